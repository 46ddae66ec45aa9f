"""Kernel times + work counters of the C2 and C5 batches on one GPU (device time from the library's own CUDA events).
Usage: python profiles/icp_quick.py [c2] [c5]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "iterative-closest-point-avmi_b200"), ROOT]
from icp_b200 import api, synth  # noqa: E402

which = [a for a in sys.argv[1:] if a in ("c2", "c5")] or ["c2", "c5"]
scans, poses = synth.make_sequence(2000, world="room", seed=0)
flat, off = synth.pack_ragged(scans)
cfg = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, method="point_to_line", normal_k=12)
api.init(0)
pin = api.pinned(flat, off)
for name in which:
    if name == "c2":
        si = np.arange(1999, dtype=np.int32); ti = si + 1
    else:
        pairs = synth.loop_closure_pairs(poses, 8192, seed=0, max_dist=3.0).astype(np.int32)
        si, ti = pairs[:, 0].copy(), pairs[:, 1].copy()
    api.icp_pair_profile(0)
    best = None
    for rep in range(4):
        t0 = time.perf_counter()
        out = api.icp_pairs(flat, off, si, ti, **cfg)
        dt = time.perf_counter() - t0
        st = api.icp_last_stats()
        if best is None or dt < best[0]:
            best = (dt, st)
    dt, st = best
    ph = api.icp_phase_profile()
    prof = api.icp_pair_profile(len(si))
    cyc = prof[:, 0]
    print(f"{name}: call {dt * 1e3:.2f} ms | voxel {st['voxel_kernel_ns'] / 1e6:.3f} normals {st['normals_kernel_ns'] / 1e6:.3f} pairs {st['pair_kernel_ns'] / 1e6:.3f} ms | "
          f"iterations {st['iterations']} swept {st['points_swept']} evals {st['sweep_pair_evals'] / 1e6:.0f} M fallbacks {st['fp64_rescans']} | "
          f"CTA cycles total {cyc.sum() / 1e9:.2f} G, longest pair {cyc.max() / 1.965e3:.0f} us | "
          f"phases(it>=8) {{{', '.join(f'{k}: {v:.0f}' for k, v in ph['phases'].items())}}} | max-iter pairs {(out['status'] == 1).sum()}")
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"icp_quick_{name}.npz"), si=si, ti=ti, prof=prof, **out)
pin.release()
