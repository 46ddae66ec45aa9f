"""Where the C5 batch (8192 loop-closure candidate pairs, identity start, up to 3 m / 0.5 rad apart) spends its time:
kernel times and work counters for a few iteration caps, on one GPU.  Usage: python profiles/c5_probe.py [n_pairs]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "iterative-closest-point-avmi_b200"), ROOT]
from icp_b200 import api, synth  # noqa: E402

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
scans, poses = synth.make_sequence(2000, world="room", seed=0)
flat, off = synth.pack_ragged(scans)
pairs = synth.loop_closure_pairs(poses, n_pairs, seed=0, max_dist=3.0).astype(np.int32)
si, ti = pairs[:, 0].copy(), pairs[:, 1].copy()
d = np.hypot(*(poses[si, :2] - poses[ti, :2]).T)
dth = np.abs((poses[si, 2] - poses[ti, 2] + np.pi) % (2 * np.pi) - np.pi)
print(f"pairs {n_pairs}: distance mean {d.mean():.2f} max {d.max():.2f} m; heading difference mean {dth.mean():.3f} max {dth.max():.3f} rad")
cfg = dict(error_threshold=1e-10, voxel_size=0.04, method="point_to_line", normal_k=12)
api.init(0)
for cap in (1, 2, 4, 8, 12, 24, 150):
    api.icp_pairs(flat, off, si, ti, max_iterations=cap, **cfg)
    t0 = time.perf_counter()
    out = api.icp_pairs(flat, off, si, ti, max_iterations=cap, **cfg)
    dt = time.perf_counter() - t0
    st = api.icp_last_stats()
    ph = api.icp_phase_profile()
    it = out["iters"]
    print(f"max_iterations {cap:3d}: call {dt * 1e3:7.2f} ms, K3 {st['pair_kernel_ns'] / 1e6:7.2f} ms, iterations {st['iterations']}, "
          f"swept {st['points_swept']} carried {st['points_carried']} evals {st['sweep_pair_evals'] / 1e6:.0f} M rescans {st['fp64_rescans']}; "
          f"converged {(out['status'] == 0).sum()} ; phases(it>=8) {{{', '.join(f'{k}: {v:.0f}' for k, v in ph['phases'].items())}}}")
hist = np.bincount(np.minimum(it, 150) // 10)
print("iterations histogram (bins of 10):", hist.tolist())
err = out["error"]
print("final error quantiles:", np.quantile(err[np.isfinite(err)], [0.1, 0.5, 0.9, 0.99]).tolist())
slow = out["status"] == 1
print(f"iteration-limit pairs: {slow.sum()}; their distance mean {d[slow].mean():.2f}, heading diff mean {dth[slow].mean():.3f}, error median {np.median(err[slow]):.4f}")

# per-pair cost of the full run: which pairs hold the tail?
api.icp_pair_profile(0)
out = api.icp_pairs(flat, off, si, ti, max_iterations=150, **cfg)
prof = api.icp_pair_profile(n_pairs)
cyc = prof[:, 0]
order = np.argsort(-cyc)
print(f"per-pair cycles: total {cyc.sum() / 1e9:.2f} G, max {cyc.max() / 1e6:.2f} M ({cyc.max() / 1.965e3:.0f} us), "
      f"top 10 share {cyc[order[:10]].sum() / cyc.sum():.2f}, top 100 share {cyc[order[:100]].sum() / cyc.sum():.2f}")
for p in order[:12]:
    print(f"  pair {p}: {cyc[p] / 1.965e3:8.0f} us, iters {prof[p, 3]}, swept {prof[p, 1]} ({prof[p, 1] / max(prof[p, 3], 1):.0f}/it), "
          f"rescans {prof[p, 2]} ({prof[p, 2] / max(prof[p, 3], 1):.1f}/it), err {out['error'][p]:.3f}, dist {d[p]:.2f} m, dth {dth[p]:.3f}")
slow = out["status"] == 1
print(f"iteration-limit pairs: cycles median {np.median(cyc[slow]) / 1.965e3:.0f} us, p90 {np.quantile(cyc[slow], 0.9) / 1.965e3:.0f} us; "
      f"converged pairs: median {np.median(cyc[~slow]) / 1.965e3:.0f} us, p90 {np.quantile(cyc[~slow], 0.9) / 1.965e3:.0f} us")
