"""A/B of the cluster mode of the hand-over launch (ICPB200_NO_CLUSTER=1 switches it off; read once per process): a
rank-sized share of the C5 batch (few expensive pairs -> each gets a cluster of 4 CTAs).  Saves the results for a bitwise
comparison between the two runs.  Usage: python profiles/cluster_ab.py <tag> [n_pairs] [first]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "iterative-closest-point-avmi_b200"), ROOT]
from icp_b200 import api, synth
from icp_b200 import dist as icpd
tag = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
scans, poses = synth.make_sequence(2000, world="room", seed=0)
flat, off = synth.pack_ragged(scans)
pairs = synth.loop_closure_pairs(poses, 8192, seed=0, max_dist=3.0).astype(np.int32)
si, ti = pairs[:, 0].copy(), pairs[:, 1].copy()
mine = icpd.plan_pair_shards(si, ti, 8192 // n)[rank]          # the share rank `rank` of 8192 / n ranks would get
si, ti = si[mine], ti[mine]
cfg = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, method="point_to_line", normal_k=12)
api.init(0)
pin = api.pinned(flat, off)
api.icp_pair_profile(0)
best = 1e9
for rep in range(4):
    t0 = time.perf_counter()
    out = api.icp_pairs(flat, off, si, ti, **cfg)
    best = min(best, time.perf_counter() - t0)
st = api.icp_last_stats()
prof = api.icp_pair_profile(len(si))
print(f"{tag}: {len(si)} pairs, call {best * 1e3:.2f} ms, hand-over launch(es) {st['pair_kernel_ns'] / 1e6:.3f} ms, longest pair {prof[:, 0].max() / 1.965e3:.0f} us, "
      f"max-iter pairs {(out['status'] == 1).sum()}, swept {st['points_swept']}, {api.icp_extra_stats()}")
ph = api.icp_phase_profile()
print("   phases(it>=8):", {k: round(v) for k, v in ph["phases"].items()})
print("   top pairs: us | iterations | swept/it | fp64 fallbacks/it | cycles per iteration >= 8: classify, NN, rest")
for i in np.argsort(-prof[:, 0])[:8]:
    it = max(prof[i, 3], 1); it8 = max(prof[i, 3] - 8, 1)
    print(f"     {prof[i, 0] / 1.965e3:6.0f} | {prof[i, 3]:3d} | {prof[i, 1] / it:5.0f} | {prof[i, 2] / it:5.1f} | "
          f"{prof[i, 4] / it8:6.0f} {prof[i, 5] / it8:6.0f} {prof[i, 6] / it8:6.0f}")
np.savez(os.path.join(ROOT, "gpurun_out", f"cluster_ab_{tag}.npz"), **out)
pin.release()
