#!/usr/bin/env python
"""Small end-to-end pass of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "iterative-closest-point-avmi_b200"))
sys.path.insert(0, ROOT)
from icp_b200 import api, synth  # noqa: E402
from utilities import OccupancyGrid2D, voxel_downsample  # noqa: E402

scans, poses = synth.make_sequence(6, world="room", seed=2)
cfg = dict(error_threshold=1e-10, max_iterations=40, voxel_size=0.04, method="point_to_line", normal_k=12)
out = api.icp_batch(scans[:-1], scans[1:], **cfg)                               # K1, K2, K3 (brute)
out2 = api.icp_batch(scans[:-1], scans[1:], nn_mode="grid", **cfg)              # big voxel? no: grid build + grid pairs
assert np.array_equal(out["iters"], out2["iters"])
big = np.vstack([synth.to_world_frame(s, p) for s, p in zip(scans, poses)])     # 6k points: radix-sort voxel kernel
ds = voxel_downsample(np.vstack([big, big + 0.013]), 0.04)
api.icp_batch([scans[0]], [np.vstack([big, big + 0.013])], 1e-10, 30, 0.04, method="point_to_point", max_corr_dist=1.5)
teapot = np.random.default_rng(0).normal(size=(400, 3))
api.icp_batch([teapot + 0.01], [teapot], 1e-12, 30, 0.05)                       # 3-D Kabsch
g = OccupancyGrid2D(-25.6, 25.6, -25.6, 25.6, resolution=0.05, p_hit=0.85, p_miss=0.42, log_odds_min=-8, log_odds_max=8)
g.update_scans([p[:2] for p in poses], [synth.to_world_frame(s, p) for s, p in zip(scans, poses)])
# device-resident entry point (offsets checked on the device, deferred statistics, eight-lane replay)
import torch  # noqa: E402
hits = [synth.to_world_frame(s, p) for s, p in zip(scans, poses)]
flat, off = synth.pack_ragged(hits)
d_org, d_hits, d_off = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (poses[:, :2].copy(), flat, off))
torch.cuda.synchronize()
for _ in range(2):
    g._dev.update_dev(len(off) - 1, d_org.data_ptr(), d_hits.data_ptr(), d_off.data_ptr(), int(off[-1]), 0)
g._invalidate()
st = g._dev.last_stats()
# pairs out of order through the host-buffer entry point (pair grouping by upload chunk)
cl, co = synth.pack_ragged(list(scans))
si = np.array([4, 0, 2, 1, 3], dtype=np.int32)
out3 = api.icp_pairs(cl, co, si, si + 1, **cfg)
assert np.array_equal(out3["iters"], out["iters"][si])
print("sanitize smoke ok", out["iters"].tolist(), len(ds), float(g.log_odds.min()), float(g.log_odds.max()))
