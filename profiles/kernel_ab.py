#!/usr/bin/env python
"""A/B of compile-time variants selected by environment knobs: kernel times of the device-resident C2 step
(CUDA events inside the library, L2 flushed between steps).  Usage: ENV=... python profiles/kernel_ab.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from icp_b200 import _lib, api  # noqa: E402

api.init(0)
lib = _lib.load()
dev = torch.device("cuda", 0)
scans, poses, flat, off, si, ti = bench.build_c2(2000, seed=0)
n_pairs, max_pts, cfg = len(si), int(np.max(np.diff(off))), bench.ICP_CFG
d = [torch.from_numpy(x).to(dev) for x in (flat, off, si, ti)]
out = [torch.empty(s, dtype=t, device=dev) for s, t in (((n_pairs, 2, 2), torch.float64), ((n_pairs, 2), torch.float64),
       ((n_pairs,), torch.float64), ((n_pairs,), torch.float64), ((n_pairs,), torch.int32), ((n_pairs,), torch.int32))]
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
rows, steps = [], []
for k in range(13):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    _lib.check(lib.icpb200_icp_pairs_dev(len(off) - 1, 2, d[0].data_ptr(), d[1].data_ptr(), max_pts, n_pairs, d[2].data_ptr(),
               d[3].data_ptr(), None, None, cfg["error_threshold"], cfg["max_iterations"], cfg["voxel_size"],
               _lib.POINT_TO_LINE, cfg["normal_k"], -1.0, _lib.NN_AUTO, *[o.data_ptr() for o in out], stream.cuda_stream), "dev")
    b.record(stream)
    torch.cuda.synchronize()
    st = api.icp_last_stats()
    if k >= 3:
        rows.append((st["voxel_kernel_ns"] / 1e3, st["normals_kernel_ns"] / 1e3, st["pair_kernel_ns"] / 1e3))
        steps.append(a.elapsed_time(b) * 1e3)
r = np.array(rows).mean(axis=0)
knobs = {k: v for k, v in os.environ.items() if k.startswith("ICPB200_")}
print(f"{knobs}: voxel {r[0]:.0f} us  normals {r[1]:.0f} us  pairs {r[2]:.0f} us  step {np.mean(steps):.0f} us  iters sum {int(out[4].sum())}")
