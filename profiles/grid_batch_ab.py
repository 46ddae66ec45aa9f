#!/usr/bin/env python
"""Scan -> submap batch (C3 shape: 1080-point scans against one 47k-voxel submap, hash-grid nearest neighbour) with
more pairs than SMs, for A/B runs of the grid-mode pair kernel's register budget."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402,F401
from icp_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
target = synth.submap_cloud(n_raw=52000, seed=3)
rng = np.random.default_rng(103)
clouds, R0, t0s = [target], [], []
while len(clouds) < 1 + n:
    c = target[rng.integers(len(target))]
    near = target[np.hypot(*(target - c).T) < 12.0]
    if len(near) < 1500:
        continue
    pts = near[rng.choice(len(near), 1080, replace=False)] + rng.normal(0, 0.01, size=(1080, 2))
    th = rng.uniform(-0.3, 0.3)
    rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    shift = rng.uniform(-5, 5, size=2)
    clouds.append((pts - shift) @ rot)
    a = th + 0.01
    R0.append([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
    t0s.append(shift + [0.05, -0.04])
flat, off = synth.pack_ragged(clouds)
kw = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, method="point_to_point", max_corr_dist=1.5,
          R_init=np.asarray(R0), t_init=np.asarray(t0s))
si, ti = np.arange(1, n + 1, dtype=np.int32), np.zeros(n, dtype=np.int32)
api.icp_pairs(flat, off, si, ti, **kw)
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    res = api.icp_pairs(flat, off, si, ti, **kw)
    ts.append(time.perf_counter() - t0)
ks = api.icp_last_stats()
knobs = {k: v for k, v in os.environ.items() if k.startswith("ICPB200_")}
print(f"{knobs}: {n} pairs, {n / min(ts):.0f} reg/s e2e, pair kernel {ks['pair_kernel_ns'] / 1e6:.3f} ms, iters sum {int(res['iters'].sum())}")
