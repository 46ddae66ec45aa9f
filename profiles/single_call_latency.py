#!/usr/bin/env python
"""Latency of ONE ICP() / update_scan() call through the drop-in shim (the slam.py main-loop shape)."""
import contextlib
import io
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "iterative-closest-point-avmi_b200"))
sys.path.insert(0, ROOT)
from icp_b200 import synth  # noqa: E402
from utilities import ICP, OccupancyGrid2D, rotation_search  # noqa: E402

scans, poses = synth.make_sequence(220, world="room", seed=0)
cfg = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, method="point_to_line", normal_k=12)
with contextlib.redirect_stdout(io.StringIO()):
    for k in range(10):
        ICP(scans[k], scans[k + 1], **cfg)
    t = []
    for k in range(10, 210):
        t0 = time.perf_counter()
        ICP(scans[k], scans[k + 1], **cfg)
        t.append(time.perf_counter() - t0)
print("ICP() single call: median %.0f us, p10 %.0f, p90 %.0f" % tuple(np.percentile(np.array(t) * 1e6, [50, 10, 90])))
grid = OccupancyGrid2D(-102.4, 102.4, -102.4, 102.4, resolution=0.05, p_hit=0.85, p_miss=0.42, log_odds_min=-8, log_odds_max=8)
hits = [synth.to_world_frame(s, p) for s, p in zip(scans, poses)]
for k in range(5):
    grid.update_scan(poses[k, :2], hits[k])
t = []
for k in range(5, 205):
    t0 = time.perf_counter()
    grid.update_scan(poses[k, :2], hits[k])
    t.append(time.perf_counter() - t0)
print("update_scan() single call: median %.0f us, p10 %.0f, p90 %.0f" % tuple(np.percentile(np.array(t) * 1e6, [50, 10, 90])))
