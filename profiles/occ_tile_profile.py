#!/usr/bin/env python
"""Per-tile timing of the occupancy tile kernel on the C4 workload (GPU box)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "iterative-closest-point-avmi_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from icp_b200 import _lib  # noqa: E402
from utilities import OccupancyGrid2D  # noqa: E402

origins, flat, off = bench.build_c4(2000, 0)
grid = OccupancyGrid2D(*bench.GRID_BOUNDS, **bench.GRID_CFG)
lib = _lib.load()
lib.icpb200_grid_tile_profile(grid._dev._h, None, 0)
for _ in range(2):
    grid.reset()
    grid._dev.update(origins, flat, off)
from icp_b200.dist import TILE  # noqa: E402
n_tiles = ((grid.nx + TILE - 1) // TILE) * ((grid.ny + TILE - 1) // TILE)
full = np.zeros((2 * n_tiles, 4), dtype=np.int64)
n = lib.icpb200_grid_tile_profile(grid._dev._h, full.ctypes.data_as(_lib.c_int64_p), 2 * n_tiles)
buf, ph = full[:n_tiles], full[n_tiles:]
act = buf[buf[:, 3] > 0]
print("tiles written", n, "active", len(act))
order = np.argsort(-act[:, 3])
print("total cycles (sum over tiles) %.1fM ; max tile %.2fM ; mean %.2fM" % (act[:, 3].sum() / 1e6, act[:, 3].max() / 1e6, act[:, 3].mean() / 1e6))
print("  tile  scans    runs   Mcycles  cyc/scan  cyc/run")
for i in order[:12]:
    t, s, r, c = act[i]
    print(f"{t:6d} {s:6d} {r:7d} {c / 1e6:9.3f} {c / max(s, 1):9.0f} {c / max(r, 1):8.1f}")
sel = act[act[:, 1] > 50]
print("cycles per scan-group: median %.0f  p10 %.0f  p90 %.0f" % tuple(np.percentile(sel[:, 3] / sel[:, 1], [50, 10, 90])))

print("thread-0 phase split of the 6 slowest tiles (Mcycles): count, barrier1, apply, barrier2")
for i in order[:6]:
    t = act[i, 0]
    print(f"{t:6d}", " ".join(f"{v / 1e6:8.3f}" for v in ph[t]))
