"""Wall-clock split of the occupancy e2e step (C4): icpb200_grid_update from page-locked host buffers, then the read-out."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "iterative-closest-point-avmi_b200"), ROOT]
import bench
from icp_b200 import api
from utilities import OccupancyGrid2D
api.init(0)
origins, flat, off = bench.build_c4(2000, seed=0)
grid = OccupancyGrid2D(*bench.GRID_BOUNDS, **bench.GRID_CFG)
host_out = np.zeros((grid.ny, grid.nx), dtype=np.float32)
pin = api.pinned(host_out, flat, origins, off)
for mode in ("dirty", "full"):
    tu, tr = [], []
    for k in range(6):
        grid.reset()
        t0 = time.perf_counter()
        grid._dev.update(origins, flat, off)
        t1 = time.perf_counter()
        if mode == "dirty":
            grid._dev.read_dirty(host_out)
        else:
            grid._dev.read(host_out)
        t2 = time.perf_counter()
        if k >= 2:
            tu.append(t1 - t0); tr.append(t2 - t1)
    print(f"{mode}: update {np.mean(tu) * 1e3:.3f} ms (H2D {flat.nbytes / 1e6:.1f} MB), read {np.mean(tr) * 1e3:.3f} ms, tiles {grid._dev.last_tiles_copied}")
pin.release()
