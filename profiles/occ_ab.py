#!/usr/bin/env python
"""A/B of occupancy variants selected by environment knobs: the device-resident C4 update, CUDA events on the stream,
L2 flushed between steps (the same loop as bench.py's occupancy leg on one GPU)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from icp_b200 import api  # noqa: E402
from utilities import OccupancyGrid2D  # noqa: E402

api.init(0)
dev = torch.device("cuda", 0)
origins, flat, off = bench.build_c4(2000, seed=0)
grid = OccupancyGrid2D(*bench.GRID_BOUNDS, **bench.GRID_CFG)
d_org, d_hits, d_off = (torch.from_numpy(x).to(dev) for x in (origins, flat, off))
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ms = []
for k in range(15):
    grid.reset()
    flush.zero_()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    grid._dev.update_dev(len(off) - 1, d_org.data_ptr(), d_hits.data_ptr(), d_off.data_ptr(), int(off[-1]), stream.cuda_stream)
    b.record(stream)
    torch.cuda.synchronize()
    if k >= 5:
        ms.append(a.elapsed_time(b))
knobs = {k: v for k, v in os.environ.items() if k.startswith("ICPB200_")}
print(f"{knobs}: update {np.mean(ms) * 1e3:.1f} us (min {np.min(ms) * 1e3:.1f}), checksum {float(np.abs(grid.log_odds).sum()):.3f}")
