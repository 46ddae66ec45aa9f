#!/usr/bin/env python
"""Stage times of one rank's share of the C4 update for a given world size (run on ONE GPU, rank by rank;
needs ICPB200_OCC_TIMING=1)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "iterative-closest-point-avmi_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from utilities import OccupancyGrid2D  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
origins, flat, off = bench.build_c4(2000, 0)
for rank in range(world):
    grid = OccupancyGrid2D(*bench.GRID_BOUNDS, **bench.GRID_CFG)
    grid._dev.set_shard(rank, world)
    for k in range(3):
        grid.reset()
        print(f"--- rank {rank}/{world} pass {k}", file=sys.stderr, flush=True)
        grid._dev.update(origins, flat, off)
    st = grid._dev.last_stats()
    print(f"rank {rank}: traversed {st['traversed']} runs {st['runs']} hits {st['hits']}", file=sys.stderr)
