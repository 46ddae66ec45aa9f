"""One rank's share of the sharded C4 update, on ONE GPU: set_shard(rank, world) without any exchange reproduces exactly
the kernels rank `rank` of `world` runs.  Usage: python profiles/occ_shard_emul.py [rank] [world] [reps]"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "iterative-closest-point-avmi_b200"), ROOT]
import bench
from icp_b200 import api
from utilities import OccupancyGrid2D
rank = int(sys.argv[1]) if len(sys.argv) > 1 else 0
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
api.init(0)
dev = torch.device("cuda", 0)
origins, flat, off = bench.build_c4(2000, seed=0)
grid = OccupancyGrid2D(*bench.GRID_BOUNDS, **bench.GRID_CFG)
if world > 1:
    grid._dev.set_shard(rank, world)
d_org, d_hits, d_off = (torch.from_numpy(x).to(dev) for x in (origins, flat, off))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ms = []
for k in range(reps):
    grid.reset()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    grid._dev.update_dev(len(off) - 1, d_org.data_ptr(), d_hits.data_ptr(), d_off.data_ptr(), int(off[-1]), stream.cuda_stream)
    b.record(stream)
    torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
print(f"rank {rank} of {world}: update {np.median(ms[2:]) * 1e3:.1f} us (median of {len(ms) - 2}), stats {grid._dev.last_stats()}")
