#!/usr/bin/env python
"""Wall-clock split of the host-buffer ICP entry point (needs ICPB200_E2E_TIMING=1; GPU box)."""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo/iterative-closest-point-avmi_b200'); sys.path.insert(0,'/root/repo')
import bench
from icp_b200 import api
scans, poses, flat, off, si, ti = bench.build_c2(2000, seed=0)
pin = api.pinned(flat, off, si, ti)
for k in range(6):
    t0=time.perf_counter(); out = api.icp_pairs(flat, off, si, ti, **bench.ICP_CFG); print("wall %.0f us" % ((time.perf_counter()-t0)*1e6), file=sys.stderr)
st = api.icp_last_stats()
print("event spans (ms): mark->K1/K2 done %.3f, ->K2 done %.3f, pairs %.3f" % (st["voxel_kernel_ns"] / 1e6, st["normals_kernel_ns"] / 1e6, st["pair_kernel_ns"] / 1e6), file=sys.stderr)
