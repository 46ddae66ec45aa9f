#!/usr/bin/env python
"""Where a nearly-converged ICP iteration spends its cycles (C2 batch, GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "iterative-closest-point-avmi_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from icp_b200 import _lib, api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
scans, poses, flat, off, si, ti = bench.build_c2(n, seed=0)
for _ in range(2):
    out = api.icp_pairs(flat, off, si, ti, **bench.ICP_CFG)
st = api.icp_last_stats()
ph = np.zeros(8, dtype=np.int64)
_lib.load().icpb200_icp_phase_profile(ph.ctypes.data_as(_lib.c_int64_p))
it = max(int(ph[5]), 1)
print("kernel ms: voxel %.3f normals %.3f pairs %.3f" % (st["voxel_kernel_ns"] / 1e6, st["normals_kernel_ns"] / 1e6, st["pair_kernel_ns"] / 1e6))
print("iterations >= 8:", it, " total iterations:", st["iterations"])
for name, v in zip(("classify", "nearest-nb", "accumulate", "solve", "apply+err"), ph[:5]):
    print(f"  {name:12s} {v / it:9.0f} cycles/iteration")
print("  sum          %9.0f cycles/iteration" % (ph[:5].sum() / it))
print("stats:", {k: st[k] for k in st if k not in ("voxel_kernel_ns", "normals_kernel_ns", "pair_kernel_ns")})
