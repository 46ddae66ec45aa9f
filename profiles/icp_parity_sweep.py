#!/usr/bin/env python
"""Iteration counts / poses of a C2 sub-batch against the CPU oracle (GPU box; parity aid)."""
import os
import sys
import multiprocessing as mp

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "iterative-closest-point-avmi_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from icp_b200 import api  # noqa: E402


def one(args):
    from oracle import icp_oracle
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        out = icp_oracle.register(args[0], args[1], **bench.ICP_CFG)
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    scans, poses, flat, off, si, ti = bench.build_c2(2000, seed=0)
    out = api.icp_pairs(flat, off, si, ti, **bench.ICP_CFG)
    sel = np.linspace(0, len(si) - 1, n).astype(int)
    # always include the pairs that hit the iteration limit
    sel = np.unique(np.concatenate([sel, np.nonzero(out["iters"] >= 150)[0][:60]]))
    with mp.get_context("spawn").Pool(min(16, os.cpu_count())) as pool:
        ref = pool.map(one, [(scans[si[i]], scans[ti[i]]) for i in sel], chunksize=4)
    bad = 0
    for k, i in enumerate(sel):
        R, t, err, it = ref[k][0], ref[k][1], ref[k][2], int(ref[k][3])
        dt = float(np.abs(out["t"][i] - t).max())
        dth = abs(np.arctan2(out["R"][i][1, 0], out["R"][i][0, 0]) - np.arctan2(R[1, 0], R[0, 0]))
        if it != int(out["iters"][i]) or dt > 1e-4 or dth > 1e-5:
            bad += 1
            if bad <= 10:
                print(f"pair {i}: iters gpu {int(out['iters'][i])} ref {it}  dt {dt:.2e} dtheta {dth:.2e}")
    print(f"checked {len(sel)} pairs, {bad} differ; total gpu iterations {int(out['iters'].sum())}")
