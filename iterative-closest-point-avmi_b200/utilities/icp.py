"""GPU-backed ``utilities.icp`` -- same call surface as the reference.

``ICP`` and ``voxel_downsample`` keep the reference's names, positional order,
defaults, return types and the one console line per call
(/root/reference/utilities/icp.py:132-134, 117, 218, 222); the arithmetic runs
in libicp_b200.so on the GPU.  No numpy/scipy fallback exists: without the
library or a CUDA device these functions raise.
"""
import os
import sys

import numpy as np

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)

from icp_b200 import api as _api          # noqa: E402
from icp_b200 import _lib as _abi         # noqa: E402


def voxel_downsample(points, voxel_size):
    """Voxel-grid mean; rows in lexicographic voxel order (icp.py:117-129)."""
    return _api.voxel_downsample(np.asarray(points, dtype=np.float64), voxel_size)


def ICP(source, target, error_threshold, max_iterations, voxel_size,
        R_init=None, t_init=None, method="point_to_point", normal_k=10,
        max_corr_dist=None):
    """Iterative Closest Point on the GPU; returns ``(R, t, error)``.

    ``p' = R p + t`` maps source onto target.  ``method`` is
    ``"point_to_point"`` or ``"point_to_line"`` (2-D only; 3-D data silently
    uses point-to-point, icp.py:162).  ``max_corr_dist`` gates correspondences
    (icp.py:183-189); ``None`` disables the gate.
    """
    src = np.ascontiguousarray(source, dtype=np.float64)
    tgt = np.ascontiguousarray(target, dtype=np.float64)
    have_init = R_init is not None and t_init is not None
    out = _api.icp_batch(
        [src], [tgt], error_threshold, max_iterations, voxel_size,
        R_init=np.asarray(R_init, dtype=np.float64)[None] if have_init else None,
        t_init=np.asarray(t_init, dtype=np.float64)[None] if have_init else None,
        method=method, normal_k=normal_k, max_corr_dist=max_corr_dist)
    error = float(out["error"][0])
    if int(out["status"][0]) == _abi.CONVERGED:
        delta = abs(float(out["prev_error"][0]) - error)
        print(f"  ICP converged: iter={int(out['iters'][0]) - 1}, error={error:.8f}, delta={delta:.2e}")
    else:
        print(f"  ICP max iterations reached: iter={max_iterations}, error={error:.8f}")
    return out["R"][0], out["t"][0], error
