"""Device-resident submap for the reference's scan-to-submap step (an OPTIONAL seam: it replaces a Python list inside
slam.py, so using it means touching three lines there -- INTEGRATION.md section 3).

The reference keeps ``submap_buffer``, a list of the last ``submap_size`` global-frame scans (slam.py:559-562), rebuilds
``submap = _build_submap(submap_buffer, submap_voxel)`` from it in every scan (slam.py:103-108, 503) and registers the
new scan against that array (``ICP(scan, submap, ...)``, slam.py:217-225), which downsamples the ~50k-point target once
more and builds a KD-tree on it.  :class:`DeviceSubmap` is that list on the GPU:

    submap_buffer = DeviceSubmap(submap_size)          # was: submap_buffer = []
    submap_buffer.append(global_points)                # unchanged; evicts the oldest scan by itself (pop(0) is a no-op)
    R, t, err = submap_buffer.ICP(points, submap_voxel, error_threshold, max_iterations, voxel_size,
                                  R_init=R_init, t_init=t_init, method="point_to_point", max_corr_dist=d)

``np.asarray(submap_buffer.build(voxel))`` still yields the reference's ``_build_submap`` array bit for bit when a caller
wants it (e.g. for ``_submap_rotation_search``)."""
import os
import sys

import numpy as np

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)

from icp_b200 import api as _api          # noqa: E402
from icp_b200 import _lib as _abi         # noqa: E402


class DeviceSubmap(_api.DeviceSubmap):
    def pop(self, index=0):
        """The reference pops the oldest scan once the list is longer than ``submap_size`` (slam.py:561-562); the device
        window evicts it on ``append``, so this only keeps that line of slam.py harmless."""
        if index != 0:
            raise IndexError("DeviceSubmap only drops its oldest scan")
        return None

    def ICP(self, source, submap_voxel, error_threshold, max_iterations, voxel_size, R_init=None, t_init=None,
            method="point_to_point", normal_k=10, max_corr_dist=None):
        """``utilities.icp.ICP(source, _build_submap(window, submap_voxel), ...)``: same return tuple and console line."""
        have = R_init is not None and t_init is not None
        out = self.icp(np.ascontiguousarray(source, dtype=np.float64), submap_voxel, error_threshold, max_iterations, voxel_size,
                       R_init=np.asarray(R_init, dtype=np.float64)[None] if have else None,
                       t_init=np.asarray(t_init, dtype=np.float64)[None] if have else None,
                       method=method, normal_k=normal_k, max_corr_dist=max_corr_dist)
        error = float(out["error"][0])
        if int(out["status"][0]) == _abi.CONVERGED:
            delta = abs(float(out["prev_error"][0]) - error)
            print(f"  ICP converged: iter={int(out['iters'][0]) - 1}, error={error:.8f}, delta={delta:.2e}")
        else:
            print(f"  ICP max iterations reached: iter={max_iterations}, error={error:.8f}")
        return out["R"][0], out["t"][0], error
