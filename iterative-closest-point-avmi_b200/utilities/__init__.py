"""Drop-in replacements for the reference's hot-path modules.

Two ways to use them: overlay ``icp.py``, ``mapping.py``, ``features.py`` and
``pose_graph.py`` on the reference's ``utilities`` package, or put this directory's parent AHEAD of the reference root on
``sys.path``.  In the second case this package shadows the reference's ``utilities``; what it does not replace (the
feature pipeline behind ``feature_based_alignment``, slam.py:9) is found through ``__path__``: every other
``utilities`` directory on ``sys.path`` is appended to it, so the unmodified ``slam.py`` imports cleanly.
"""
import os as _os
import sys as _sys

_HERE = _os.path.dirname(_os.path.abspath(__file__))


def _fall_through_dirs():
    out = []
    for p in _sys.path:
        d = _os.path.abspath(_os.path.join(p or ".", "utilities"))
        if d != _HERE and d not in out and _os.path.isfile(_os.path.join(d, "__init__.py")):
            out.append(d)
    return out


for _d in _fall_through_dirs():
    if _d not in __path__:
        __path__.append(_d)

from .icp import ICP, voxel_downsample                                      # noqa: E402
from .mapping import OccupancyGrid2D                                        # noqa: E402
from .features import rotation_search, submap_rotation_search               # noqa: E402
from .features import feature_based_alignment                               # noqa: E402  (the reference's, when it is on the path)
from .submap import DeviceSubmap                                            # noqa: E402

__all__ = ["ICP", "voxel_downsample", "OccupancyGrid2D", "rotation_search", "submap_rotation_search",
           "feature_based_alignment", "DeviceSubmap"]

from .pose_graph import (PoseGraph2D, pose_matrix_to_vec, pose_vec_to_matrix,      # noqa: E402,F401  (utilities/__init__.py:4-9)
                         relative_transform_vec)

__all__ += ["PoseGraph2D", "pose_matrix_to_vec", "pose_vec_to_matrix", "relative_transform_vec"]
