"""Drop-in replacements for the reference's hot-path modules.

Overlay ``icp.py`` and ``mapping.py`` on the reference's ``utilities`` package
(or put this directory's parent ahead of it on ``sys.path``): ``slam.py`` and
``demos/teapot_icp_demo.py`` import ``utilities.icp.ICP``,
``utilities.icp.voxel_downsample`` and ``utilities.mapping.OccupancyGrid2D``
by these names (slam.py:8-10, demos/teapot_icp_demo.py:23).
"""
from .icp import ICP, voxel_downsample
from .mapping import OccupancyGrid2D
from .features import rotation_search, submap_rotation_search

__all__ = ["ICP", "voxel_downsample", "OccupancyGrid2D", "rotation_search", "submap_rotation_search"]
