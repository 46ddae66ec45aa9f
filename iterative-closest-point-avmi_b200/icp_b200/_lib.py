"""ctypes binding of libicp_b200.so (C ABI: include/icp_b200.h).

The library is the product; this module only declares prototypes and turns
error codes into exceptions.  There is deliberately no fallback: if the shared
library is missing the import fails, and if no CUDA device is usable every
compute call raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libicp_b200.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_int64_p = ctypes.POINTER(ctypes.c_int64)

OK, ERR_CUDA, ERR_ARG, ERR_LIMIT = 0, -1, -2, -3
CONVERGED, MAX_ITER, FEW_INLIERS, BAD_VOXELS, SINGULAR = 0, 1, 2, 3, 4
POINT_TO_POINT, POINT_TO_LINE = 0, 1
NN_AUTO, NN_BRUTE, NN_GRID = 0, 1, 2

_ICP_TAIL = [ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int,
             ctypes.c_double, ctypes.c_int]
_ICP_OUT = [c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p, c_int32_p]

# name -> (restype, argtypes); mirrors include/icp_b200.h one to one
PROTOTYPES = {
    "icpb200_init": (ctypes.c_int, [ctypes.c_int]),
    "icpb200_shutdown": (None, []),
    "icpb200_last_error": (ctypes.c_char_p, []),
    "icpb200_launch_count": (ctypes.c_int64, []),
    "icpb200_built_arch": (ctypes.c_int, []),
    "icpb200_icp_batch": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_double_p, c_int64_p, c_double_p,
                                         c_int64_p, c_double_p, c_double_p] + _ICP_TAIL + _ICP_OUT),
    "icpb200_icp_pairs": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_double_p, c_int64_p, ctypes.c_int,
                                         c_int32_p, c_int32_p, c_double_p, c_double_p] + _ICP_TAIL + _ICP_OUT),
    "icpb200_icp_pairs_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p] + _ICP_TAIL +
                              [ctypes.c_void_p] * 6 + [ctypes.c_void_p]),
    "icpb200_icp_trace": (ctypes.c_int, [ctypes.c_int, c_double_p, ctypes.c_int64, c_double_p, ctypes.c_int64,
                                         c_double_p, c_double_p] + _ICP_TAIL + _ICP_OUT +
                          [c_double_p, c_int64_p, c_double_p, c_int64_p, c_double_p, c_int32_p, ctypes.c_int]),
    "icpb200_icp_last_stats": (ctypes.c_int, [c_int64_p]),
    "icpb200_icp_phase_profile": (ctypes.c_int, [c_int64_p]),
    "icpb200_icp_pair_profile": (ctypes.c_int, [c_int64_p, ctypes.c_int64]),
    "icpb200_icp_extra_stats": (ctypes.c_int, [c_int64_p]),
    "icpb200_voxel_downsample": (ctypes.c_int, [c_double_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double,
                                                c_double_p, c_int64_p]),
    "icpb200_grid_create": (ctypes.c_void_p, [ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 7),
    "icpb200_grid_destroy": (None, [ctypes.c_void_p]),
    "icpb200_grid_set_shard": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "icpb200_grid_update": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p, c_double_p, c_int64_p]),
    "icpb200_grid_update_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "icpb200_grid_read": (ctypes.c_int, [ctypes.c_void_p, c_float_p]),
    "icpb200_grid_read_view": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_float_p, c_int32_p]),
    "icpb200_grid_reset": (ctypes.c_int, [ctypes.c_void_p]),
    "icpb200_grid_rebuild": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p, c_double_p, c_int64_p]),
    "icpb200_grid_ipc_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p]),
    "icpb200_grid_ipc_attach": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]),
    "icpb200_grid_push_tiles": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "icpb200_grid_device_ptr": (ctypes.c_void_p, [ctypes.c_void_p]),
    "icpb200_grid_tile_profile": (ctypes.c_int, [ctypes.c_void_p, c_int64_p, ctypes.c_int64]),
    "icpb200_grid_last_stats": (ctypes.c_int, [ctypes.c_void_p, c_int64_p]),
    "icpb200_submap_create": (ctypes.c_void_p, [ctypes.c_int, ctypes.c_int]),
    "icpb200_submap_destroy": (None, [ctypes.c_void_p]),
    "icpb200_submap_push": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.c_int64]),
    "icpb200_submap_clear": (ctypes.c_int, [ctypes.c_void_p]),
    "icpb200_submap_size": (ctypes.c_int, [ctypes.c_void_p, c_int64_p, c_int64_p]),
    "icpb200_submap_build": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, c_double_p, ctypes.c_int64, c_int64_p]),
    "icpb200_submap_icp": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_int, c_double_p, c_int64_p, c_double_p,
                                          c_double_p] + _ICP_TAIL + _ICP_OUT),
    "icpb200_rotation_scores": (ctypes.c_int, [ctypes.c_int, c_double_p, c_int64_p, c_double_p, c_int64_p, c_double_p,
                                               c_int64_p, c_double_p, c_double_p, c_double_p, c_int32_p]),
    "icpb200_pose_graph_optimize": (ctypes.c_int, [ctypes.c_int64, c_double_p, ctypes.c_int64, c_int32_p, c_int32_p, c_double_p,
                                                   c_double_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, c_int32_p,
                                                   c_double_p, c_int32_p]),
    "icpb200_pin_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t]),
    "icpb200_unpin_host": (ctypes.c_int, [ctypes.c_void_p]),
}

_lib = None


def load():
    """Load libicp_b200.so once; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C iterative-closest-point-avmi_b200/csrc` "
                "(or __graft_entry__.build()).  There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().icpb200_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != OK:
        kind = {ERR_CUDA: "CUDA failure", ERR_ARG: "invalid argument", ERR_LIMIT: "limit exceeded"}.get(rc, f"error {rc}")
        raise RuntimeError(f"{what}: {kind}: {last_error()}")
