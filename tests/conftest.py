"""Shared pytest plumbing: path setup, the ``gpu`` marker, fixture loaders."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "iterative-closest-point-avmi_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def icp_cases(npz):
    """Group the flat ``case/key`` arrays of icp2d.npz by case."""
    cases = {}
    for key in npz.files:
        case, field = key.split("/", 1)
        cases.setdefault(case, {})[field] = npz[key]
    return cases


def case_kwargs(case):
    kw = {}
    for key, val in case.items():
        if key.startswith("kw_"):
            name = key[3:]
            if name in ("method",):
                kw[name] = str(val)
            elif name in ("max_iterations", "normal_k"):
                kw[name] = int(val)
            elif name in ("R_init", "t_init"):
                kw[name] = np.asarray(val, dtype=np.float64)
            else:
                kw[name] = float(val)
    return kw


@pytest.fixture(scope="session")
def harness():
    """Host build of the device helper headers (tests/host_harness)."""
    d = os.path.join(ROOT, "tests", "host_harness")
    subprocess.check_call(["make", "-s", "-C", d, "libharness.so"])
    lib = ctypes.CDLL(os.path.join(d, "libharness.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    lib.harness_ray_cells.restype = ctypes.c_int64
    lib.harness_ray_cells.argtypes = [ctypes.c_int] * 7 + [ctypes.POINTER(ctypes.c_int32), ctypes.c_int64]
    lib.harness_minor_steps.argtypes = [ctypes.c_int] * 5
    for fn in (lib.harness_ray_runs_iter, lib.harness_ray_runs_cb):
        fn.restype = ctypes.c_int64
        fn.argtypes = [ctypes.c_int] * 6 + [ctypes.POINTER(ctypes.c_int32), ctypes.c_int64]
    lib.harness_sat_cell.argtypes = [ctypes.c_double]
    lib.harness_solve3.argtypes = [dp, dp, dp]
    lib.harness_kabsch2.argtypes = [dp, dp]
    lib.harness_kabsch3.argtypes = [dp, dp]
    lib.harness_eigvec2.argtypes = [ctypes.c_double] * 3 + [dp]
    return lib


def rot2(theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, -s], [s, c]])


def pose_delta(R_a, t_a, R_b, t_b):
    """(translation difference in metres, rotation difference in radians)."""
    dt = float(np.max(np.abs(np.asarray(t_a) - np.asarray(t_b))))
    rel = np.asarray(R_a) @ np.asarray(R_b).T
    if rel.shape == (2, 2):
        dr = abs(float(np.arctan2(rel[1, 0], rel[0, 0])))
    else:
        dr = float(np.arccos(np.clip((np.trace(rel) - 1.0) / 2.0, -1.0, 1.0)))
    return dt, dr
