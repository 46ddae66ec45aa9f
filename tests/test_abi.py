"""The C-ABI library loads without a GPU, exports every symbol the header
declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "icp_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(icpb200_\w+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from icp_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 27
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/icp_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names      # the ctypes binding covers the whole header
    assert lib.icpb200_built_arch() == 100


def test_sass_is_sm100():
    """The shared library carries sm_100a code (cuobjdump is in the image)."""
    import shutil
    import subprocess
    from icp_b200 import _lib
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([tool, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from icp_b200 import api
    from utilities import ICP, OccupancyGrid2D, voxel_downsample
    pts = np.random.default_rng(0).normal(size=(50, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ICP(pts, pts, 1e-7, 10, 0.1)
    with pytest.raises(RuntimeError):
        voxel_downsample(pts, 0.1)
    with pytest.raises(RuntimeError):
        OccupancyGrid2D(-1, 1, -1, 1)
    with pytest.raises(RuntimeError):
        api.icp_pairs(pts, [0, 50], [0], [0], 1e-7, 10, 0.1)


def test_argument_errors_do_not_need_a_gpu():
    from icp_b200 import _lib
    lib = _lib.load()
    rc = lib.icpb200_voxel_downsample(None, 0, 2, 0.1, None, None)
    assert rc == _lib.ERR_ARG and b"bad argument" in lib.icpb200_last_error()
    out = (ctypes.c_double * 4)()
    iters = (ctypes.c_int32 * 1)()
    off = (ctypes.c_int64 * 2)(0, 5)
    pts = (ctypes.c_double * 10)()
    rc = lib.icpb200_icp_batch(1, 4, pts, off, pts, off, None, None, 1e-7, 10, 0.1, 0, 10, -1.0, 0,
                               out, out, out, out, iters, iters)
    assert rc == _lib.ERR_ARG and b"dim must be 2 or 3" in lib.icpb200_last_error()
    rc = lib.icpb200_icp_batch(1, 2, pts, off, pts, off, None, None, 1e-7, 10, -1.0, 0, 10, -1.0, 0,
                               out, out, out, out, iters, iters)
    assert rc == _lib.ERR_ARG and b"voxel_size" in lib.icpb200_last_error()
    rc = lib.icpb200_icp_batch(1, 2, pts, off, pts, off, None, None, 1e-7, 10, 0.1, 1, 64, -1.0, 0,
                               out, out, out, out, iters, iters)
    assert rc == _lib.ERR_LIMIT
