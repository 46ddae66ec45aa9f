"""Locate the UNMODIFIED reference for tests and the CPU-baseline legs -- TEST INFRASTRUCTURE ONLY.

`reference_root()` returns a directory holding the reference's own `slam.py`, `utilities/`, `services/`:
the archive `oracle/_ref/reference_py.tar.gz` (made by `oracle/make_ref.py` in the build container, shipped to the
GPU box by gpurun) unpacked into a temporary directory, else `/root/reference` itself, else None.
`import_reference(*names)` imports modules from it with a stub `pyvista` (utilities/mapping.py:2 imports it at
module top; it is display-only and absent from the image) WITHOUT leaving the reference on sys.path or in
sys.modules, so the drop-in shim's `utilities` package is unaffected.
"""
import atexit
import importlib
import os
import shutil
import sys
import tarfile
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ARCHIVE = os.path.join(HERE, "_ref", "reference_py.tar.gz")
_root = None


def reference_root():
    global _root
    if _root is not None:
        return _root or None
    if os.path.exists(ARCHIVE):
        tmp = tempfile.mkdtemp(prefix="icpb200_ref_")
        with tarfile.open(ARCHIVE, "r:gz") as tar:
            tar.extractall(tmp, filter="data")
        atexit.register(shutil.rmtree, tmp, ignore_errors=True)
        _root = tmp
    elif os.path.isdir("/root/reference/utilities"):
        _root = "/root/reference"
    else:
        _root = ""
    return _root or None


def import_reference(*names):
    """Import the reference's own modules (e.g. "utilities.icp", "utilities.mapping") in isolation."""
    root = reference_root()
    if root is None:
        raise ImportError("the reference is not available (no oracle/_ref archive, no /root/reference)")
    saved = {k: v for k, v in sys.modules.items() if k == "utilities" or k.startswith("utilities.") or k == "pyvista"
             or k == "services" or k.startswith("services.") or k == "slam"}
    for k in saved:
        del sys.modules[k]
    sys.modules["pyvista"] = types.ModuleType("pyvista")
    sys.path.insert(0, root)
    dont = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        mods = [importlib.import_module(n) for n in names]
    finally:
        sys.dont_write_bytecode = dont
        sys.path.remove(root)
        for k in [k for k in sys.modules if k == "utilities" or k.startswith("utilities.") or k == "pyvista"
                  or k == "services" or k.startswith("services.") or k == "slam"]:
            del sys.modules[k]
        sys.modules.update(saved)
    return mods if len(mods) > 1 else mods[0]
