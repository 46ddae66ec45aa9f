"""Pin oracle/features_oracle.py to the live reference and write tests/golden/rotation.npz.

Run in the BUILD container only (it reads /root/reference):

    python oracle/pin_rotation.py

``rotation_search`` is imported unmodified from /root/reference/utilities/features.py.
``_submap_rotation_search`` lives in slam.py, whose module top imports pyvista and the
services; its FunctionDef is cut out of the unmodified source with ``ast`` and executed
with the reference's own ``voxel_downsample``.  Every case must agree BIT-FOR-BIT with
the oracle before the fixture is written.
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import pin_against_reference as pin                                   # noqa: E402  (loader + helpers)
from icp_b200 import synth                                             # noqa: E402
from oracle import features_oracle as fo                               # noqa: E402


def load_reference_functions():
    import importlib.util
    ref_icp, _ = pin.load_reference()
    spec = importlib.util.spec_from_file_location("refutil.features", os.path.join(pin.REF, "utilities", "features.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["refutil.features"] = mod
    spec.loader.exec_module(mod)
    tree = ast.parse(open(os.path.join(pin.REF, "slam.py")).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "_submap_rotation_search")
    ns = {"np": np, "voxel_downsample": ref_icp.voxel_downsample}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), os.path.join(pin.REF, "slam.py"), "exec"), ns)
    return mod.rotation_search, ns["_submap_rotation_search"]


def rot(theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, -s], [s, c]])


def main():
    ref_rs, ref_sub = load_reference_functions()
    scans, poses = synth.make_sequence(60, world="room", seed=3)
    out = {}
    # ---- rotation_search: consecutive scans, a large artificial rotation, the reference's config values
    cases = [("cfg", scans[10], scans[11], dict(voxel_size=0.15, angle_step_coarse=1.5, angle_step_fine=0.1)),
             ("defaults", scans[20], scans[24], {}),
             ("turned", scans[30] @ rot(2.1).T + [0.4, -0.3], scans[31], dict(voxel_size=0.15, angle_step_coarse=1.5, angle_step_fine=0.1)),
             ("tiny", scans[5][:40], scans[6][:40], dict(voxel_size=2.0))]
    for name, s, t, kw in cases:
        R0, t0, sc0 = pin.quiet(ref_rs, s, t, **kw)
        R1, t1, sc1, coarse, fine, fine_angles = fo.rotation_search(s, t, **kw)
        pin.check(pin.same_bits(R0, R1) and pin.same_bits(t0, t1) and np.float64(sc0).tobytes() == np.float64(sc1).tobytes(),
                  f"rotation_search {name}: oracle == reference bit-for-bit (score {sc0:.6g})")
        out[f"rs_{name}_src"], out[f"rs_{name}_tgt"] = s, t
        out[f"rs_{name}_kw"] = np.array([kw.get("voxel_size", 0.3), kw.get("angle_step_coarse", 2.0), kw.get("angle_step_fine", 0.2)])
        out[f"rs_{name}_R"], out[f"rs_{name}_t"], out[f"rs_{name}_score"] = R0, t0, np.float64(sc0)
        if coarse is not None:
            out[f"rs_{name}_coarse"], out[f"rs_{name}_fine"] = coarse, fine
    # ---- _submap_rotation_search: a scan against the 25 scans before it, predicted pose off by a few degrees
    for name, k, dth, dxy, kw in [("cfg", 40, 0.06, (0.08, -0.05), dict(angle_range=60.0, angle_step=0.8, fine_step=0.1, voxel_size=0.2)),
                                  ("defaults", 55, -0.2, (0.0, 0.1), {})]:
        submap = np.vstack([synth.to_world_frame(scans[i], poses[i]) for i in range(k - 25, k)])
        x, y, th = poses[k]
        P = np.eye(3)
        P[:2, :2] = rot(th + dth)
        P[:2, 2] = [x + dxy[0], y + dxy[1]]
        R0, t0 = pin.quiet(ref_sub, scans[k], submap, P, **kw)
        R1, t1 = fo.submap_rotation_search(scans[k], submap, P, **kw)
        pin.check(pin.same_bits(R0, R1) and pin.same_bits(t0, t1), f"_submap_rotation_search {name}: oracle == reference bit-for-bit")
        out[f"sub_{name}_src"], out[f"sub_{name}_map"], out[f"sub_{name}_pose"] = scans[k], submap, P
        out[f"sub_{name}_kw"] = np.array([kw.get("angle_range", 60.0), kw.get("angle_step", 2.0), kw.get("fine_step", 0.5), kw.get("voxel_size", 0.3)])
        out[f"sub_{name}_R"], out[f"sub_{name}_t"] = R0, t0
    path = os.path.join(pin.GOLDEN, "rotation.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
