"""The reference's OWN slam.py, unmodified, driven against the drop-in shim.

`oracle/_ref/reference_py.tar.gz` (made by oracle/make_ref.py in the build container, shipped by gpurun; git-ignored) is
unpacked to a temporary directory; the shim package `iterative-closest-point-avmi_b200/` goes AHEAD of it on sys.path
(second option), so `slam.py`'s own imports (slam.py:6-16) bind `utilities.icp.ICP`,
`utilities.icp.voxel_downsample`, `utilities.mapping.OccupancyGrid2D` and `utilities.pose_graph.PoseGraph2D` to the
library's versions while `feature_based_alignment` and `services.*` stay the reference's.  `slam.run_slam(cfg)` then runs
on the synthetic lidar CSV the golden fixture was made from (oracle/make_slam_golden.py ran the same slam.py on the
reference's own numpy/scipy utilities) and the trajectory and the map are compared.

Three configurations, each with its own golden run of the reference: the plain scan-to-scan loop; the same with the
rotation-search pre-alignment (slam.py:60-66 -> the shim's GPU rotation_search); and submap + loop closure enabled
(slam.py:103-225, 230-277, 564-620: _build_submap, _submap_rotation_search, gated ICP against the submap, loop-closure
candidates, the pose graph (the library's block-skyline solver against the golden run's dense solve), _rebuild_map)."""
import contextlib
import io
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import PKG, ROOT, load_golden

pytestmark = pytest.mark.gpu

DRIVER = textwrap.dedent('''
    import contextlib, io, json, os, sys, types
    import numpy as np
    pkg, root, ref, csv, out, mode_patch = sys.argv[1:7]
    sys.path[:0] = [pkg, ref, root]
    sys.modules["pyvista"] = types.ModuleType("pyvista")          # slam.py:4 imports it at module top; display only
    sys.dont_write_bytecode = True
    import slam                                                    # the reference's file
    assert os.path.dirname(os.path.abspath(slam.__file__)) == os.path.abspath(ref)
    assert slam.ICP.__module__ == "utilities.icp" and "iterative-closest-point-avmi_b200" in sys.modules["utilities.icp"].__file__
    assert "iterative-closest-point-avmi_b200" in sys.modules["utilities.mapping"].__file__
    assert "iterative-closest-point-avmi_b200" in sys.modules["utilities.pose_graph"].__file__
    cfg = dict(data_file=csv, imu=dict(enabled=False),
               icp=dict(method="point_to_line", normal_k=12, voxel_size=0.04, error_threshold=1e-10,
                        max_iterations=150, error_reject_threshold=0.15),
               features=dict(method="none"), submap=dict(enabled=False), loop_closure=dict(enabled=False),
               filter=dict(z_min=1.0, z_max=1.4),
               mapping=dict(resolution=0.05, margin=50.0, p_hit=0.85, p_miss=0.42, log_odds_min=-8.0, log_odds_max=8.0),
               service=dict(sleep_s=0.0, loop=False), display=dict(live_map=False), num_scans=None)
    for key, val in json.loads(mode_patch).items():
        cfg[key] = val
    log = io.StringIO()
    with contextlib.redirect_stdout(log):
        pose, traj, mapper = slam.run_slam(cfg)
    lo = mapper.log_odds
    nz = np.flatnonzero(lo)
    np.savez(out, trajectory=np.asarray(traj), final_pose=pose, grid_shape=np.asarray(lo.shape),
             grid_bounds=np.asarray([mapper.min_x, mapper.max_x, mapper.min_y, mapper.max_y]),
             nz_index=nz.astype(np.int64), nz_value=lo.ravel()[nz], log_lines=np.asarray(len(log.getvalue().splitlines())))
''')


def _run(mode, tmp_path):
    from oracle import ref_loader
    ref = ref_loader.reference_root()
    if ref is None:
        pytest.skip("oracle/_ref/reference_py.tar.gz is absent (run oracle/make_ref.py in the build container)")
    g = load_golden("slam_loop.npz")
    off = g["scan_off"]
    csv = tmp_path / "lidar.csv"
    with open(csv, "w") as f:                                   # services/lidar_service.py:5-19 format, as make_slam_golden.py
        for i in range(len(off) - 1):
            vals = ";".join(f"{float(p[0])!r};{float(p[1])!r};1.2" for p in g["scans"][off[i]:off[i + 1]])
            f.write(f"{1000000 * i};{vals}\n")
    out = tmp_path / f"{mode}.npz"
    drv = tmp_path / "drive_slam.py"
    drv.write_text(DRIVER)
    import json
    from oracle.make_slam_golden import MODES             # the configurations the golden fixture was made with
    proc = subprocess.run([sys.executable, str(drv), PKG, ROOT, ref, str(csv), str(out), json.dumps(MODES.get(mode, {}))],
                          capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-4000:]
    return g, np.load(out)


def test_unmodified_slam_scan_to_scan_matches_reference_run(tmp_path):
    g, got = _run("plain", tmp_path)
    want = g["trajectory"]
    assert got["trajectory"].shape == want.shape
    for k, (a, b) in enumerate(zip(got["trajectory"], want)):
        assert np.abs(a[:2, 2] - b[:2, 2]).max() < 1e-4, k                                       # metres
        assert abs(np.arctan2(a[1, 0], a[0, 0]) - np.arctan2(b[1, 0], b[0, 0])) < 1e-5, k        # radians
    assert np.abs(got["final_pose"] - g["final_pose"]).max() < 1e-7
    assert tuple(got["grid_shape"]) == tuple(g["grid_shape"]) and np.array_equal(got["grid_bounds"], g["grid_bounds"])
    ref_map = np.zeros(tuple(g["grid_shape"]), dtype=np.float32)
    ref_map.ravel()[g["nz_index"]] = g["nz_value"]
    mine = np.zeros_like(ref_map)
    mine.ravel()[got["nz_index"]] = got["nz_value"]
    differ = np.count_nonzero(mine != ref_map)
    assert differ <= 20, f"{differ} of {ref_map.size} cells differ from the reference run's map"


def _compare_mode(g, got, mode):
    want = g[f"trajectory_{mode}"]
    assert got["trajectory"].shape == want.shape
    for k, (a, b) in enumerate(zip(got["trajectory"], want)):
        assert np.abs(a[:2, 2] - b[:2, 2]).max() < 1e-4, (mode, k)
        assert abs(np.arctan2(a[1, 0], a[0, 0]) - np.arctan2(b[1, 0], b[0, 0])) < 1e-5, (mode, k)
    assert tuple(got["grid_shape"]) == tuple(g[f"grid_shape_{mode}"])
    ref_map = np.zeros(tuple(g[f"grid_shape_{mode}"]), dtype=np.float32)
    ref_map.ravel()[g[f"nz_index_{mode}"]] = g[f"nz_value_{mode}"]
    mine = np.zeros_like(ref_map)
    mine.ravel()[got["nz_index"]] = got["nz_value"]
    differ = np.count_nonzero(mine != ref_map)
    assert differ <= 40, f"{mode}: {differ} of {ref_map.size} cells differ from the reference run's map"


def test_unmodified_slam_with_rotation_search_prealignment(tmp_path):
    """slam.py:60-66 -> the shim's GPU rotation_search in front of every ICP call."""
    g, got = _run("rotation", tmp_path)
    _compare_mode(g, got, "rotation")


def test_unmodified_slam_with_submap_and_loop_closure(tmp_path):
    """slam.py:103-225 (submap build, rotation search around the prediction, gated ICP against the submap), 230-268 +
    564-597 (loop-closure candidates registered one by one), 599-620 (pose graph, submap and map rebuilt)."""
    g, got = _run("submap_lc", tmp_path)
    _compare_mode(g, got, "submap_lc")
