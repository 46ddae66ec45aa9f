"""Drop-in check at the level of the reference's main loop: the call sequence
of slam.py:377-557 (scan-to-scan ICP -> pose accumulation -> update_scan),
restated here, runs on the shim and is compared with the trajectory and map
the UNMODIFIED reference slam.py produced on the same synthetic lidar file
(oracle/make_slam_golden.py, run in the build container)."""
import contextlib
import io

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_slam_main_loop_matches_reference():
    from utilities import ICP, OccupancyGrid2D
    g = load_golden("slam_loop.npz")
    off = g["scan_off"]
    scans = [g["scans"][off[i]:off[i + 1]] for i in range(len(off) - 1)]
    gkw = dict(resolution=0.05, p_hit=0.85, p_miss=0.42, log_odds_min=-8.0, log_odds_max=8.0)

    def new_mapper(points):                                   # slam.py:31-36, 398-405
        return OccupancyGrid2D(float(points[:, 0].min() - 50.0), float(points[:, 0].max() + 50.0),
                               float(points[:, 1].min() - 50.0), float(points[:, 1].max() + 50.0), **gkw)

    pose = np.eye(3)
    traj, prev, mapper = [], None, None
    with contextlib.redirect_stdout(io.StringIO()):
        for points in scans:
            if prev is None:
                prev, mapper = points, new_mapper(points)
                mapper.update_scan(pose[:2, 2], points @ pose[:2, :2].T + pose[:2, 2])
                continue
            r, t, err = ICP(prev, points, error_threshold=1e-10, max_iterations=150, voxel_size=0.04,
                            R_init=None, t_init=None, method="point_to_line", normal_k=12)      # slam.py:90-98
            if err > 0.15:                                    # slam.py:485-490
                prev = points
                continue
            t_inv = np.eye(3)                                 # slam.py:38-43
            t_inv[:2, :2] = r.T
            t_inv[:2, 2] = -r.T @ t
            pose = pose @ t_inv
            traj.append(pose.copy())
            mapper.update_scan(pose[:2, 2], points @ pose[:2, :2].T + pose[:2, 2])              # slam.py:552-557
            prev = points
    want = g["trajectory"]
    assert len(traj) == len(want)
    for k, (a, b) in enumerate(zip(traj, want)):
        assert np.abs(a[:2, 2] - b[:2, 2]).max() < 1e-4, k                                       # metres
        assert abs(np.arctan2(a[1, 0], a[0, 0]) - np.arctan2(b[1, 0], b[0, 0])) < 1e-5, k        # radians
    assert np.abs(traj[-1] - g["final_pose"]).max() < 1e-7
    assert mapper.log_odds.shape == tuple(g["grid_shape"])
    assert np.array_equal([mapper.min_x, mapper.max_x, mapper.min_y, mapper.max_y], g["grid_bounds"])
    ref_map = np.zeros(mapper.log_odds.shape, dtype=np.float32)
    ref_map.ravel()[g["nz_index"]] = g["nz_value"]
    differ = np.count_nonzero(mapper.log_odds != ref_map)
    # poses agree to ~1e-12, so at most a stray endpoint on a cell border may move
    assert differ <= 20, f"{differ} of {ref_map.size} cells differ from the reference map"

    # the raycast alone, fed the reference's own poses, must be bit-exact
    exact = new_mapper(scans[0])
    poses = [np.eye(3)] + list(want)
    used = [scans[0]]
    prev_pose_count = 0
    # scans that the reference skipped (error gate) have no pose; replay only the mapped ones
    mapped = [0]
    k = 0
    with contextlib.redirect_stdout(io.StringIO()):
        prev = scans[0]
        for i, points in enumerate(scans[1:], start=1):
            r, t, err = ICP(prev, points, 1e-10, 150, 0.04, method="point_to_line", normal_k=12)
            prev = points
            if err > 0.15:
                continue
            mapped.append(i)
    assert len(mapped) == len(poses)
    exact.update_scans([p[:2, 2] for p in poses],
                       [scans[i] @ p[:2, :2].T + p[:2, 2] for i, p in zip(mapped, poses)])
    assert exact.log_odds.tobytes() == ref_map.tobytes()
