"""Run the reference's slam.py (unmodified, imported from /root/reference) on a
synthetic lidar CSV and store its trajectory and final map as a golden fixture.

Build container only.  Configuration: scan-to-scan point_to_line ICP with the
config.yaml ICP / mapping parameters, no IMU, no pre-alignment
(features.method "none"), submap and loop closure off, live_map off -- i.e.
BASELINE.json configs[1] run through the reference's own main loop
(slam.py:282-657).  tests/test_gpu_slam_loop.py replays the same call sequence
against the drop-in shim and compares pose by pose and cell by cell.
"""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, "iterative-closest-point-avmi_b200"))
from icp_b200 import synth  # noqa: E402

N_SCANS = 30
MODES = {
    "rotation": dict(features=dict(method="rotation_search", rotation_voxel_size=0.15, angle_step_coarse=1.5, angle_step_fine=0.1)),
    "submap_lc": dict(submap=dict(enabled=True, size=10, voxel_size=0.04, max_corr_dist=1.5, rotation_range=20.0,
                                  rotation_step=2.0, rotation_fine_step=0.5, rotation_voxel_size=0.25),
                      loop_closure=dict(enabled=True, distance_threshold=3.0, min_interval=10, max_candidates=2,
                                        min_cumulative_travel=0.5, error_threshold=0.3)),
}


def main():
    scans, poses = synth.make_sequence(N_SCANS, world="room", seed=17, traj_seed=11)
    tmp = tempfile.mkdtemp()
    csv = os.path.join(tmp, "lidar.csv")
    with open(csv, "w") as f:                                   # services/lidar_service.py:5-19 format
        for i, s in enumerate(scans):
            vals = ";".join(f"{float(p[0])!r};{float(p[1])!r};1.2" for p in s)
            f.write(f"{1000000 * i};{vals}\n")
    cfg = dict(data_file=csv, imu=dict(enabled=False),
               icp=dict(method="point_to_line", normal_k=12, voxel_size=0.04, error_threshold=1e-10,
                        max_iterations=150, error_reject_threshold=0.15),
               features=dict(method="none"), submap=dict(enabled=False), loop_closure=dict(enabled=False),
               filter=dict(z_min=1.0, z_max=1.4),
               mapping=dict(resolution=0.05, margin=50.0, p_hit=0.85, p_miss=0.42, log_odds_min=-8.0, log_odds_max=8.0),
               service=dict(sleep_s=0.0, loop=False), display=dict(live_map=False), num_scans=None)
    sys.modules["pyvista"] = types.ModuleType("pyvista")
    sys.path.insert(0, REF)
    import slam                                                  # the reference, unmodified
    with contextlib.redirect_stdout(io.StringIO()):
        global_pose, trajectory, mapper = slam.run_slam(cfg)
    lo = mapper.log_odds
    nz = np.flatnonzero(lo)
    print(f"{len(trajectory)} poses, grid {lo.shape}, {len(nz)} non-zero cells, final pose\n{global_pose}")
    flat, off = synth.pack_ragged(scans)
    extra = {}
    # two more configurations of the same main loop (tests/test_gpu_slam_unmodified.py runs them on the shim):
    # the rotation-search pre-alignment in front of every ICP call (slam.py:60-66), and submap alignment + loop
    # closure + map rebuild (slam.py:103-225, 230-277, 564-620)
    for mode, patch in MODES.items():
        c2 = {k: (dict(v) if isinstance(v, dict) else v) for k, v in cfg.items()}
        for key, val in patch.items():
            c2[key] = val
        log = io.StringIO()
        with contextlib.redirect_stdout(log):
            pose_m, traj_m, mapper_m = slam.run_slam(c2)
        lo_m = mapper_m.log_odds
        nz_m = np.flatnonzero(lo_m)
        text = log.getvalue()
        print(f"[{mode}] {len(traj_m)} poses, {len(nz_m)} non-zero cells, loop closures accepted: {text.count('Loop closure accepted') + text.count('ACCEPTED')}")
        extra.update({f"trajectory_{mode}": np.asarray(traj_m), f"final_pose_{mode}": pose_m,
                      f"grid_shape_{mode}": np.asarray(lo_m.shape), f"nz_index_{mode}": nz_m.astype(np.int64),
                      f"nz_value_{mode}": lo_m.ravel()[nz_m]})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "slam_loop.npz"), scans=flat, scan_off=off,
                        trajectory=np.asarray(trajectory), final_pose=global_pose,
                        grid_shape=np.asarray(lo.shape), grid_bounds=np.asarray([mapper.min_x, mapper.max_x, mapper.min_y, mapper.max_y]),
                        nz_index=nz.astype(np.int64), nz_value=lo.ravel()[nz], **extra)
    print("wrote tests/golden/slam_loop.npz", os.path.getsize(os.path.join(ROOT, "tests", "golden", "slam_loop.npz")))


if __name__ == "__main__":
    main()
