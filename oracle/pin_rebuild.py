"""Pin oracle.occupancy_oracle.rebuild_map / transform_points_2d against the reference's own slam._rebuild_map and
slam.transform_points_2d (imported unmodified from /root/reference; build container only) and write
tests/golden/rebuild.npz.  The history holds 14 scans with their poses, one empty scan, one scan whose endpoints leave
the grid, and poses that change between two rebuilds (what a loop closure does, slam.py:596-603)."""
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, "iterative-closest-point-avmi_b200"))
sys.path.insert(0, ROOT)
from icp_b200 import synth  # noqa: E402
from oracle import occupancy_oracle as oo  # noqa: E402

GKW = dict(resolution=0.05, p_hit=0.85, p_miss=0.42, log_odds_min=-8.0, log_odds_max=8.0)
BOUNDS = (-22.0, 22.0, -14.0, 14.0)


def pose_matrix(x, y, th):
    c, s = np.cos(th), np.sin(th)
    return np.array([[c, -s, x], [s, c, y], [0.0, 0.0, 1.0]])


def main():
    sys.modules["pyvista"] = types.ModuleType("pyvista")
    sys.path.insert(0, REF)
    import slam                                                   # the reference, unmodified
    from utilities.mapping import OccupancyGrid2D as RefGrid
    scans, poses = synth.make_sequence(14, world="room", seed=23, traj_seed=5)
    scans = [s.copy() for s in scans]
    scans[4] = np.zeros((0, 2))                                    # an empty scan
    scans[9] = scans[9] * 1.8                                      # endpoints beyond the grid
    rng = np.random.default_rng(7)
    out = {}
    for variant in range(2):                                       # second pass: corrected poses, same scans
        P = poses + (rng.normal(0, [0.05, 0.05, 0.01], size=poses.shape) if variant else 0.0)
        history = [(s, pose_matrix(*p)) for s, p in zip(scans, P)]
        ref = RefGrid(*BOUNDS, **GKW)
        ref.update_scan(np.zeros(2), scans[0])                     # something to clear
        with contextlib.redirect_stdout(io.StringIO()):
            slam._rebuild_map(ref, history)
        for cls in (oo.GridOracleC, oo.GridOraclePy):
            mine = cls(*BOUNDS, **GKW)
            mine.update_scan(np.zeros(2), scans[0])
            oo.rebuild_map(mine, history)
            assert mine.log_odds.tobytes() == ref.log_odds.tobytes(), (variant, cls.__name__)
        for s, M in history[:3]:
            assert oo.transform_points_2d(s, M).tobytes() == slam.transform_points_2d(s, M).tobytes()
        nz = np.flatnonzero(ref.log_odds)
        out[f"poses_{variant}"] = np.stack([M for _, M in history])
        out[f"nz_index_{variant}"] = nz.astype(np.int64)
        out[f"nz_value_{variant}"] = ref.log_odds.ravel()[nz]
        print(f"variant {variant}: {len(nz)} non-zero cells, oracle == reference (C and Python restatements)")
    flat, off = synth.pack_ragged(scans)
    path = os.path.join(ROOT, "tests", "golden", "rebuild.npz")
    np.savez_compressed(path, scans=flat, scan_off=off, bounds=np.asarray(BOUNDS), grid_shape=np.asarray(ref.log_odds.shape), **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
