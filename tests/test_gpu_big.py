"""GPU parity for clouds beyond one CTA's shared memory: the global-memory
voxel downsample (radix sort), the hash-grid nearest neighbour (scan -> submap,
slam.py:217-225) and the grid mode forced on scan-sized clouds."""
import numpy as np
import pytest

from conftest import load_golden, pose_delta
from icp_b200 import api, synth
from oracle import icp_oracle

pytestmark = pytest.mark.gpu

POS_TOL, ROT_TOL = 1e-4, 1e-5          # north_star tolerances
CFG = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, method="point_to_line", normal_k=12)
GATE = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, method="point_to_point", max_corr_dist=1.5)


def test_big_voxel_downsample_bit_exact():
    from utilities import voxel_downsample
    g = load_golden("submap.npz")
    out = voxel_downsample(g["raw_sub"], 0.04)                     # 21k points: radix-sort kernel
    assert out.tobytes() == g["submap"].tobytes()
    sub = synth.submap_cloud(n_raw=52000, seed=3)                  # SURVEY 8(d) C3 recipe
    for v in (0.04, 0.5):
        out = voxel_downsample(sub, v)
        want = icp_oracle.voxel_means(sub, v)
        assert out.shape == want.shape and out.tobytes() == want.tobytes(), v
    p3 = np.random.default_rng(1).normal(size=(30000, 3))
    assert voxel_downsample(p3, 0.15).tobytes() == icp_oracle.voxel_means(p3, 0.15).tobytes()


def test_golden_scan_to_submap():
    g = load_golden("submap.npz")
    for tag in ("a", "b"):
        for mode in ("auto", "grid"):
            tr = api.icp_trace(g[f"{tag}/src"], g["submap"], R_init=g[f"{tag}/R_init"], t_init=g[f"{tag}/t_init"],
                               nn_mode=mode, trace_iters=1, **GATE)
            assert tr["status"] == int(g[f"{tag}/status"]) and tr["iters"] == int(g[f"{tag}/iters"]), (tag, mode)
            dt, dr = pose_delta(tr["R"], tr["t"], g[f"{tag}/R"], g[f"{tag}/t"])
            assert dt < POS_TOL and dr < ROT_TOL, (tag, mode, dt, dr)
            assert np.array_equal(tr["matches"][0], g[f"{tag}/first_match"]), (tag, mode)
    # raw 21k-point target: big voxel pass + hash grid inside the call
    tr = api.icp_trace(g["a/src"], g["raw_sub"], R_init=g["a/R_init"], t_init=g["a/t_init"], trace_iters=1, **GATE)
    assert tr["iters"] == int(g["raw/iters"]) and tr["status"] == int(g["raw/status"])
    assert len(tr["tgt"]) == int(g["raw/n_tgt"]) and np.array_equal(tr["matches"][0], g["raw/first_match"])
    dt, dr = pose_delta(tr["R"], tr["t"], g["raw/R"], g["raw/t"])
    assert dt < POS_TOL and dr < ROT_TOL
    # point-to-line against the submap (normals on the hash grid)
    tr = api.icp_trace(g["a/src"], g["submap"], R_init=g["a/R_init"], t_init=g["a/t_init"], nn_mode="grid",
                       trace_iters=1, **CFG)
    assert tr["iters"] == int(g["p2l/iters"]) and tr["status"] == int(g["p2l/status"])
    dt, dr = pose_delta(tr["R"], tr["t"], g["p2l/R"], g["p2l/t"])
    assert dt < POS_TOL and dr < ROT_TOL


def c3_case(n_sources=4, seed=3):
    """~50k-point submap target (random wall segments) and 1080-point scans cut out of it."""
    target = synth.submap_cloud(n_raw=52000, seed=seed)
    rng = np.random.default_rng(seed + 100)
    clouds, R0, t0 = [target], [], []
    while len(clouds) < 1 + n_sources:
        c = target[rng.integers(len(target))]
        near = target[np.hypot(*(target - c).T) < 12.0]
        if len(near) < 1500:
            continue
        pts = near[rng.choice(len(near), 1080, replace=False)] + rng.normal(0, 0.01, size=(1080, 2))
        th = rng.uniform(-0.3, 0.3)
        rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        shift = rng.uniform(-5, 5, size=2)
        clouds.append((pts - shift) @ rot)                       # sensor frame: p_world = rot p + shift
        a = th + 0.01                                            # perturbed initial guess (slam.py:209-215)
        R0.append([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        t0.append(shift + [0.05, -0.04])
    return clouds, np.asarray(R0), np.asarray(t0)


def test_large_submap_vs_oracle():
    """~47k-voxel target after downsampling; several scans against the one target through the pairs API."""
    clouds, R0, t0 = c3_case()
    flat, off = synth.pack_ragged(clouds)
    out = api.icp_pairs(flat, off, [1, 2, 3, 4], [0, 0, 0, 0], R_init=R0, t_init=t0, **GATE)
    assert len(icp_oracle.voxel_means(clouds[0], 0.04)) > 40000
    for p in range(4):
        R, t, err, iters, status = icp_oracle.register(clouds[1 + p], clouds[0], R_init=R0[p], t_init=t0[p], **GATE)
        assert out["status"][p] == status and out["iters"][p] == iters, (p, out["iters"][p], iters)
        dt, dr = pose_delta(out["R"][p], out["t"][p], R, t)
        assert dt < POS_TOL and dr < ROT_TOL, (p, dt, dr)
        assert abs(out["error"][p] - err) < 1e-9


def test_grid_mode_equals_brute_on_scans():
    scans, _ = synth.make_sequence(8, world="room", seed=13)
    brute = api.icp_batch(scans[:-1], scans[1:], nn_mode="brute", **CFG)
    grid = api.icp_batch(scans[:-1], scans[1:], nn_mode="grid", **CFG)
    assert np.array_equal(brute["iters"], grid["iters"]) and np.array_equal(brute["status"], grid["status"])
    assert np.abs(brute["R"] - grid["R"]).max() < 1e-9 and np.abs(brute["t"] - grid["t"]).max() < 1e-9
