"""GPU parity: ICP registration through the C ABI vs golden fixtures written
from the live reference, and vs the oracle on seeded inputs.

Tolerances (BASELINE.json north_star): final poses within 1e-4 m and 1e-5 rad;
correspondence indices identical (except documented exact-distance ties);
voxel-downsample output is integer/sort work plus an input-order sum, so it is
held to bit-exactness."""
import numpy as np
import pytest

from conftest import case_kwargs, icp_cases, load_golden, pose_delta
from icp_b200 import api, synth
from oracle import icp_oracle

pytestmark = pytest.mark.gpu

POS_TOL = 1e-4      # metres   (north_star)
ROT_TOL = 1e-5      # radians  (north_star)
CFG = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, method="point_to_line", normal_k=12)


def test_voxel_downsample_bit_exact():
    from utilities import voxel_downsample
    g = load_golden("voxel.npz")
    for tag in sorted({k[:-3] for k in g.files if k.endswith("_in")}):
        out = voxel_downsample(g[f"{tag}_in"], float(g[f"{tag}_v"]))
        want = g[f"{tag}_out"]
        assert out.shape == want.shape, tag
        assert out.tobytes() == want.tobytes(), f"{tag}: max diff {np.abs(out - want).max()}"
    rng = np.random.default_rng(0)
    pts = rng.uniform(-5, 5, size=(3000, 2))
    for v in (0.05, 0.7, 30.0):
        assert voxel_downsample(pts, v).tobytes() == icp_oracle.voxel_means(pts, v).tobytes()
    p3 = rng.normal(size=(1500, 3))
    assert voxel_downsample(p3, 0.2).tobytes() == icp_oracle.voxel_means(p3, 0.2).tobytes()


def test_teapot_known_answer_and_fixture():
    from utilities import ICP
    g = load_golden("teapot.npz")
    R, t, err = ICP(source=g["moved"], target=g["teapot"], error_threshold=1e-12, max_iterations=300,
                    voxel_size=0.005, method="point_to_point")
    assert R.shape == (3, 3) and t.shape == (3,) and isinstance(err, float)
    # analytic answer of demos/teapot_icp_demo.py:38-47
    assert np.abs(R - g["ry"].T).max() < 1e-9 and np.abs(t + g["ry"].T @ g["shift"]).max() < 1e-9
    dt, dr = pose_delta(R, t, g["R"], g["t"])
    assert dt < POS_TOL and dr < ROT_TOL
    assert abs(err - float(g["err"])) < 1e-12
    # 3-D "point_to_line" means point_to_point (icp.py:162)
    R2, t2, _ = ICP(g["moved"], g["teapot"], 1e-12, 300, 0.005, method="point_to_line")
    assert R2.tobytes() == R.tobytes() and t2.tobytes() == t.tobytes()


def test_golden_2d_cases():
    """Every 2-D case pinned from the reference: p2l with config.yaml and code
    default parameters, initial guess, p2p, gated scan->submap, the
    too-few-inliers break (error stays +inf) and the max-iterations exit."""
    cases = icp_cases(load_golden("icp2d.npz"))
    for name, c in cases.items():
        kw = case_kwargs(c)
        tr = api.icp_trace(c["src"], c["tgt"], trace_iters=1, **kw)
        assert tr["status"] == int(c["status"]), name
        assert tr["iters"] == int(c["iters"]), f"{name}: {tr['iters']} iterations, reference {int(c['iters'])}"
        dt, dr = pose_delta(tr["R"], tr["t"], c["R"], c["t"])
        assert dt < POS_TOL and dr < ROT_TOL, f"{name}: dt={dt:.3e} dr={dr:.3e}"
        if np.isinf(c["err"]):
            assert np.isinf(tr["error"]), name
        else:
            assert abs(tr["error"] - float(c["err"])) < 1e-9, name
        assert len(tr["src"]) == int(c["n_src"]) and len(tr["tgt"]) == int(c["n_tgt"]), name
        if len(c["first_match"]):
            assert np.array_equal(tr["matches"][0], c["first_match"]), f"{name}: first-iteration correspondences"


def test_normals_match_reference_up_to_sign():
    g = load_golden("normals.npz")
    for tag in ("scan0_k12", "scan1_k10"):
        cloud, k = g[f"{tag}_in"], int(g[f"{tag}_k"])
        # voxel far below the point spacing leaves the cloud unchanged, so the
        # trace's normals are the normals of `cloud` itself
        tr = api.icp_trace(cloud, cloud, 1e-10, 1, 1e-7, method="point_to_line", normal_k=k, trace_iters=0)
        assert len(tr["tgt"]) == len(cloud)
        order = np.lexsort((cloud[:, 1], cloud[:, 0]))     # trace rows are in voxel (lexicographic) order
        want = g[f"{tag}_out"][order]
        assert np.allclose(tr["tgt"], cloud[order], rtol=0, atol=0)
        dots = np.abs(np.sum(tr["normals"] * want, axis=1))
        assert dots.min() > 1.0 - 1e-9, f"{tag}: worst |n.n_ref| = {dots.min()}"


def test_correspondences_identical_every_iteration():
    scans, _ = synth.make_sequence(6, world="room", seed=11)
    for i in range(4):
        trace = {}
        R, t, err, iters, status = icp_oracle.register(scans[i], scans[i + 1], trace=trace, **CFG)
        tr = api.icp_trace(scans[i], scans[i + 1], trace_iters=iters, **CFG)
        assert tr["src"].tobytes() == trace["src"].tobytes() and tr["tgt"].tobytes() == trace["tgt"].tobytes()
        assert tr["iters"] == iters and tr["status"] == status
        for it in range(iters):
            same = tr["matches"][it] == trace["matches"][it]
            assert same.all(), f"pair {i} iteration {it}: {np.count_nonzero(~same)} correspondences differ"
        dt, dr = pose_delta(tr["R"], tr["t"], R, t)
        assert dt < 1e-9 and dr < 1e-9


def test_batch_equals_single_and_pairs_api():
    scans, poses = synth.make_sequence(24, world="room", seed=3)
    batch = api.icp_batch(scans[:-1], scans[1:], **CFG)
    flat, off = synth.pack_ragged(scans)
    idx = np.arange(len(scans) - 1, dtype=np.int32)
    pairs = api.icp_pairs(flat, off, idx, idx + 1, **CFG)
    for key in ("R", "t", "error", "iters", "status"):
        assert batch[key].tobytes() == pairs[key].tobytes(), key     # bitwise reproducible
    for i in (0, 7, 22):
        one = api.icp_batch([scans[i]], [scans[i + 1]], **CFG)
        assert one["R"][0].tobytes() == batch["R"][i].tobytes()
        R, t, err, iters, status = icp_oracle.register(scans[i], scans[i + 1], **CFG)
        dt, dr = pose_delta(batch["R"][i], batch["t"][i], R, t)
        assert dt < POS_TOL and dr < ROT_TOL and batch["iters"][i] == iters and batch["status"][i] == status
    # loop-closure style pairs (far apart in time, close in space), with initial guesses
    lc = synth.loop_closure_pairs(poses, 6, seed=1, max_dist=1.0, min_gap=3).astype(np.int32)
    th = poses[lc[:, 1], 2] - poses[lc[:, 0], 2]
    R0 = np.stack([[[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]] for a in -th])
    t0 = np.zeros((len(lc), 2))
    out = api.icp_pairs(flat, off, lc[:, 0], lc[:, 1], R_init=R0, t_init=t0, **CFG)
    for p, (i, j) in enumerate(lc):
        R, t, err, iters, status = icp_oracle.register(scans[i], scans[j], R_init=R0[p], t_init=t0[p], **CFG)
        dt, dr = pose_delta(out["R"][p], out["t"][p], R, t)
        if status == icp_oracle.STATUS_CONVERGED:
            assert dt < POS_TOL and dr < ROT_TOL, (p, dt, dr)
        assert out["status"][p] == status


def test_shim_prints_reference_line(capsys):
    from utilities import ICP
    scans, _ = synth.make_sequence(2, world="room", seed=4)
    R, t, err = ICP(scans[0], scans[1], 1e-10, 150, 0.04, method="point_to_line", normal_k=12)
    line = capsys.readouterr().out
    Ro, to, eo, iters, status = icp_oracle.register(scans[0], scans[1], **CFG)
    if status == icp_oracle.STATUS_CONVERGED:     # icp.py:218
        assert line.startswith(f"  ICP converged: iter={iters - 1}, error={eo:.8f}, delta=")
    else:                                         # icp.py:222
        assert line == f"  ICP max iterations reached: iter=150, error={eo:.8f}\n"
    R1, t1, err1 = ICP(scans[0], scans[1], 1e-7, 100, 0.06, method="point_to_line")   # slam.py:92-97 defaults
    assert capsys.readouterr().out.startswith("  ICP converged: iter=")
    R2, t2, err2 = ICP(scans[0], scans[1], 1e-10, 2, 0.04, method="point_to_line", normal_k=12)
    assert capsys.readouterr().out.startswith("  ICP max iterations reached: iter=2, error=")
    # one of R_init / t_init alone is ignored (icp.py:153)
    R3, t3, _ = ICP(scans[0], scans[1], 1e-10, 150, 0.04, R_init=np.eye(2) * 0.5, method="point_to_line", normal_k=12)
    assert R3.tobytes() == R.tobytes()
    # inputs are not mutated and need not be contiguous / float64
    a = np.asfortranarray(scans[0]).astype(np.float32)
    keep = a.copy()
    ICP(a, scans[1], 1e-10, 5, 0.04)
    assert np.array_equal(a, keep)


def test_limits_and_errors():
    pts = np.random.default_rng(0).normal(size=(5000, 2))
    with pytest.raises(RuntimeError, match="source clouds of more than 4096"):
        api.icp_batch([pts], [pts], 1e-7, 5, 0.01)
    with pytest.raises(RuntimeError, match="nn_mode brute supports targets"):
        api.icp_batch([pts[:100]], [pts], 1e-7, 5, 0.01, nn_mode="brute")
    with pytest.raises(RuntimeError, match="2-D only"):
        api.icp_batch([np.zeros((10, 3))], [np.random.default_rng(1).normal(size=(5000, 3))], 1e-7, 5, 0.01)
    with pytest.raises(RuntimeError):
        api.icp_batch([np.zeros((0, 2))], [pts[:10]], 1e-7, 5, 0.01)


def test_two_phase_schedule_matches_single_launch_batches():
    """Batches of >= 2 x SM-count pairs run 256-thread CTAs, two per SM, and hand slow pairs to a second launch after 12
    iterations; batches of at most one pair per SM run the 512-thread variant from start to end.  The two variants add
    their fp64 sums in different (each fixed, reproducible) orders, so results agree to rounding, not bit for bit:
    iteration counts and exit status equal, poses within 1e-10.  Two runs of the same batch ARE bitwise identical."""
    scans, _ = synth.make_sequence(40, world="room", seed=6)
    flat, off = synth.pack_ragged(scans)
    rng = np.random.default_rng(0)
    si = rng.integers(0, 38, size=640).astype(np.int32)
    ti = (si + rng.integers(1, 3, size=640)).astype(np.int32)
    big = api.icp_pairs(flat, off, si, ti, **CFG)
    again = api.icp_pairs(flat, off, si, ti, **CFG)
    for key in ("R", "t", "error", "prev_error", "iters", "status"):
        assert again[key].tobytes() == big[key].tobytes(), key
    assert big["iters"].max() > 12 and (big["status"] == 0).sum() > 500
    for lo in range(0, 640, 128):
        part = api.icp_pairs(flat, off, si[lo:lo + 128], ti[lo:lo + 128], **CFG)
        assert np.array_equal(part["iters"], big["iters"][lo:lo + 128]) and np.array_equal(part["status"], big["status"][lo:lo + 128])
        assert np.abs(part["R"] - big["R"][lo:lo + 128]).max() < 1e-10 and np.abs(part["t"] - big["t"][lo:lo + 128]).max() < 1e-10
        ok = np.isfinite(part["error"])
        assert np.abs(part["error"][ok] - big["error"][lo:lo + 128][ok]).max() < 1e-12


def test_c2_batch_sweep_against_the_oracle():
    """A slice of the headline batch (400 consecutive scans, the reference's config.yaml parameters) through the
    batch entry point against the oracle on every 6th pair plus every pair that hits the iteration limit: same
    iteration counts and exit status, poses within the north-star tolerance.  This is the regime the pair kernel's
    shortcuts live in (slab sweep, two-candidate words, net-displacement bound, hand-over launch)."""
    from oracle import icp_oracle
    scans, _ = synth.make_sequence(400, world="room", seed=0)
    cfg = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, method="point_to_line", normal_k=12)
    flat, off = synth.pack_ragged(scans)
    si = np.arange(len(scans) - 1, dtype=np.int32)
    out = api.icp_pairs(flat, off, si, si + 1, **cfg)
    assert out["iters"].max() == 150 and (out["status"] == 1).sum() >= 5        # the batch does contain limit cycles
    sel = np.unique(np.concatenate([np.arange(0, len(si), 6), np.nonzero(out["iters"] >= 150)[0]]))
    for i in sel:
        R, t, err, iters, status = icp_oracle.register(scans[i], scans[i + 1], **cfg)
        assert int(out["iters"][i]) == iters and int(out["status"][i]) == status, f"pair {i}"
        assert np.abs(out["t"][i] - t).max() <= 1e-4
        dth = np.arctan2(out["R"][i][1, 0], out["R"][i][0, 0]) - np.arctan2(R[1, 0], R[0, 0])
        assert abs(dth) <= 1e-5


def test_chunked_upload_with_pairs_in_any_order_is_bitwise_identical():
    """icpb200_icp_pairs uploads a large cloud set in four chunks and registers a pair as soon as the chunks of both
    its clouds have landed (pairs grouped by the last chunk they need, an order list on the device).  The results must be
    the ones of the plain batch entry point, bit for bit, in the caller's order -- for pairs that arrive in time order,
    shuffled, reversed, and with initial guesses."""
    scans, poses = synth.make_sequence(1100, world="room", seed=5)        # 18 MB of points: four upload chunks
    flat, off = synth.pack_ragged(scans)
    assert flat.nbytes >= 16 << 20
    rng = np.random.default_rng(11)
    src = rng.integers(0, 1099, size=700).astype(np.int32)
    tgt = np.clip(src + rng.integers(-2, 3, size=700), 0, 1099).astype(np.int32)
    tgt[src == tgt] = src[src == tgt] + 1
    ref = api.icp_batch([scans[i] for i in src], [scans[j] for j in tgt], **CFG)
    assert (ref["status"] == 0).sum() > 500
    for name, perm in (("shuffled", np.arange(700)), ("sorted", np.argsort(np.maximum(src, tgt), kind="stable")),
                       ("reversed", np.argsort(-np.maximum(src, tgt), kind="stable"))):
        out = api.icp_pairs(flat, off, src[perm], tgt[perm], **CFG)
        for key in ("R", "t", "error", "iters", "status"):
            assert out[key].tobytes() == ref[key][perm].tobytes(), (name, key)
    th = rng.normal(0.0, 0.01, size=700)
    R0 = np.stack([[[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]] for a in th])
    t0 = rng.normal(0.0, 0.02, size=(700, 2))
    a = api.icp_batch([scans[i] for i in src], [scans[j] for j in tgt], R_init=R0, t_init=t0, **CFG)
    b = api.icp_pairs(flat, off, src, tgt, R_init=R0, t_init=t0, **CFG)
    for key in ("R", "t", "error", "iters", "status"):
        assert a[key].tobytes() == b[key].tobytes(), ("init", key)


def test_far_away_sources_get_the_exact_nearest_neighbours():
    """A registration that diverges keeps iterating to max_iterations in the reference (icp.py:177-223) with the source
    metres, kilometres or thousands of kilometres from the target.  fp32 cannot rank the candidates out there; the kernel
    then decides many points at once with the lock-step fp64 refine (mid range) or against the target's front set (far
    field).  The first-iteration correspondences must still be scipy's, whatever the direction of the offset."""
    from scipy.spatial import KDTree
    scans, _ = synth.make_sequence(3, world="room", seed=21)
    src0, tgt0 = scans[0], scans[1]
    for dist in (30.0, 400.0, 2.0e4, 3.0e6):
        for ang in (0.0, 0.7, 1.5708, 2.9, 4.0, 5.5):
            shift = dist * np.array([np.cos(ang), np.sin(ang)])
            th = 0.3 * np.sin(ang + dist)
            R0 = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
            trace = {}
            icp_oracle.register(src0, tgt0, 1e-10, 1, 0.04, R_init=R0, t_init=shift, method="point_to_point", trace=trace)
            tr = api.icp_trace(src0, tgt0, 1e-10, 1, 0.04, R_init=R0, t_init=shift, method="point_to_point", trace_iters=1)
            want, got = trace["matches"][0], tr["matches"][0]
            if not np.array_equal(want, got):
                # only exact-distance ties may differ (documented: lowest index here, unspecified in scipy)
                cur = trace["src"] @ R0.T + shift
                bad = np.flatnonzero(want != got)
                d_w = np.sum((cur[bad] - trace["tgt"][want[bad]]) ** 2, axis=1)
                d_g = np.sum((cur[bad] - trace["tgt"][got[bad]]) ** 2, axis=1)
                assert np.array_equal(d_w, d_g), (dist, ang, len(bad), np.abs(d_w - d_g).max())
    # and a whole diverging run stays finite and ends the way the reference's loop does (iteration limit or convergence)
    out = api.icp_batch([src0], [tgt0], 1e-10, 40, 0.04, R_init=np.eye(2)[None], t_init=np.array([[5.0e5, -2.0e5]]),
                        method="point_to_line", normal_k=12)
    assert out["status"][0] in (0, 1) and np.isfinite(out["t"]).all()


def test_3d_point_to_point_batch_against_the_oracle():
    """C1's path in a batch: the teapot cloud under twelve random rigid motions, 3-D point-to-point.
    Same iteration counts and exit status as the oracle, poses within the north-star tolerance -- covers the Kabsch
    solve's warm start (the right singular basis is carried from one iteration to the next, csrc/linalg_small.cuh)."""
    g = load_golden("teapot.npz")
    base = np.ascontiguousarray(g["teapot"])
    rng = np.random.default_rng(11)
    srcs = []
    for _ in range(12):
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        ang = rng.uniform(-0.35, 0.35)
        K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        rot = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        srcs.append(base @ rot.T + rng.uniform(-0.3, 0.3, size=3))
    cfg = dict(error_threshold=1e-10, max_iterations=80, voxel_size=0.01, method="point_to_point")
    out = api.icp_batch(srcs, [base] * len(srcs), **cfg)
    for i, s in enumerate(srcs):
        R, t, err, iters, status = icp_oracle.register(s, base, **cfg)
        assert int(out["iters"][i]) == iters and int(out["status"][i]) == status, f"pair {i}: {out['iters'][i]} vs {iters}"
        dt, dr = pose_delta(out["R"][i], out["t"][i], R, t)
        assert dt < POS_TOL and dr < ROT_TOL, f"pair {i}: dt={dt:.3e} dr={dr:.3e}"
        assert abs(out["error"][i] - err) <= 1e-9 * max(1.0, abs(err))


def test_pair_profile_columns():
    """icpb200_icp_pair_profile: eight counters per pair; the phase cycles (iterations >= 8) stay below the pair's total."""
    scans, _ = synth.make_sequence(40, world="room", seed=4)
    cfg = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, method="point_to_line", normal_k=12)
    flat, off = synth.pack_ragged(scans)
    si = np.arange(len(scans) - 1, dtype=np.int32)
    api.icp_pair_profile(0)                                  # switches the counters on
    out = api.icp_pairs(flat, off, si, si + 1, **cfg)
    prof = api.icp_pair_profile(len(si))
    assert prof.shape == (len(si), 8)
    assert np.array_equal(prof[:, 3], out["iters"])          # iterations
    assert (prof[:, 0] > 0).all() and (prof[:, 1] > 0).all()
    assert (prof[:, 4:7].sum(axis=1) <= prof[:, 0]).all() and (prof[:, 7] == 0).all()
    long_ones = out["iters"] > 8
    assert (prof[long_ones, 4:7].sum(axis=1) > 0).all()
