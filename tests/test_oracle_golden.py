"""The oracle restatement reproduces every golden fixture bit-for-bit.

The fixtures were written by oracle/pin_against_reference.py from the live
reference (run in the build container); this test keeps the oracle pinned to
them wherever the suite runs (no /root/reference needed).
"""
import numpy as np

from conftest import case_kwargs, icp_cases, load_golden
from oracle import icp_oracle, occupancy_oracle


def same_bits(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.tobytes() == b.tobytes()


def test_teapot_fixture_and_known_answer():
    g = load_golden("teapot.npz")
    R, t, err, iters, status = icp_oracle.register(
        g["moved"], g["teapot"], error_threshold=1e-12, max_iterations=300, voxel_size=0.005,
        method="point_to_point")
    assert same_bits(R, g["R"]) and same_bits(t, g["t"]) and np.float64(err).tobytes() == g["err"].tobytes()
    assert iters == int(g["iters"]) and status == int(g["status"]) == icp_oracle.STATUS_CONVERGED
    # analytic answer: inverse of x -> Ry x + shift (demos/teapot_icp_demo.py:38-47)
    assert np.abs(R - g["ry"].T).max() < 1e-12
    assert np.abs(t + g["ry"].T @ g["shift"]).max() < 1e-12


def test_voxel_fixtures():
    g = load_golden("voxel.npz")
    tags = sorted({k[:-3] for k in g.files if k.endswith("_in")})
    assert len(tags) == 5
    for tag in tags:
        out = icp_oracle.voxel_means(g[f"{tag}_in"], float(g[f"{tag}_v"]))
        assert same_bits(out, g[f"{tag}_out"]), tag


def test_normals_fixtures():
    g = load_golden("normals.npz")
    for tag in ("scan0_k12", "scan1_k10", "tiny_k12"):
        out = icp_oracle.pca_normals_2d(g[f"{tag}_in"], k=int(g[f"{tag}_k"]))
        assert same_bits(out, g[f"{tag}_out"]), tag
        assert np.allclose(np.linalg.norm(out, axis=1), 1.0)


def test_icp2d_fixtures():
    cases = icp_cases(load_golden("icp2d.npz"))
    assert len(cases) == 12
    for name, c in cases.items():
        R, t, err, iters, status = icp_oracle.register(c["src"], c["tgt"], **case_kwargs(c))
        assert same_bits(R, c["R"]) and same_bits(t, c["t"]), name
        assert np.float64(err).tobytes() == c["err"].tobytes(), name
        assert iters == int(c["iters"]) and status == int(c["status"]), name
    assert int(cases["gate_break"]["status"]) == icp_oracle.STATUS_FEW_INLIERS
    assert np.isinf(cases["gate_break"]["err"])
    assert int(cases["few_iters"]["status"]) == icp_oracle.STATUS_MAX_ITER


def test_bresenham_fixture():
    g = load_golden("bresenham.npz")
    ends, cells, off = g["ends"], g["cells"], g["off"]
    for i, (x0, y0, x1, y1) in enumerate(ends):
        want = cells[off[i]:off[i + 1]]
        got_py = occupancy_oracle.line_cells_py(int(x0), int(y0), int(x1), int(y1))
        got_c = occupancy_oracle.line_cells_c(int(x0), int(y0), int(x1), int(y1))
        assert len(want) == max(abs(x1 - x0), abs(y1 - y0))
        assert np.array_equal(np.asarray(got_py, dtype=np.int32).reshape(-1, 2), want)
        assert np.array_equal(got_c.astype(np.int32), want)


def _replay(g, cls, n, **kw):
    b = g["bounds"]
    grid = cls(b[0], b[1], b[2], b[3], resolution=float(g["resolution"]), p_hit=float(g["p_hit"]),
               p_miss=float(g["p_miss"]), log_odds_min=kw.get("lo", float(g["log_odds_min"])),
               log_odds_max=kw.get("hi", float(g["log_odds_max"])))
    off = g["hit_off"]
    for s in range(n):
        grid.update_scan(g["origins"][s], g["hits"][off[s]:off[s + 1]])
    return grid


def test_occupancy_fixture_c_oracle():
    g = load_golden("occupancy.npz")
    grid = _replay(g, occupancy_oracle.GridOracleC, 1)
    assert same_bits(grid.log_odds, g["snap_0"])
    grid = _replay(g, occupancy_oracle.GridOracleC, 14)
    assert same_bits(grid.log_odds, g["snap_13"])
    grid = _replay(g, occupancy_oracle.GridOracleC, 40)
    assert same_bits(grid.log_odds, g["snap_39"])
    assert grid.l_hit == 1.7346010553881064 and grid.l_miss == -0.3227733922630512
    odd = _replay(g, occupancy_oracle.GridOracleC, 3, lo=0.5, hi=3.0)
    assert same_bits(odd.log_odds, g["odd_final"])


def test_occupancy_fixture_py_oracle_and_batch():
    g = load_golden("occupancy.npz")
    grid = _replay(g, occupancy_oracle.GridOraclePy, 1)
    assert same_bits(grid.log_odds, g["snap_0"])
    b = g["bounds"]
    many = occupancy_oracle.GridOracleC(b[0], b[1], b[2], b[3], resolution=0.05, p_hit=0.85, p_miss=0.42,
                                        log_odds_min=-8.0, log_odds_max=8.0)
    cells = many.update_many(g["origins"], g["hits"], g["hit_off"], fast=True)
    assert same_bits(many.log_odds, g["snap_39"]) and cells > 0


def test_submap_fixture():
    g = load_golden("submap.npz")
    assert same_bits(icp_oracle.voxel_means(g["raw_sub"], 0.04), g["submap"])
    for tag in ("a", "b"):
        R, t, err, iters, status = icp_oracle.register(
            g[f"{tag}/src"], g["submap"], 1e-10, 150, 0.04, R_init=g[f"{tag}/R_init"], t_init=g[f"{tag}/t_init"],
            method="point_to_point", max_corr_dist=1.5)
        assert same_bits(R, g[f"{tag}/R"]) and same_bits(t, g[f"{tag}/t"]) and iters == int(g[f"{tag}/iters"])


def test_rotation_search_oracle_reproduces_the_reference():
    """tests/golden/rotation.npz holds the live reference's outputs (oracle/pin_rotation.py)."""
    from oracle import features_oracle as fo
    g = load_golden("rotation.npz")
    for name in ("cfg", "defaults", "turned", "tiny"):
        v, c, f = g[f"rs_{name}_kw"]
        R, t, score = fo.rotation_search(g[f"rs_{name}_src"], g[f"rs_{name}_tgt"], voxel_size=v, angle_step_coarse=c,
                                         angle_step_fine=f)[:3]
        assert R.tobytes() == g[f"rs_{name}_R"].tobytes() and t.tobytes() == g[f"rs_{name}_t"].tobytes()
        assert np.float64(score).tobytes() == g[f"rs_{name}_score"].tobytes()
    for name in ("cfg", "defaults"):
        ar, st, fs, v = g[f"sub_{name}_kw"]
        R, t = fo.submap_rotation_search(g[f"sub_{name}_src"], g[f"sub_{name}_map"], g[f"sub_{name}_pose"], angle_range=ar,
                                         angle_step=st, fine_step=fs, voxel_size=v)
        assert R.tobytes() == g[f"sub_{name}_R"].tobytes() and t.tobytes() == g[f"sub_{name}_t"].tobytes()


def _rebuild_history(g, variant):
    off = g["scan_off"]
    return [(g["scans"][off[s]:off[s + 1]], g[f"poses_{variant}"][s]) for s in range(len(off) - 1)]


def test_rebuild_map_oracle_reproduces_the_reference():
    """slam.py:271-277 `_rebuild_map` + slam.py:46-50 `transform_points_2d`: the reference's own maps for a 14-scan
    history (an empty scan, endpoints beyond the grid), before and after its poses were corrected
    (tests/golden/rebuild.npz, written by oracle/pin_rebuild.py from the live reference)."""
    g = load_golden("rebuild.npz")
    kw = dict(resolution=0.05, p_hit=0.85, p_miss=0.42, log_odds_min=-8.0, log_odds_max=8.0)
    grid = occupancy_oracle.GridOracleC(*g["bounds"], **kw)
    assert grid.log_odds.shape == tuple(g["grid_shape"])
    for variant in (0, 1, 0):                      # a rebuild starts from a cleared grid whatever was in it
        occupancy_oracle.rebuild_map(grid, _rebuild_history(g, variant))
        want = np.zeros(grid.log_odds.size, dtype=np.float32)
        want[g[f"nz_index_{variant}"]] = g[f"nz_value_{variant}"]
        assert same_bits(grid.log_odds.ravel(), want), variant
