"""Device helper headers (bres.cuh, linalg_small.cuh) compiled for the host and
checked against the oracle / numpy.  These are the exact functions the CUDA
kernels call; a GPU is not needed to prove their arithmetic."""
import ctypes

import numpy as np

from oracle import occupancy_oracle as oo

DP = ctypes.POINTER(ctypes.c_double)


def P(a):
    return a.ctypes.data_as(DP)


def ray_cells(h, ts, ox, oy, hx, hy, nx, ny):
    cap = max(abs(hx - ox), abs(hy - oy)) + 1
    buf = np.empty((cap, 3), np.int32)
    n = h.harness_ray_cells(ts, ox, oy, hx, hy, nx, ny, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), cap)
    assert 0 <= n <= cap
    return buf[:n]


def test_closed_form_bresenham_exhaustive(harness):
    """Every endpoint in a 43x43 window, several origins (inside, outside, far
    outside the grid), two tile sizes: cells, order and tile ids all match the
    reference walk clipped to the grid."""
    for ts in (4, 8):
        for ox, oy, nx, ny in ((5, 6, 24, 17), (-3, 2, 24, 17), (30, 30, 24, 17), (0, 0, 9, 9)):
            tiles_x = (nx + ts - 1) // ts
            for hx in range(ox - 21, ox + 22):
                for hy in range(oy - 21, oy + 22):
                    ref = [c for c in oo.line_cells_py(ox, oy, hx, hy) if 0 <= c[0] < nx and 0 <= c[1] < ny]
                    got = ray_cells(harness, ts, ox, oy, hx, hy, nx, ny)
                    assert len(ref) == len(got), (ts, ox, oy, hx, hy)
                    if ref:
                        assert np.array_equal(np.asarray(ref), got[:, :2]), (ts, ox, oy, hx, hy)
                        assert np.array_equal(got[:, 2], (got[:, 1] // ts) * tiles_x + got[:, 0] // ts)


def test_closed_form_bresenham_long_rays(harness):
    rng = np.random.default_rng(0)
    for _ in range(1500):
        ox, oy = (int(v) for v in rng.integers(-500, 4600, 2))
        hx, hy = (int(v) for v in rng.integers(-3000, 7000, 2))
        ref = oo.line_cells_c(ox, oy, hx, hy)
        ref = ref[(ref[:, 0] >= 0) & (ref[:, 0] < 4096) & (ref[:, 1] >= 0) & (ref[:, 1] < 4096)]
        got = ray_cells(harness, 64, ox, oy, hx, hy, 4096, 4096)
        assert len(ref) == len(got) and np.array_equal(ref, got[:, :2])


def test_run_iterator_equals_callback_form(harness):
    rng = np.random.default_rng(9)
    a, b = np.empty((256, 4), np.int32), np.empty((256, 4), np.int32)
    pa, pb = (x.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) for x in (a, b))
    for _ in range(2000):
        ox, oy = (int(v) for v in rng.integers(-300, 4400, 2))
        hx, hy = (int(v) for v in rng.integers(-2000, 6000, 2))
        na = harness.harness_ray_runs_iter(ox, oy, hx, hy, 4096, 4096, pa, 256)
        nb = harness.harness_ray_runs_cb(ox, oy, hx, hy, 4096, 4096, pb, 256)
        assert na == nb and np.array_equal(a[:na], b[:nb])
        if na:
            assert a[:na, 3].min() >= 1 and a[:na, 3].max() <= 64


def test_minor_steps_matches_walk(harness):
    rng = np.random.default_rng(3)
    for _ in range(300):
        ox, oy, hx, hy = (int(v) for v in rng.integers(-40, 40, 4))
        cells = oo.line_cells_py(ox, oy, hx, hy)
        xmajor = abs(hx - ox) >= abs(hy - oy)
        for n, (x, y) in enumerate(cells):
            j = abs(y - oy) if xmajor else abs(x - ox)
            assert harness.harness_minor_steps(ox, oy, hx, hy, n) == j


def test_sat_cell(harness):
    assert harness.harness_sat_cell(12.0) == 12 and harness.harness_sat_cell(-7.0) == -7
    assert harness.harness_sat_cell(1e300) == 1 << 27 and harness.harness_sat_cell(-1e300) == -(1 << 27)
    assert harness.harness_sat_cell(float("nan")) == -(1 << 27)


def test_solve3_matches_lapack(harness):
    rng = np.random.default_rng(1)
    for _ in range(500):
        a = rng.normal(size=(200, 3)) * [10, 1, 1]
        m = np.ascontiguousarray(a.T @ a)
        b = rng.normal(size=3)
        x = np.empty(3)
        assert harness.harness_solve3(P(m), P(b), P(x)) == 0
        ref = np.linalg.solve(m, b)
        assert np.abs(x - ref).max() <= 1e-13 * np.abs(ref).max()
    x = np.empty(3)
    # exactly singular -> 1 (numpy raises LinAlgError; the reference then takes the identity step)
    assert harness.harness_solve3(P(np.zeros((3, 3))), P(np.ones(3)), P(x)) == 1
    assert harness.harness_solve3(P(np.array([[1., 2, 3], [2, 4, 6], [1, 1, 1]])), P(np.ones(3)), P(x)) == 1


def _ref_kabsch(w):
    u, _, vt = np.linalg.svd(w)
    r = vt.T @ u.T
    if np.linalg.det(r) < 0:
        vt[-1, :] *= -1
        r = vt.T @ u.T
    return r


def test_kabsch_matches_svd_path(harness):
    rng = np.random.default_rng(2)
    for k in range(1500):
        w = rng.normal(size=(3, 3))
        if k % 5 == 0:
            w = w @ np.diag([1, 1, -1])          # reflection: exercises the det < 0 fix
        if k % 7 == 0:
            w[:, 2] = 0                          # rank 2 (planar data)
        if k % 11 == 0:
            w *= 1e-9
        r = np.empty((3, 3))
        harness.harness_kabsch3(P(np.ascontiguousarray(w)), P(r))
        assert np.abs(r - _ref_kabsch(w.copy())).max() < 1e-12
        assert abs(np.linalg.det(r) - 1.0) < 1e-12
        w2 = rng.normal(size=(2, 2)) @ (np.diag([1, -1]) if k % 3 == 0 else np.eye(2))
        r2 = np.empty((2, 2))
        harness.harness_kabsch2(P(np.ascontiguousarray(w2)), P(r2))
        assert np.abs(r2 - _ref_kabsch(w2.copy())).max() < 1e-13


def test_eigvec2_matches_eigh(harness):
    rng = np.random.default_rng(4)
    n = np.empty(2)
    for k in range(2000):
        pts = rng.normal(size=(13, 2)) * rng.uniform(0.01, 3, size=2)
        if k % 4 == 0:
            pts[:, 1] = 0.3 * pts[:, 0] + 1e-3 * rng.normal(size=13)
        c = np.cov(pts.T)
        ref = np.linalg.eigh(c)[1][:, 0]
        harness.harness_eigvec2(c[0, 0], c[0, 1], c[1, 1], P(n))
        assert min(np.abs(n - ref).max(), np.abs(n + ref).max()) < 1e-12
    for m, want in (([[0, 0], [0, 0]], [1, 0]), ([[2, 0], [0, 2]], [1, 0]), ([[1, 0], [0, 3]], [1, 0]),
                    ([[3, 0], [0, 1]], [0, 1])):
        m = np.asarray(m, float)
        harness.harness_eigvec2(m[0, 0], m[0, 1], m[1, 1], P(n))
        assert n.tolist() == want                # degenerate cases as eigh returns them
