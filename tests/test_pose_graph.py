"""the pose-graph optimiser (utilities/pose_graph.py of the reference).

No GPU needed: the solver is the host side of libicp_b200.so (iterative-closest-point-avmi_b200/host/pose_graph.cpp).
The oracle (oracle/pose_graph_oracle.py, a numpy restatement with the reference's dense solve) is pinned bit for bit
against the live reference through tests/golden/pose_graph.npz (oracle/make_pose_graph_golden.py); the library is
checked against the golden poses, against the oracle on a larger seeded graph, and on the reference's edge cases."""
import contextlib
import io
import time

import numpy as np

from conftest import load_golden
from oracle import pose_graph_oracle


def _graph(g, name):
    est = g[f"{name}_nodes_in"]
    edges = [(int(a), int(b), z, om) for (a, b), z, om in zip(g[f"{name}_edge_ij"], g[f"{name}_edge_z"], g[f"{name}_edge_info"])]
    return est, edges, eval(str(g[f"{name}_kw"]))


def _optimize(est, edges, **kw):
    from utilities.pose_graph import PoseGraph2D
    pg = PoseGraph2D()
    for p in est:
        pg.add_node(p)
    for e in edges:
        pg.add_edge(*e)
    log = io.StringIO()
    with contextlib.redirect_stdout(log):
        pg.optimize(**kw)
    return pg, log.getvalue().strip()


def _pose_diff(a, b):
    d = np.asarray(a) - np.asarray(b)
    d[:, 2] = (d[:, 2] + np.pi) % (2 * np.pi) - np.pi
    return np.abs(d).max()


def test_oracle_reproduces_the_reference_bit_for_bit():
    g = load_golden("pose_graph.npz")
    for name in ("small", "loops", "anchor"):
        est, edges, kw = _graph(g, name)
        nodes, it, step, status = pose_graph_oracle.optimize(est, edges, **kw)
        assert nodes.tobytes() == g[f"{name}_nodes_out"].tobytes(), name
        assert status == 0 and f"iter={it}," in str(g[f"{name}_line"])


def test_library_matches_the_reference_run():
    g = load_golden("pose_graph.npz")
    for name in ("small", "loops", "anchor"):
        est, edges, kw = _graph(g, name)
        pg, line = _optimize(est, edges, **kw)
        assert _pose_diff(pg.nodes, g[f"{name}_nodes_out"]) < 1e-10, name        # observed <= 7e-15
        assert line == str(g[f"{name}_line"]), name                              # the reference's console line
        assert abs(pg.total_error() - float(g[f"{name}_total_error"])) < 1e-9
        assert len(pg.get_poses_as_matrices()) == len(est)


def test_library_matches_the_oracle_on_a_larger_graph_and_is_faster():
    rng = np.random.default_rng(5)
    n = 500
    truth = np.cumsum(np.column_stack([0.25 * np.cos(np.linspace(0, 4 * np.pi, n)), 0.25 * np.sin(np.linspace(0, 4 * np.pi, n)),
                                       np.full(n, 4 * np.pi / n)]), axis=0)
    truth[:, 2] = pose_graph_oracle.wrap(truth[:, 2])

    def rel(a, b):
        c, s = np.cos(a[2]), np.sin(a[2])
        d = b[:2] - a[:2]
        return np.array([c * d[0] + s * d[1], -s * d[0] + c * d[1], pose_graph_oracle.wrap(b[2] - a[2])])

    est = truth + rng.normal(0, [0.05, 0.05, 0.01], size=truth.shape)
    est[0] = truth[0]
    edges = [(k - 1, k, rel(truth[k - 1], truth[k]) + rng.normal(0, 0.002, 3), np.diag([150.0, 150.0, 500.0])) for k in range(1, n)]
    for _ in range(80):
        a, b = rng.choice(n, size=2, replace=False)
        edges.append((int(a), int(b), rel(truth[a], truth[b]) + rng.normal(0, 0.002, 3), np.diag([300.0, 300.0, 900.0])))
    t0 = time.perf_counter()
    want, it, step, status = pose_graph_oracle.optimize(est, edges, n_iterations=25)
    t_oracle = time.perf_counter() - t0
    t0 = time.perf_counter()
    pg, line = _optimize(est, edges, n_iterations=25)
    t_lib = time.perf_counter() - t0
    assert status == 0 and f"converged: iter={it}," in line
    assert _pose_diff(pg.nodes, want) < 1e-9
    assert t_lib < t_oracle                    # (dense 1500 x 1500 solve per iteration against a skyline of 500 block rows)


def test_edge_cases():
    from utilities.pose_graph import PoseGraph2D, normalize_angle, pose_matrix_to_vec, pose_vec_to_matrix, relative_transform_vec
    pg = PoseGraph2D()
    pg.optimize()                                                   # pose_graph.py:90-92: nothing to do, no output
    pg.add_node([0, 0, 0])
    pg.add_node([1, 0, 0.1])
    pg.optimize()
    assert np.array_equal(pg.nodes[1], [1, 0, 0.1])
    # an unconstrained node makes the normal matrix singular: the reference stops with a message (pose_graph.py:115-119)
    pg.add_node([2, 0, 0.0])
    pg.add_edge(0, 1, [1.0, 0.0, 0.1])
    log = io.StringIO()
    with contextlib.redirect_stdout(log):
        pg.optimize()
    assert "singular H at iter 0" in log.getvalue()
    # iteration limit line (pose_graph.py:131-134)
    g = load_golden("pose_graph.npz")
    est, edges, _ = _graph(g, "loops")
    _, line = _optimize(est, edges, n_iterations=1)
    assert line.startswith("PoseGraph max iterations: iter=1,")
    # an edge to a node that does not exist: the reference raises IndexError, this raises with the library's message
    bad = PoseGraph2D()
    bad.add_node([0, 0, 0]); bad.add_node([1, 0, 0])
    bad.add_edge(0, 5, [1.0, 0.0, 0.0])
    import pytest
    with pytest.raises(RuntimeError, match="edge 0 joins 0 and 5"):
        bad.optimize()
    # helpers (pose_graph.py:15-37)
    v = np.array([0.3, -1.2, 2.9])
    assert np.allclose(pose_matrix_to_vec(pose_vec_to_matrix(v)), v)
    assert abs(normalize_angle(3 * np.pi + 0.1) - (-np.pi + 0.1)) < 1e-12
    t1, t2 = pose_vec_to_matrix([1.0, 2.0, 0.5]), pose_vec_to_matrix([1.5, 2.5, 0.9])
    z = relative_transform_vec(t1, t2)
    assert np.allclose(pose_vec_to_matrix([1.0, 2.0, 0.5]) @ pose_vec_to_matrix(z), t2)
