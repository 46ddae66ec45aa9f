"""CPU oracle of the reference's 2-D pose graph (test infrastructure: only tests/, smoke() and bench.py's CPU legs may
import it -- the product path is icpb200_pose_graph_optimize in libicp_b200.so).

A plain numpy restatement of utilities/pose_graph.py of the reference, dense normal matrix and np.linalg.solve included
(pose_graph.py:83-134), so that it reproduces the reference's numbers; pinned against the live reference by
tests/golden/pose_graph.npz (oracle/make_pose_graph_golden.py)."""
import numpy as np


def wrap(a):
    """pose_graph.py:15-17."""
    return (a + np.pi) % (2 * np.pi) - np.pi


def edge_terms(xi, xj, z):
    """Error and Jacobians of one edge (pose_graph.py:138-178)."""
    c, s = np.cos(xi[2]), np.sin(xi[2])
    rt = np.array([[c, s], [-s, c]])
    dt = xj[:2] - xi[:2]
    dth = wrap(xj[2] - xi[2])
    pred = rt @ dt
    e = np.array([pred[0] - z[0], pred[1] - z[1], wrap(dth - z[2])])
    drt = np.array([[-s, c], [-c, -s]]) @ dt
    a = np.zeros((3, 3))
    a[:2, :2] = -rt
    a[:2, 2] = drt
    a[2, 2] = -1.0
    b = np.zeros((3, 3))
    b[:2, :2] = rt
    b[2, 2] = 1.0
    return e, a, b


def optimize(nodes, edges, n_iterations=20, fix_node=0, convergence_eps=1e-6):
    """nodes: (n, 3) array, updated copy returned; edges: list of (i, j, z (3,), omega (3, 3)).
    Returns (nodes, iteration index the loop ended on, last step norm, status 0 converged / 1 limit / 4 singular)."""
    nodes = np.array(nodes, dtype=float)
    n = len(nodes)
    if n < 2 or len(edges) == 0:
        return nodes, 0, 0.0, 0
    step = 0.0
    for it in range(n_iterations):
        h = np.zeros((3 * n, 3 * n))
        g = np.zeros(3 * n)
        for i, j, z, om in edges:
            e, a, b = edge_terms(nodes[i], nodes[j], z)
            si, sj = 3 * i, 3 * j
            h[si:si + 3, si:si + 3] += a.T @ om @ a          # pose_graph.py:99-105
            h[si:si + 3, sj:sj + 3] += a.T @ om @ b
            h[sj:sj + 3, si:si + 3] += b.T @ om @ a
            h[sj:sj + 3, sj:sj + 3] += b.T @ om @ b
            g[si:si + 3] += a.T @ om @ e
            g[sj:sj + 3] += b.T @ om @ e
        sf = 3 * fix_node                                    # pose_graph.py:107-112
        h[sf:sf + 3, :] = 0
        h[:, sf:sf + 3] = 0
        h[sf:sf + 3, sf:sf + 3] = np.eye(3) * 1e10
        g[sf:sf + 3] = 0
        try:
            dx = np.linalg.solve(h, -g)
        except np.linalg.LinAlgError:
            return nodes, it, step, 4
        nodes[:, 0] += dx[0::3]                              # pose_graph.py:121-125
        nodes[:, 1] += dx[1::3]
        nodes[:, 2] = wrap(nodes[:, 2] + dx[2::3])
        step = float(np.linalg.norm(dx))
        if step < convergence_eps:
            return nodes, it, step, 0
    return nodes, n_iterations, step, 1
