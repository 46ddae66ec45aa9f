#!/usr/bin/env python
"""Golden vectors of the reference's PoseGraph2D (run in the build container: imports /root/reference).

Three graphs: a short chain with one loop closure; a 300-node noisy trajectory with 25 loop closures (some pointing
backwards, as slam.py:593 adds them: cur -> candidate) and non-identity information matrices; a graph anchored at a node
other than 0.  For each: the inputs, the reference's optimised poses, and the line it printed."""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ICP_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REF)
import types  # noqa: E402
sys.modules.setdefault("pyvista", types.ModuleType("pyvista"))      # utilities/mapping.py:2 imports it; display only
from utilities.pose_graph import PoseGraph2D  # noqa: E402  (the reference)


def rel(pi, pj):
    c, s = np.cos(pi[2]), np.sin(pi[2])
    d = pj[:2] - pi[:2]
    return np.array([c * d[0] + s * d[1], -s * d[0] + c * d[1], (pj[2] - pi[2] + np.pi) % (2 * np.pi) - np.pi])


def make_graph(n, n_loops, seed, info_scale=1.0):
    rng = np.random.default_rng(seed)
    truth = np.zeros((n, 3))
    for k in range(1, n):                         # a closed rounded path
        th = truth[k - 1, 2] + 2 * np.pi / n + rng.normal(0, 0.002)
        truth[k] = [truth[k - 1, 0] + 0.3 * np.cos(th), truth[k - 1, 1] + 0.3 * np.sin(th), th]
    truth[:, 2] = (truth[:, 2] + np.pi) % (2 * np.pi) - np.pi
    est = np.zeros_like(truth)
    edges = []
    for k in range(1, n):                         # odometry with drift
        z = rel(truth[k - 1], truth[k]) + rng.normal(0, [0.01, 0.01, 0.004])
        c, s = np.cos(est[k - 1, 2]), np.sin(est[k - 1, 2])
        est[k] = [est[k - 1, 0] + c * z[0] - s * z[1], est[k - 1, 1] + s * z[0] + c * z[1], est[k - 1, 2] + z[2]]
        edges.append((k - 1, k, z, np.diag([100.0, 100.0, 400.0]) * info_scale))
    est[:, 2] = (est[:, 2] + np.pi) % (2 * np.pi) - np.pi
    for _ in range(n_loops):
        a, b = sorted(rng.choice(n, size=2, replace=False))
        if b - a < 5:
            continue
        i, j = (b, a) if rng.random() < 0.7 else (a, b)          # slam.py:593: (cur_idx, cand_idx), cur > cand
        z = rel(truth[i], truth[j]) + rng.normal(0, [0.005, 0.005, 0.002])
        m = rng.normal(size=(3, 3)) * 0.1
        edges.append((i, j, z, (np.diag([200.0, 200.0, 800.0]) + m @ m.T) * info_scale))
    return est, edges


def run(est, edges, **kw):
    pg = PoseGraph2D()
    for p in est:
        pg.add_node(p)
    for i, j, z, om in edges:
        pg.add_edge(i, j, z, om)
    log = io.StringIO()
    with contextlib.redirect_stdout(log):
        pg.optimize(**kw)
    return np.array(pg.nodes), log.getvalue().strip(), pg.total_error()


out = {}
for name, (n, loops, seed, kw) in {"small": (12, 1, 1, {}), "loops": (300, 25, 2, dict(n_iterations=30)),
                                   "anchor": (40, 4, 3, dict(fix_node=17, n_iterations=25, convergence_eps=1e-9))}.items():
    est, edges = make_graph(n, max(loops, 1), seed)
    nodes, line, total = run(est, edges, **kw)
    out[f"{name}_nodes_in"] = est
    out[f"{name}_edge_ij"] = np.array([[e[0], e[1]] for e in edges], dtype=np.int32)
    out[f"{name}_edge_z"] = np.array([e[2] for e in edges])
    out[f"{name}_edge_info"] = np.array([e[3] for e in edges])
    out[f"{name}_nodes_out"] = nodes
    out[f"{name}_line"] = np.array(line)
    out[f"{name}_total_error"] = np.array(total)
    out[f"{name}_kw"] = np.array(repr(kw))
    print(name, len(est), "nodes", len(edges), "edges:", line, "total error", total)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pose_graph.npz"), **out)
