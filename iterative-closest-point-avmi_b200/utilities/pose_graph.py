"""``utilities.pose_graph`` with the solver in libicp_b200.so -- same call surface as the reference.

``PoseGraph2D`` keeps the reference's attributes (``nodes``: list of [x, y, theta] arrays, ``edges``: list of
(i, j, z, omega)), methods, defaults and the one console line per ``optimize`` call
(/root/reference/utilities/pose_graph.py:41-134, 182-194); the helpers keep their names and conventions
(pose_graph.py:15-37).  ``optimize`` hands nodes and edges to ``icpb200_pose_graph_optimize``: Gauss-Newton with a
block-skyline Cholesky on the host instead of a dense 3n x 3n ``np.linalg.solve`` per iteration -- the same poses to
rounding, at a cost that grows with the nodes and the loop lengths instead of n^3.  No numpy fallback: without the
library ``optimize`` raises."""
import ctypes

import numpy as np

from icp_b200 import _lib as _abi


def normalize_angle(a):
    """Wrap an angle to [-pi, pi) the way the reference does (pose_graph.py:15-17)."""
    return (a + np.pi) % (2 * np.pi) - np.pi


def pose_matrix_to_vec(T):
    """3 x 3 homogeneous matrix -> [x, y, theta] (pose_graph.py:20-22)."""
    return np.array([T[0, 2], T[1, 2], np.arctan2(T[1, 0], T[0, 0])])


def pose_vec_to_matrix(v):
    """[x, y, theta] -> 3 x 3 homogeneous matrix (pose_graph.py:25-31)."""
    c, s = np.cos(v[2]), np.sin(v[2])
    return np.array([[c, -s, v[0]], [s, c, v[1]], [0, 0, 1]])


def relative_transform_vec(T_i, T_j):
    """z_ij = T_i^-1 T_j as [dx, dy, dtheta] (pose_graph.py:34-37)."""
    return pose_matrix_to_vec(np.linalg.inv(T_i) @ T_j)


class PoseGraph2D:
    """2-D pose graph, Gauss-Newton on SE(2) (pose_graph.py:41-55 for the usage)."""

    def __init__(self):
        self.nodes = []
        self.edges = []

    def add_node(self, pose_vec):
        """Append a pose [x, y, theta] (copied); returns its index (pose_graph.py:63-66)."""
        index = len(self.nodes)
        self.nodes.append(np.array(pose_vec, dtype=np.float64))
        return index

    def add_edge(self, i, j, measurement, information=None):
        """Constraint i -> j: relative pose [dx, dy, dtheta] in frame i and its 3 x 3 information matrix, identity when
        omitted (pose_graph.py:68-79).  Both are copied."""
        weight = np.identity(3) if information is None else np.array(information, dtype=np.float64)
        self.edges.append((i, j, np.array(measurement, dtype=np.float64), weight))

    def _pack(self):
        poses = np.ascontiguousarray(np.array(self.nodes, dtype=np.float64).reshape(-1, 3))
        ei = np.array([e[0] for e in self.edges], dtype=np.int32)
        ej = np.array([e[1] for e in self.edges], dtype=np.int32)
        z = np.ascontiguousarray(np.array([e[2] for e in self.edges], dtype=np.float64).reshape(-1, 3))
        om = np.ascontiguousarray(np.array([e[3] for e in self.edges], dtype=np.float64).reshape(-1, 9))
        return poses, ei, ej, z, om

    def optimize(self, n_iterations=20, fix_node=0, convergence_eps=1e-6):
        """Minimise the total edge error with the pose ``fix_node`` held (pose_graph.py:83-134)."""
        n = len(self.nodes)
        if n < 2 or len(self.edges) == 0:
            return
        poses, ei, ej, z, om = self._pack()
        iters, status, step = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_double(0.0)
        dp, ip = _abi.c_double_p, _abi.c_int32_p
        rc = _abi.load().icpb200_pose_graph_optimize(
            n, poses.ctypes.data_as(dp), len(self.edges), ei.ctypes.data_as(ip), ej.ctypes.data_as(ip),
            z.ctypes.data_as(dp), om.ctypes.data_as(dp), int(n_iterations), int(fix_node), float(convergence_eps),
            ctypes.byref(iters), ctypes.byref(step), ctypes.byref(status))
        _abi.check(rc, "icpb200_pose_graph_optimize")
        for k in range(n):
            self.nodes[k][:] = poses[k]
        if status.value == _abi.SINGULAR:
            print(f"  PoseGraph: singular H at iter {iters.value}, stopping")
        elif status.value == _abi.CONVERGED:
            print(f"  PoseGraph converged: iter={iters.value}, ||Δx||={step.value:.2e}")
        else:
            print(f"  PoseGraph max iterations: iter={n_iterations}, ||Δx||={step.value:.2e}")

    def get_poses_as_matrices(self):
        return [pose_vec_to_matrix(v) for v in self.nodes]

    def total_error(self):
        """Sum over the edges of e^T Omega e (pose_graph.py:186-192)."""
        total = 0.0
        for i, j, z, omega in self.edges:
            xi, xj = self.nodes[i], self.nodes[j]
            c, s = np.cos(xi[2]), np.sin(xi[2])
            d = xj[:2] - xi[:2]
            e = np.array([c * d[0] + s * d[1] - z[0], -s * d[0] + c * d[1] - z[1],
                          normalize_angle(normalize_angle(xj[2] - xi[2]) - z[2])])
            total += e @ omega @ e
        return total
