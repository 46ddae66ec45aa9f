"""CPU oracle for the ICP registration path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy restatement of the reference's registration loop.
It exists so the CUDA path can be checked against the reference's arithmetic
on the GPU box, where /root/reference is not present.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it; the product (``libicp_b200.so`` and the
``utilities`` shim) never does and has no CPU fallback.

Pinning: ``oracle/pin_against_reference.py`` runs this file and the live
reference (imported from /root/reference) on the same seeded inputs and
requires bit-identical R, t, error on every case before it writes the golden
fixtures under ``tests/golden/``.  The reference itself ships no tests or
golden vectors ("parity unpinned" by the reference); the
pins therefore come from the live reference run in the build container.

Third-party arithmetic reached from here (not under /root/reference):
scipy.spatial.KDTree (reference pins scipy==1.17.0, requirements.txt:2) and
numpy/LAPACK solve/svd/eigh (numpy==2.3.5, requirements.txt:1).  The same
calls are made in the same order as the reference so results are bit-equal.

Every function cites the reference lines it restates (paths relative to
/root/reference).
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import KDTree

STATUS_CONVERGED = 0      # utilities/icp.py:217-219
STATUS_MAX_ITER = 1       # utilities/icp.py:222-223 (loop exhausted)
STATUS_FEW_INLIERS = 2    # utilities/icp.py:186-187 (break)


def voxel_means(cloud, voxel):
    """Voxel-grid mean downsample.  Restates utilities/icp.py:117-129.

    Rows come out in lexicographic order of the integer voxel index (the order
    ``np.unique(axis=0)`` produces); each mean is an input-order sum divided by
    the member count.
    """
    lo = np.min(cloud, axis=0)
    cell = np.floor((cloud - lo) / voxel).astype(int)
    uniq, member_of = np.unique(cell, axis=0, return_inverse=True)
    member_of = np.asarray(member_of).reshape(-1)
    n_out = len(uniq)
    weight = np.bincount(member_of, minlength=n_out).astype(np.float64)
    out = np.empty((n_out, cloud.shape[1]))
    for axis in range(cloud.shape[1]):
        out[:, axis] = np.bincount(member_of, weights=cloud[:, axis], minlength=n_out)
    out /= weight[:, np.newaxis]
    return out


def pca_normals_2d(cloud, k=10, return_neighbours=False):
    """Unit normals from PCA of the (k+1)-NN set (self included).

    Restates utilities/icp.py:51-76: k is capped at n-1, the covariance is
    ``np.cov`` (ddof=1), the normal is the eigenvector of the smallest
    eigenvalue from ``eigh``, divided by max(norm, 1e-10).
    """
    n = len(cloud)
    k = min(k, n - 1)
    _, nbr = KDTree(cloud).query(cloud, k=k + 1)
    normals = np.zeros_like(cloud)
    for i in range(n):
        cov = np.cov(cloud[nbr[i]].T)
        _, vec = np.linalg.eigh(cov)
        normals[i] = vec[:, 0]
    length = np.linalg.norm(normals, axis=1, keepdims=True)
    normals /= np.maximum(length, 1e-10)
    if return_neighbours:
        return normals, nbr
    return normals


def point_to_line_step(moving, target, normals, match):
    """One linearised point-to-line step.  Restates utilities/icp.py:79-115.

    Unknowns x = [theta, tx, ty]; rows [ny*px - nx*py, nx, ny]; right-hand side
    -(n . (p - q)); 3x3 normal equations solved by LAPACK gesv; a singular
    system yields the identity step (icp.py:105-108).
    """
    q = target[match]
    nm = normals[match]
    nx, ny = nm[:, 0], nm[:, 1]
    px, py = moving[:, 0], moving[:, 1]
    dx, dy = px - q[:, 0], py - q[:, 1]
    c = ny * px - nx * py
    a = np.column_stack([c, nx, ny])
    b = -(nx * dx + ny * dy)
    ata = a.T @ a
    atb = a.T @ b
    try:
        sol = np.linalg.solve(ata, atb)
    except np.linalg.LinAlgError:
        return np.eye(2), np.zeros(2)
    th, tx, ty = sol
    ct, st = np.cos(th), np.sin(th)
    return np.array([[ct, -st], [st, ct]]), np.array([tx, ty])


def kabsch_step(moving, matched):
    """One point-to-point (Kabsch/SVD) step.  Restates utilities/icp.py:196-207
    (centroids per icp.py:32-33)."""
    mu_s = np.array(np.mean(moving, axis=0))
    mu_t = np.array(np.mean(matched, axis=0))
    w = np.dot((moving - mu_s).T, matched - mu_t)
    u, _, vt = np.linalg.svd(w)
    r = np.dot(vt.T, u.T)
    if np.linalg.det(r) < 0:
        vt[-1, :] *= -1
        r = np.dot(vt.T, u.T)
    return r, mu_t - np.dot(r, mu_s)


def register(source, target, error_threshold, max_iterations, voxel_size,
             R_init=None, t_init=None, method="point_to_point", normal_k=10,
             max_corr_dist=None, trace=None):
    """Full registration loop.  Restates utilities/icp.py:132-223.

    Returns ``(R, t, error, iters, status)``; the first three are exactly the
    reference's return tuple.  ``iters`` counts completed solve steps and
    ``status`` is one of the STATUS_* codes.  If ``trace`` is a dict it receives
    the downsampled clouds, normals and per-iteration correspondence indices
    (used by the correspondence-parity tests).
    """
    src = voxel_means(source, voxel_size)                       # icp.py:150
    tgt = voxel_means(target, voxel_size)                       # icp.py:151
    dim = src.shape[1]

    if R_init is not None and t_init is not None:               # icp.py:153-160
        cur = src @ R_init.T + t_init
        r_tot = R_init.copy()
        t_tot = t_init.copy()
    else:
        cur = src.copy()
        r_tot = np.eye(dim)
        t_tot = np.zeros(dim)

    p2l = (method == "point_to_line" and dim == 2)              # icp.py:162
    normals = pca_normals_2d(tgt, k=normal_k) if p2l else None  # icp.py:165-167
    gate = max_corr_dist ** 2 if max_corr_dist is not None else None   # icp.py:169
    tree = KDTree(tgt)                                          # icp.py:173

    if trace is not None:
        trace.update(src=src, tgt=tgt, normals=normals, matches=[], errors=[])

    prev = float("inf")
    err = float("inf")
    iters = 0
    status = STATUS_MAX_ITER
    for _ in range(max_iterations):                             # icp.py:177
        dist, match = tree.query(cur)                           # icp.py:179
        near = tgt[match]                                       # icp.py:180
        if trace is not None:
            trace["matches"].append(match.copy())
        if gate is not None:                                    # icp.py:183-189
            keep = dist ** 2 < gate
            if keep.sum() < max(3, len(cur) // 10):
                status = STATUS_FEW_INLIERS
                break
        else:
            keep = np.ones(len(cur), dtype=bool)

        if p2l:                                                 # icp.py:192-195
            r, t = point_to_line_step(cur[keep], tgt, normals, match[keep])
        else:                                                   # icp.py:197-207
            r, t = kabsch_step(cur[keep], near[keep])

        r_tot = np.dot(r, r_tot)                                # icp.py:210
        t_tot = np.dot(t_tot, r.T) + t                          # icp.py:211
        cur = np.dot(cur, r.T) + t                              # icp.py:212
        iters += 1

        err = np.mean(np.sum((near - cur) ** 2, axis=1))        # icp.py:215
        if trace is not None:
            trace["errors"].append(float(err))
        delta = abs(prev - err)                                 # icp.py:216
        if delta < error_threshold:                             # icp.py:217-219
            status = STATUS_CONVERGED
            break
        prev = err                                              # icp.py:220
    return r_tot, t_tot, err, iters, status


def result_line(err, iters, status, max_iterations):
    """The one console line the reference prints per call (icp.py:218, 222).

    The converged line also carries ``delta``; callers that need it pass the
    text through from their own bookkeeping -- here only the parts derivable
    from (err, iters, status) are produced, for shim-format tests.
    """
    if status == STATUS_CONVERGED:
        return f"  ICP converged: iter={iters - 1}, error={err:.8f}"
    return f"  ICP max iterations reached: iter={max_iterations}, error={err:.8f}"
