"""CPU oracle for the rotation-search pre-alignment -- TEST INFRASTRUCTURE ONLY.

numpy/scipy restatement of /root/reference/utilities/features.py:165-242
(``rotation_search``) and /root/reference/slam.py:111-183
(``_submap_rotation_search``), making the same library calls in the same order
(``scipy.spatial.KDTree``, ``np.mean``, ``np.argmin``, ``np.percentile``).
Pinned bit-exact against the live reference by ``oracle/pin_rotation.py``
(tests/golden/rotation.npz).  Only ``tests/`` and ``bench.py``'s CPU-baseline
legs may import this module; the product path never does.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import KDTree

from .icp_oracle import voxel_means as voxel_downsample


def sweep_scores(src, tgt, angles, shift):
    """features.py:205-211 / slam.py:138-143 for a list of angles."""
    tree = KDTree(tgt)
    out = np.empty(len(angles))
    for k, a in enumerate(angles):
        ca, sa = np.cos(a), np.sin(a)
        R = np.array([[ca, -sa], [sa, ca]])
        dists, _ = tree.query(src @ R.T + shift)
        out[k] = np.mean(dists ** 2)
    return out


def rotation_search(source, target, voxel_size=0.3, angle_step_coarse=2.0, angle_step_fine=0.2):
    """features.py:165-242; returns (R, t, score, coarse scores, fine scores, fine angles)."""
    src = voxel_downsample(source, voxel_size)
    tgt = voxel_downsample(target, voxel_size)
    if len(src) < 5 or len(tgt) < 5:
        return np.eye(2), np.zeros(2), float("inf"), None, None, None
    mu_s, mu_t = src.mean(axis=0), tgt.mean(axis=0)
    src_c = src - mu_s
    angles_coarse = np.deg2rad(np.arange(-180, 180, angle_step_coarse))
    sc = sweep_scores(src_c, tgt, angles_coarse, mu_t)
    best_angle = angles_coarse[int(np.argmin(sc))]
    lo, hi = best_angle - np.deg2rad(angle_step_coarse), best_angle + np.deg2rad(angle_step_coarse)
    angles_fine = np.arange(lo, hi, np.deg2rad(angle_step_fine))
    sf = sweep_scores(src_c, tgt, angles_fine, mu_t)
    k = int(np.argmin(sf))
    best_angle, best_score = angles_fine[k], sf[k]
    ca, sa = np.cos(best_angle), np.sin(best_angle)
    R = np.array([[ca, -sa], [sa, ca]])
    return R, mu_t - R @ mu_s, best_score, sc, sf, angles_fine


def submap_rotation_search(source_local, submap_global, predicted_pose, angle_range=60.0, angle_step=2.0,
                           fine_step=0.5, voxel_size=0.3):
    """slam.py:111-183; returns (R, t)."""
    src = voxel_downsample(source_local, voxel_size)
    tgt = voxel_downsample(submap_global, voxel_size)
    if len(src) < 5 or len(tgt) < 5:
        return predicted_pose[:2, :2], predicted_pose[:2, 2]
    pred_t = predicted_pose[:2, 2]
    pred_theta = np.arctan2(predicted_pose[1, 0], predicted_pose[0, 0])
    tree = KDTree(tgt)
    offsets = np.deg2rad(np.arange(-angle_range, angle_range + angle_step, angle_step))
    angles = pred_theta + offsets
    scores = sweep_scores(src, tgt, angles, pred_t)
    best_angle = angles[int(np.argmin(scores))]
    fine_angles = np.arange(best_angle - np.deg2rad(angle_step), best_angle + np.deg2rad(angle_step), np.deg2rad(fine_step))
    if len(fine_angles) > 0:
        best_angle = fine_angles[int(np.argmin(sweep_scores(src, tgt, fine_angles, pred_t)))]
    ca, sa = np.cos(best_angle), np.sin(best_angle)
    R_best = np.array([[ca, -sa], [sa, ca]])
    rotated_src = src @ R_best.T
    nn_dists, nn_idx = tree.query(rotated_src + pred_t)
    nn_dists_sq = nn_dists ** 2
    inlier_mask = nn_dists_sq <= np.percentile(nn_dists_sq, 80)
    if inlier_mask.sum() >= 5:
        refined_t = np.mean(tgt[nn_idx][inlier_mask] - rotated_src[inlier_mask], axis=0)
    else:
        refined_t = pred_t
    return R_best, refined_t
