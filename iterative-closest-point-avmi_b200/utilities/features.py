"""GPU-backed pre-alignment sweeps -- the callers' step in front of every ICP call.

``rotation_search`` keeps the reference's name, arguments, defaults, return
tuple and console line (/root/reference/utilities/features.py:165-242; called
from slam.py:60-66).  ``submap_rotation_search`` restates
``slam.py:111-183`` (``_submap_rotation_search``; it lives in slam.py upstream,
so using it means replacing that one function -- see INTEGRATION.md section 3).
Both run every angle of a sweep in ONE launch of libicp_b200's rotation-score
kernel: the angle lists, the argmin and the output transform are computed on
the host exactly as the reference computes them.  The rest of
``utilities/features.py`` (curvature key points, descriptors, RANSAC behind
``feature_based_alignment``, features.py:22-160, 247-315) is outside the
accelerated path and stays the reference's: when this package shadows the
reference's ``utilities`` (second option) those names are
adopted from the reference's own ``features.py`` found through the package path.
"""
import os
import sys

import numpy as np

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)

from icp_b200 import api as _api          # noqa: E402


def rotation_search(source, target, voxel_size=0.3, angle_step_coarse=2.0, angle_step_fine=0.2):
    """Correlative rotation search; returns ``(R (2,2), t (2,), score)``."""
    src = _api.voxel_downsample(np.asarray(source, dtype=np.float64), voxel_size)       # features.py:198-199
    tgt = _api.voxel_downsample(np.asarray(target, dtype=np.float64), voxel_size)
    if len(src) < 5 or len(tgt) < 5:                                                   # features.py:201-202
        return np.eye(2), np.zeros(2), float("inf")
    mu_s = src.mean(axis=0)
    mu_t = tgt.mean(axis=0)
    src_c = src - mu_s
    angles_coarse = np.deg2rad(np.arange(-180, 180, angle_step_coarse))                # features.py:214
    scores_coarse = _api.rotation_scores([src_c], [tgt], [angles_coarse], [mu_t])[0]
    best_angle = angles_coarse[int(np.argmin(scores_coarse))]
    lo = best_angle - np.deg2rad(angle_step_coarse)                                    # features.py:220-223
    hi = best_angle + np.deg2rad(angle_step_coarse)
    angles_fine = np.arange(lo, hi, np.deg2rad(angle_step_fine))
    scores_fine = _api.rotation_scores([src_c], [tgt], [angles_fine], [mu_t])[0]
    best_idx_f = int(np.argmin(scores_fine))
    best_angle = angles_fine[best_idx_f]
    best_score = scores_fine[best_idx_f]
    ca, sa = np.cos(best_angle), np.sin(best_angle)
    R = np.array([[ca, -sa], [sa, ca]])
    t = mu_t - R @ mu_s
    print(f"  Rotation search: best angle {np.degrees(best_angle):.1f}\u00b0, "
          f"score {best_score:.4f}")
    return R, t, best_score


def submap_rotation_search(source_local, submap_global, predicted_pose, angle_range=60.0, angle_step=2.0,
                           fine_step=0.5, voxel_size=0.3):
    """slam.py:111-183: rotation sweep around the predicted pose, then one nearest-neighbour
    translation step on the closest 80 % of the correspondences.  Returns ``(R (2,2), t (2,))``."""
    predicted_pose = np.asarray(predicted_pose, dtype=np.float64)
    src = _api.voxel_downsample(np.asarray(source_local, dtype=np.float64), voxel_size)
    tgt = _api.voxel_downsample(np.asarray(submap_global, dtype=np.float64), voxel_size)
    if len(src) < 5 or len(tgt) < 5:                                                   # slam.py:127-128
        return predicted_pose[:2, :2], predicted_pose[:2, 2]
    pred_t = predicted_pose[:2, 2]
    pred_theta = np.arctan2(predicted_pose[1, 0], predicted_pose[0, 0])
    offsets = np.deg2rad(np.arange(-angle_range, angle_range + angle_step, angle_step))   # slam.py:146-148
    angles = pred_theta + offsets
    scores = _api.rotation_scores([src], [tgt], [angles], [pred_t])[0]
    best_angle = angles[int(np.argmin(scores))]
    fine_lo = best_angle - np.deg2rad(angle_step)                                      # slam.py:153-158
    fine_hi = best_angle + np.deg2rad(angle_step)
    fine_angles = np.arange(fine_lo, fine_hi, np.deg2rad(fine_step))
    if len(fine_angles) > 0:
        fine_scores = _api.rotation_scores([src], [tgt], [fine_angles], [pred_t])[0]
        best_angle = fine_angles[int(np.argmin(fine_scores))]
    correction = np.degrees(best_angle - pred_theta)
    if abs(correction) > 1.0:
        print(f"  Submap rotation correction: {correction:+.1f}\u00b0")
    ca, sa = np.cos(best_angle), np.sin(best_angle)
    R_best = np.array([[ca, -sa], [sa, ca]])
    rotated_src = src @ R_best.T                                                       # slam.py:166-181
    _, nn_d, nn_i = _api.rotation_scores([src], [tgt], [np.array([best_angle])], [pred_t], want_nn=True)
    nn_dists_sq = nn_d[0] ** 2
    dist_thresh = np.percentile(nn_dists_sq, 80)
    inlier_mask = nn_dists_sq <= dist_thresh
    if inlier_mask.sum() >= 5:
        matched = tgt[nn_i[0]]
        refined_t = np.mean(matched[inlier_mask] - rotated_src[inlier_mask], axis=0)
    else:
        refined_t = pred_t
    return R_best, refined_t


# ---- the rest of the reference's module, unchanged (fall-through) -----------------------------------
def _adopt_reference_features():
    """Load the reference's own utilities/features.py (the next one on the package path) as a private submodule and
    re-export every public name this file does not replace.  Its `from .icp import voxel_downsample`
    (features.py:198, 276) then resolves to this package's GPU `voxel_downsample`."""
    import importlib.util
    pkg = sys.modules.get(__package__)
    for d in list(getattr(pkg, "__path__", []))[1:]:
        path = os.path.join(d, "features.py")
        if not os.path.isfile(path):
            continue
        spec = importlib.util.spec_from_file_location(__package__ + "._reference_features", path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        for name in dir(mod):
            if not name.startswith("__") and name not in globals():
                globals()[name] = getattr(mod, name)
        return True
    return False


if not _adopt_reference_features():
    def feature_based_alignment(*args, **kwargs):
        """features.py:247-315 is not part of the accelerated path; it is the reference's own function whenever the
        reference is importable (see the module docstring)."""
        raise ImportError("utilities.features.feature_based_alignment is the reference's own function: put the reference "
                          "root on sys.path behind this package (INTEGRATION.md section 1)")
