"""GPU parity: rotation-search pre-alignment (features.py:165-242, slam.py:111-183) through the C ABI
against the oracle and the reference's golden outputs.  The scores are fp64 means of exact nearest
distances: tolerance 1e-12 relative (summation order); the winning angle must be the reference's."""
import numpy as np
import pytest

from conftest import load_golden
from icp_b200 import api, synth
from oracle import features_oracle as fo
from oracle.icp_oracle import voxel_means

pytestmark = pytest.mark.gpu


def test_scores_match_the_oracle_for_every_angle():
    scans, _ = synth.make_sequence(4, world="room", seed=11)
    src, tgt = voxel_means(scans[0], 0.15), voxel_means(scans[2], 0.15)
    src_c, mu_t = src - src.mean(axis=0), tgt.mean(axis=0)
    angles = np.deg2rad(np.arange(-180, 180, 1.5))
    got = api.rotation_scores([src_c], [tgt], [angles], [mu_t])[0]
    ref = fo.sweep_scores(src_c, tgt, angles, mu_t)
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=0)
    assert int(np.argmin(got)) == int(np.argmin(ref))


def test_rotation_search_reproduces_the_reference(capsys):
    from utilities import rotation_search
    g = load_golden("rotation.npz")
    for name in ("cfg", "defaults", "turned"):
        v, c, f = g[f"rs_{name}_kw"]
        R, t, score = rotation_search(g[f"rs_{name}_src"], g[f"rs_{name}_tgt"], voxel_size=v, angle_step_coarse=c,
                                      angle_step_fine=f)
        line = capsys.readouterr().out
        assert line.startswith("  Rotation search: best angle ") and "score" in line
        np.testing.assert_allclose(R, g[f"rs_{name}_R"], rtol=0, atol=1e-15)         # same angle => same matrix
        np.testing.assert_allclose(t, g[f"rs_{name}_t"], rtol=0, atol=1e-12)
        assert abs(score - float(g[f"rs_{name}_score"])) <= 1e-12 * float(g[f"rs_{name}_score"])
    # too few points after the coarse voxel grid: identity, inf (features.py:201-202)
    v, c, f = g["rs_tiny_kw"]
    R, t, score = rotation_search(g["rs_tiny_src"], g["rs_tiny_tgt"], voxel_size=v, angle_step_coarse=c, angle_step_fine=f)
    assert np.array_equal(R, np.eye(2)) and np.array_equal(t, np.zeros(2)) and score == float("inf")


def test_submap_rotation_search_reproduces_the_reference():
    from utilities import submap_rotation_search
    g = load_golden("rotation.npz")
    for name in ("cfg", "defaults"):
        ar, st, fs, v = g[f"sub_{name}_kw"]
        R, t = submap_rotation_search(g[f"sub_{name}_src"], g[f"sub_{name}_map"], g[f"sub_{name}_pose"], angle_range=ar,
                                      angle_step=st, fine_step=fs, voxel_size=v)
        np.testing.assert_allclose(R, g[f"sub_{name}_R"], rtol=0, atol=1e-15)
        np.testing.assert_allclose(t, g[f"sub_{name}_t"], rtol=0, atol=1e-9)


def test_nearest_neighbour_output_and_batching():
    """want_nn returns what KDTree.query returns; many problems in one call equal one call each."""
    from scipy.spatial import KDTree
    rng = np.random.default_rng(2)
    probs = []
    for k in range(5):
        n_s, n_t = int(rng.integers(5, 700)), int(rng.integers(5, 3000))
        probs.append((rng.normal(scale=5, size=(n_s, 2)), rng.normal(scale=5, size=(n_t, 2)),
                      rng.uniform(-3, 3, size=int(rng.integers(1, 40))), rng.normal(size=2)))
    many = api.rotation_scores([p[0] for p in probs], [p[1] for p in probs], [p[2] for p in probs], [p[3] for p in probs])
    for k, (s, t, a, sh) in enumerate(probs):
        one = api.rotation_scores([s], [t], [a], [sh])[0]
        assert one.tobytes() == many[k].tobytes()
        np.testing.assert_allclose(one, fo.sweep_scores(s, t, a, sh), rtol=1e-12)
        _, d, i = api.rotation_scores([s], [t], [a[:1]], [sh], want_nn=True)
        ca, sa = np.cos(a[0]), np.sin(a[0])
        dd, ii = KDTree(t).query(s @ np.array([[ca, -sa], [sa, ca]]).T + sh)
        assert np.array_equal(i[0], ii.astype(np.int32))
        np.testing.assert_allclose(d[0], dd, rtol=1e-12)


def test_argument_errors():
    s = np.zeros((6, 2))
    with pytest.raises(RuntimeError, match="one angle"):
        api.rotation_scores([s], [s], [np.zeros(3)], [np.zeros(2)], want_nn=True)


def test_targets_beyond_shared_memory_are_swept_in_slices():
    """A downsampled submap can exceed the 8192 points a CTA stages (slam.py:125-126 puts no bound on it): the target is
    swept in slices with the running nearest neighbour of every (angle, point) kept in HBM.  Scores within 1e-12 of the
    oracle, nearest indices those of KDTree, and a mixed batch (one big target, one small) equals separate calls."""
    from scipy.spatial import KDTree
    rng = np.random.default_rng(5)
    big = rng.uniform(-40, 40, size=(20000, 2))
    small = rng.uniform(-40, 40, size=(900, 2))
    src = rng.uniform(-30, 30, size=(777, 2))
    angles = rng.uniform(-3, 3, size=11)
    shift = np.array([0.7, -1.3])
    got = api.rotation_scores([src], [big], [angles], [shift])[0]
    np.testing.assert_allclose(got, fo.sweep_scores(src, big, angles, shift), rtol=1e-12)
    _, d, i = api.rotation_scores([src], [big], [angles[:1]], [shift], want_nn=True)
    ca, sa = np.cos(angles[0]), np.sin(angles[0])
    dd, ii = KDTree(big).query(src @ np.array([[ca, -sa], [sa, ca]]).T + shift)
    assert np.array_equal(i[0], ii.astype(np.int32))
    np.testing.assert_allclose(d[0], dd, rtol=1e-12)
    both = api.rotation_scores([src, src[:100]], [big, small], [angles, angles[:3]], [shift, shift])
    assert both[0].tobytes() == got.tobytes()
    assert both[1].tobytes() == api.rotation_scores([src[:100]], [small], [angles[:3]], [shift])[0].tobytes()
