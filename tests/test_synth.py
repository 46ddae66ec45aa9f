"""Synthetic generator: deterministic, right shapes, usable as bench input."""
import numpy as np

from icp_b200 import synth


def test_sequences_are_deterministic_and_shaped():
    a, pa = synth.make_sequence(3, world="room", seed=0)
    b, pb = synth.make_sequence(3, world="room", seed=0)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and np.array_equal(pa, pb)
    for s in a:
        assert s.ndim == 2 and s.shape[1] == 2 and 900 <= len(s) <= synth.N_BEAMS
        assert np.all(np.hypot(s[:, 0], s[:, 1]) <= synth.MAX_RANGE + 0.1)
    step = np.hypot(*(pa[1:, :2] - pa[:-1, :2]).T)
    assert np.all(step < 0.2)
    c, _ = synth.make_sequence(2, world="campus", seed=1)
    assert all(len(s) > 300 for s in c)


def test_pack_and_pairs():
    clouds = [np.zeros((3, 2)), np.ones((5, 2))]
    flat, off = synth.pack_ragged(clouds)
    assert flat.shape == (8, 2) and off.tolist() == [0, 3, 8] and flat.flags.c_contiguous
    _, poses = synth.make_sequence(1, world="room")
    poses = synth.room_trajectory(400)
    pairs = synth.loop_closure_pairs(poses, 64, seed=0)
    d = np.hypot(*(poses[pairs[:, 0], :2] - poses[pairs[:, 1], :2]).T)
    assert pairs.shape == (64, 2) and np.all(d < 3.0) and np.all(pairs[:, 0] != pairs[:, 1])
    sub = synth.submap_cloud(n_raw=5000)
    assert sub.shape == (5000, 2)
