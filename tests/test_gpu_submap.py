"""Device-resident submap (slam.py:103-108, 217-225, 559-562, 611-615) against the reference's way of
doing it: a Python list of global-frame scans, np.vstack + voxel_downsample every scan, then ICP(scan, submap, ...)."""
import numpy as np
import pytest

from conftest import pose_delta
from icp_b200 import api, synth
from oracle import icp_oracle

pytestmark = pytest.mark.gpu

KW = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, method="point_to_point", max_corr_dist=1.5)


def _rot(th):
    return np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])


def test_window_build_and_registration_match_the_reference_recipe():
    scans, poses = synth.make_sequence(60, world="room", seed=12)
    world_scans = [synth.to_world_frame(s, p) for s, p in zip(scans, poses)]
    sm = api.DeviceSubmap(capacity_scans=40)
    window = []
    for k in range(55):                                       # slam.py:559-562
        sm.append(world_scans[k]); window.append(world_scans[k])
        if len(window) > 40:
            window.pop(0)
    assert len(sm) == 40 and sm.size()[1] == sum(len(w) for w in window)
    SV = 0.01                                                 # submap voxel: fine enough that the window stays a big target
    want = icp_oracle.voxel_means(np.vstack(window), SV)     # slam.py:103-108
    got = sm.build(SV)
    assert len(want) > 4096                                  # a big target: the hash-grid path
    assert got.shape == want.shape and got.tobytes() == want.tobytes()
    # register the next scans against the window (slam.py:217-225), initial guess = true pose perturbed
    srcs, R0, t0 = [], [], []
    for k in range(55, 59):
        srcs.append(scans[k]); R0.append(_rot(poses[k, 2] + 0.01)); t0.append(poses[k, :2] + [0.05, -0.04])
    R0, t0 = np.asarray(R0), np.asarray(t0)
    out = sm.icp(srcs, SV, R_init=R0, t_init=t0, **KW)
    again = sm.icp(srcs, SV, R_init=R0, t_init=t0, **KW)   # the window did not change: cached preprocessing, same bits
    plain = api.icp_batch(srcs, [want] * 4, R_init=R0, t_init=t0, **KW)
    for key in ("R", "t", "error", "iters", "status"):
        assert again[key].tobytes() == out[key].tobytes(), key
        assert plain[key].tobytes() == out[key].tobytes(), key
    for p in range(4):
        R, t, err, iters, status = icp_oracle.register(srcs[p], want, R_init=R0[p], t_init=t0[p], **KW)
        dt, dr = pose_delta(out["R"][p], out["t"][p], R, t)
        assert dt < 1e-4 and dr < 1e-5 and int(out["iters"][p]) == iters and int(out["status"][p]) == status
        assert np.abs(out["t"][p] - poses[55 + p, :2]).max() < 0.05          # and it found the pose
    one = sm.icp(srcs[0], SV, R_init=R0[:1], t_init=t0[:1], **KW)          # a single (N, 2) array is one source
    assert one["R"][0].tobytes() == out["R"][0].tobytes()
    # another ICP voxel size re-does the second downsample; another submap voxel the first
    coarse = sm.icp(srcs[:1], 0.1, R_init=R0[:1], t_init=t0[:1], **dict(KW, voxel_size=0.06))
    ref_c = api.icp_batch(srcs[:1], [icp_oracle.voxel_means(np.vstack(window), 0.1)], R_init=R0[:1], t_init=t0[:1], **dict(KW, voxel_size=0.06))
    assert coarse["t"].tobytes() == ref_c["t"].tobytes() and coarse["iters"][0] == ref_c["iters"][0]
    # the window moves on: results follow it
    sm.append(world_scans[55]); window.append(world_scans[55]); window.pop(0)
    want2 = icp_oracle.voxel_means(np.vstack(window), SV)
    out2 = sm.icp(srcs[1:2], SV, R_init=R0[1:2], t_init=t0[1:2], **KW)
    plain2 = api.icp_batch(srcs[1:2], [want2], R_init=R0[1:2], t_init=t0[1:2], **KW)
    assert out2["t"].tobytes() == plain2["t"].tobytes()
    # loop closure: the buffer is rebuilt from corrected poses (slam.py:611-615)
    sm.clear()
    assert len(sm) == 0
    with pytest.raises(RuntimeError, match="empty"):
        sm.icp(srcs[:1], SV, R_init=R0[:1], t_init=t0[:1], **KW)
    for k in range(5):
        sm.append(world_scans[k])
    small = sm.build(0.04)
    assert small.tobytes() == icp_oracle.voxel_means(np.vstack(world_scans[:5]), 0.04).tobytes()
    sm.close()


def test_small_window_takes_the_brute_path_and_point_to_line():
    """A window of two scans downsampled coarsely stays under 4096 points: shared-memory brute-force path, with normals."""
    scans, poses = synth.make_sequence(6, world="room", seed=3)
    world_scans = [synth.to_world_frame(s, p) for s, p in zip(scans, poses)]
    sm = api.DeviceSubmap(capacity_scans=2)
    for k in range(3):
        sm.append(world_scans[k])
    target = icp_oracle.voxel_means(np.vstack(world_scans[1:3]), 0.1)
    assert len(target) <= 4096 and sm.build(0.1).tobytes() == target.tobytes()
    kw = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, method="point_to_line", normal_k=12)
    R0, t0 = _rot(poses[3, 2])[None], poses[3:4, :2].copy()
    out = sm.icp(scans[3], 0.1, R_init=R0, t_init=t0, **kw)
    R, t, err, iters, status = icp_oracle.register(scans[3], target, R_init=R0[0], t_init=t0[0], **kw)
    dt, dr = pose_delta(out["R"][0], out["t"][0], R, t)
    assert dt < 1e-4 and dr < 1e-5 and int(out["iters"][0]) == iters and int(out["status"][0]) == status
