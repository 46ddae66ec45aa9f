"""GPU parity: occupancy raycast through the C ABI vs the oracle / golden
fixtures.  Bar: bit-exact float32 log-odds (north_star: Bresenham cell sets
bit-exact, log-odds within 1e-9 abs -- i.e. identical float32 values)."""
import os

import numpy as np
import pytest

from conftest import load_golden
from icp_b200 import api, synth
from oracle import occupancy_oracle as oo

pytestmark = pytest.mark.gpu

GKW = dict(resolution=0.05, p_hit=0.85, p_miss=0.42, log_odds_min=-8.0, log_odds_max=8.0)


@pytest.fixture(params=["fast", "ordered"], autouse=True)
def occ_path(request, monkeypatch):
    """Every test runs on both device paths: the order-free pipeline (default)
    and the ordered tile replay (read by icpb200_grid_create)."""
    monkeypatch.setenv("ICPB200_OCC_PATH", request.param)
    return request.param


def make_pair(bounds, **kw):
    from utilities import OccupancyGrid2D
    return OccupancyGrid2D(*bounds, **kw), oo.GridOracleC(*bounds, **kw)


def assert_same(gpu_grid, ref_grid, what=""):
    a, b = gpu_grid.log_odds, ref_grid.log_odds
    assert a.dtype == np.float32 and a.shape == b.shape
    if a.tobytes() != b.tobytes():
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(bad)} cells differ, first {bad[:5].tolist()} "
                             f"gpu={a[tuple(bad[0])]!r} ref={b[tuple(bad[0])]!r}")


def test_golden_fixture_scan_by_scan():
    """The reference's own outputs (tests/golden/occupancy.npz) reproduced
    through update_scan, one scan per call: duplicates, zero-length rays, an
    empty scan, an out-of-grid origin and out-of-bounds endpoints included."""
    from utilities import OccupancyGrid2D
    g = load_golden("occupancy.npz")
    grid = OccupancyGrid2D(*g["bounds"], **GKW)
    assert (grid.ny, grid.nx) == g["snap_0"].shape
    assert grid.l_hit == 1.7346010553881064 and grid.l_miss == -0.3227733922630512
    off = g["hit_off"]
    for s in range(40):
        grid.update_scan(g["origins"][s], g["hits"][off[s]:off[s + 1]])
        if s in (0, 13, 39):
            assert grid.log_odds.tobytes() == g[f"snap_{s}"].tobytes(), f"after scan {s}"
    grid.reset()
    assert not grid.log_odds.any()
    grid.update_scan(g["origins"][0], g["hits"][off[0]:off[1]])
    assert grid.log_odds.tobytes() == g["snap_0"].tobytes()


def test_golden_fixture_batched_and_clamp_excluding_zero():
    from utilities import OccupancyGrid2D
    g = load_golden("occupancy.npz")
    grid = OccupancyGrid2D(*g["bounds"], **GKW)
    grid._dev.update(g["origins"], g["hits"], g["hit_off"])
    assert grid._dev.read().tobytes() == g["snap_39"].tobytes()
    st = grid._dev.last_stats()
    assert st["rays"] == int(g["hit_off"][-1]) and st["traversed"] > 0 and st["hits"] > 0
    odd = OccupancyGrid2D(*g["bounds"], **dict(GKW, log_odds_min=0.5, log_odds_max=3.0))
    off = g["hit_off"]
    for s in range(3):
        odd.update_scan(g["origins"][s], g["hits"][off[s]:off[s + 1]])
    assert odd.log_odds.tobytes() == g["odd_final"].tobytes()
    odd2 = OccupancyGrid2D(*g["bounds"], **dict(GKW, log_odds_min=0.5, log_odds_max=3.0))
    odd2._dev.update(g["origins"][:3], g["hits"][:off[3]], off[:4])
    assert odd2._dev.read().tobytes() == g["odd_final"].tobytes()


def test_random_scans_vs_oracle_non_multiple_of_tile():
    """Grid whose size is not a multiple of the 64-cell tile, rays from random
    origins (inside and outside), many collisions in few cells."""
    rng = np.random.default_rng(5)
    bounds = (-7.3, 9.1, -4.4, 6.05)                       # 328 x 209 cells
    gpu, ref = make_pair(bounds, **GKW)
    for s in range(25):
        n = int(rng.integers(1, 700))
        org = rng.uniform([-9, -6], [11, 8])
        pts = org + rng.normal(size=(n, 2)) * rng.uniform(0.05, 6.0)
        if s % 5 == 0:
            pts[: n // 2] = pts[0]                         # heavy duplicates in one cell
        gpu.update_scan(org, pts)
        ref.update_scan(org, pts)
        if s % 6 == 0:
            assert_same(gpu, ref, f"scan {s}")
    assert_same(gpu, ref, "final")


def test_saturation_and_sign_conventions():
    """Many repeats drive cells into both clamps; p_miss > 0.5 flips the sign
    of l_miss (no early-exit shortcut may change the result)."""
    rng = np.random.default_rng(8)
    bounds = (-4.0, 4.0, -4.0, 4.0)
    for kw in (dict(GKW, log_odds_min=-2.0, log_odds_max=3.5),
               dict(resolution=0.1, p_hit=0.3, p_miss=0.6, log_odds_min=-1.0, log_odds_max=1.5),
               dict(resolution=0.1, p_hit=0.7, p_miss=0.5, log_odds_min=-5.0, log_odds_max=5.0)):
        gpu, ref = make_pair(bounds, **kw)
        pts0 = rng.uniform(-3.5, 3.5, size=(300, 2))
        origins = [np.array([0.2, -0.1])] * 30
        clouds = [pts0 + rng.normal(0, 0.01, size=pts0.shape) for _ in range(30)]
        gpu.update_scans(origins, clouds)
        for o, c in zip(origins, clouds):
            ref.update_scan(o, c, fast=True)
        assert_same(gpu, ref, str(kw))


def test_campus_replay_batched_vs_oracle():
    """300 synthetic scans replayed in one call (the _rebuild_map shape) on a
    2048 x 2048 grid, against the C oracle scan by scan."""
    scans, poses = synth.make_sequence(300, world="campus", seed=1)
    hits = [synth.to_world_frame(s, p) for s, p in zip(scans, poses)]
    bounds = (poses[:, 0].mean() - 51.2, poses[:, 0].mean() + 51.2,
              poses[:, 1].mean() - 51.2, poses[:, 1].mean() + 51.2)
    gpu, ref = make_pair(bounds, **GKW)
    assert gpu.nx == 2048 and gpu.ny == 2048
    flat, off = synth.pack_ragged(hits)
    gpu._dev.update(poses[:, :2].copy(), flat, off)
    cells = ref.update_many(poses[:, :2].copy(), flat, off, fast=True)
    assert_same(gpu, ref, "campus replay")
    st = gpu._dev.last_stats()
    assert st["traversed"] == cells                       # same number of free-cell updates
    # splitting the batch anywhere gives the same map (scan order is preserved)
    gpu2, _ = make_pair(bounds, **GKW)
    gpu2._dev.update(poses[:120, :2].copy(), flat[:off[120]], off[:121])
    gpu2._dev.update(poses[120:, :2].copy(), flat[off[120]:], off[120:] - off[120])
    assert gpu2.log_odds.tobytes() == gpu.log_odds.tobytes()


def test_spatial_shards_sum_to_whole_map():
    """Multi-GPU seam emulated on one GPU: world = 4 and world = 3 shards (bands of 64 rows dealt round-robin, the rule
    of icp_b200.dist.owned_rows), each replaying all scans clipped to its own bands.  Every shard writes only rows it
    owns, bit-exactly the rows of the unsharded map, so the gather of the owned rows is the whole map -- also in a second
    round without a reset, when the rows a shard does not own hold stale values."""
    from icp_b200 import dist as icpd
    scans, poses = synth.make_sequence(40, world="room", seed=2)
    hits = [synth.to_world_frame(s, p) for s, p in zip(scans, poses)]
    flat, off = synth.pack_ragged(hits)
    half = int(off[20])
    bounds = (-25.6, 25.6, -25.6, 25.6)
    whole, ref = make_pair(bounds, **GKW)
    whole._dev.update(poses[:, :2].copy(), flat, off)
    ref.update_many(poses[:, :2].copy(), flat, off)
    assert_same(whole, ref, "unsharded")
    first, _ = make_pair(bounds, **GKW)
    first._dev.update(poses[:20, :2].copy(), flat[:half], off[:21])
    first_map = first.log_odds.copy()
    for world in (4, 3):
        total = np.zeros_like(whole.log_odds)
        for rank in range(world):
            part, _ = make_pair(bounds, **GKW)
            part._dev.set_shard(rank, world)
            rows = icpd.owned_rows(part.ny, rank, world)
            # round 1: the first 20 scans; round 2, no reset: the other 20 on top
            part._dev.update(poses[:20, :2].copy(), flat[:half], off[:21])
            piece = part._dev.read()                  # (the shim's log_odds property caches: read the device directly)
            assert not piece[~rows].any(), f"world {world} rank {rank} wrote rows it does not own"
            assert piece[rows].tobytes() == first_map[rows].tobytes()
            part._dev.update(poses[20:, :2].copy(), flat[half:], off[20:] - off[20])
            piece = part._dev.read()
            assert not piece[~rows].any()
            total[rows] = piece[rows]
        assert total.tobytes() == whole.log_odds.tobytes(), world


def test_empty_inputs_and_limits():
    from utilities import OccupancyGrid2D
    grid = OccupancyGrid2D(-1, 1, -1, 1, **GKW)
    grid.update_scan([0.0, 0.0], np.zeros((0, 2)))          # mapping.py:113-114
    assert not grid.log_odds.any()
    grid.update_scan([0.0, 0.0], np.array([[0.001, 0.001]]))  # hit in the origin cell: no free cells
    lo = grid.log_odds
    assert np.count_nonzero(lo) == 1 and lo.max() == np.float32(1.7346010553881064)
    grid.update_scan([0.0, 0.0], np.array([[1e12, -3e11]]))   # absurdly far endpoint: clipped walk
    assert np.isfinite(grid.log_odds).all()
    with pytest.raises(RuntimeError):
        grid._dev.update(np.zeros((1, 2)), np.zeros((3, 2)), np.array([1, 3]))   # hit_off[0] != 0


def test_many_scans_cross_the_chunk_boundary_and_neutral_miss():
    """2300 small scans in one call (the device path works in chunks of 2048
    scans) and p_miss = 0.5 (l_miss == 0: free cells must not move)."""
    rng = np.random.default_rng(5)
    bounds = (-6.4, 6.4, -6.4, 6.4)
    n = 2300
    origins = rng.uniform(-5, 5, size=(n, 2))
    clouds = [o + rng.normal(scale=2.5, size=(int(rng.integers(0, 6)), 2)) for o in origins]
    flat, off = synth.pack_ragged(clouds)
    for kw in (GKW, dict(GKW, p_miss=0.5), dict(GKW, p_miss=0.49999)):
        gpu, ref = make_pair(bounds, **kw)
        gpu._dev.update(origins, flat, off)
        ref.update_many(origins, flat, off, fast=True)
        assert_same(gpu, ref, str(kw))


def test_hit_field_overflow_is_reported_and_the_handle_survives():
    """More than 4095 endpoints of ONE scan in ONE cell exceed the packed hit counter: both device
    paths must refuse loudly (no silent wrap), and the grid must keep working afterwards."""
    bounds = (-3.2, 3.2, -3.2, 3.2)
    gpu, ref = make_pair(bounds, **GKW)
    pts = np.tile([[1.011, 0.512]], (5000, 1))
    with pytest.raises(RuntimeError, match="4095"):
        gpu.update_scan(np.zeros(2), pts)
    gpu.reset()
    ok = np.tile([[1.011, 0.512]], (4095, 1))
    gpu.update_scan(np.zeros(2), ok)
    ref.update_scan(np.zeros(2), ok, fast=True)
    assert_same(gpu, ref, "after the refused call")


def test_randomised_configurations_match_the_oracle():
    """Twelve random grids (sizes that are no multiple of the tile, origins inside and outside, endpoints far out,
    p_hit / p_miss on both sides of 0.5, lop-sided clamps that still contain 0, duplicate endpoints, empty scans):
    the device map must equal the oracle's bit for bit on both device paths."""
    rng = np.random.default_rng(20261018)
    for case in range(12):
        w, h = rng.uniform(3.0, 14.0, size=2)
        res = float(rng.choice([0.05, 0.07, 0.1]))
        bounds = (-w / 2, w / 2, -h / 3, 2 * h / 3)
        kw = dict(resolution=res, p_hit=float(rng.uniform(0.08, 0.95)), p_miss=float(rng.uniform(0.08, 0.95)),
                  log_odds_min=-float(rng.uniform(0.3, 9.0)), log_odds_max=float(rng.uniform(0.3, 9.0)))
        n_scans = int(rng.integers(1, 90))
        origins = rng.uniform(-0.7 * max(w, h), 0.7 * max(w, h), size=(n_scans, 2))
        clouds = []
        for o in origins:
            n = int(rng.integers(0, 120))
            pts = o + rng.normal(scale=rng.choice([0.2, 2.0, 30.0]), size=(n, 2))
            if n > 4:
                pts[1] = pts[0]                           # duplicate endpoint
                pts[2] = o                                # zero-length ray
            clouds.append(pts)
        flat, off = synth.pack_ragged(clouds)
        gpu, ref = make_pair(bounds, **kw)
        gpu._dev.update(origins, flat, off)
        ref.update_many(origins, flat, off, fast=True)
        assert_same(gpu, ref, f"case {case}: {kw} bounds {bounds} scans {n_scans}")
        # ... and scan by scan on top of the batch (state carried between calls)
        for s in range(min(n_scans, 5)):
            gpu.update_scan(origins[s], clouds[s])
            ref.update_scan(origins[s], clouds[s], fast=True)
        assert_same(gpu, ref, f"case {case} after single scans")


def test_device_resident_entry_point_is_stream_ordered_and_reports_late():
    """icpb200_grid_update_dev (device pointers, caller's stream): same map as the oracle, statistics available after
    the call, offsets that do not add up refused with nothing written, and the hit-field overflow -- detected on the
    device after the call has returned -- reported by the next call on the grid."""
    import torch
    dev = torch.device("cuda", 0)
    scans, poses = synth.make_sequence(120, world="campus", seed=3)
    hits = [synth.to_world_frame(s, p) for s, p in zip(scans, poses)]
    bounds = (poses[:, 0].mean() - 25.6, poses[:, 0].mean() + 25.6, poses[:, 1].mean() - 25.6, poses[:, 1].mean() + 25.6)
    gpu, ref = make_pair(bounds, **GKW)
    flat, off = synth.pack_ragged(hits)
    org = poses[:, :2].copy()
    d_org, d_hits, d_off = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (org, flat, off))
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()

    class _Dev:                                       # the shim caches its host mirror: drop it after every device update
        def update_dev(self, *a):
            gpu._invalidate()
            gpu._dev.update_dev(*a)
    dev_entry = _Dev()
    for rep in range(2):                              # the second call takes the speculative fill (buffers already sized)
        dev_entry.update_dev(len(off) - 1, d_org.data_ptr(), d_hits.data_ptr(), d_off.data_ptr(), int(off[-1]), stream.cuda_stream)
        cells = ref.update_many(org, flat, off, fast=True)
        assert gpu._dev.last_stats()["traversed"] == cells
        assert_same(gpu, ref, f"device-resident update {rep}")
    # offsets that do not add up: refused, map untouched
    bad = off.copy()
    bad[5] = bad[4] - 1
    d_bad = torch.from_numpy(bad).to(dev)
    with pytest.raises(RuntimeError, match="hit_off"):
        dev_entry.update_dev(len(off) - 1, d_org.data_ptr(), d_hits.data_ptr(), d_bad.data_ptr(), int(off[-1]), stream.cuda_stream)
    with pytest.raises(RuntimeError, match="hit_off"):
        dev_entry.update_dev(len(off) - 1, d_org.data_ptr(), d_hits.data_ptr(), d_off.data_ptr(), int(off[-1]) - 1, stream.cuda_stream)
    assert_same(gpu, ref, "after the refused calls")
    dev_entry.update_dev(len(off) - 1, d_org.data_ptr(), d_hits.data_ptr(), d_off.data_ptr(), int(off[-1]), stream.cuda_stream)
    ref.update_many(org, flat, off, fast=True)
    assert_same(gpu, ref, "the handle works after the refused calls")
    # a scan with more than 4095 rays: the fill pass needs its checking variant (decided on the device)
    big = np.column_stack([np.linspace(bounds[0] + 1, bounds[1] - 1, 6000), np.full(6000, bounds[2] + 2.0)])
    o1 = np.array([[poses[0, 0], poses[0, 1]]])
    boff = np.array([0, 6000], dtype=np.int64)
    d_o1, d_big, d_boff = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (o1, big, boff))
    dev_entry.update_dev(1, d_o1.data_ptr(), d_big.data_ptr(), d_boff.data_ptr(), 6000, stream.cuda_stream)
    ref.update_many(o1, big, boff, fast=True)
    assert_same(gpu, ref, "6000-ray scan")
    # 5000 endpoints of one scan in one cell: reported late (by the next call), then the handle works again
    pts = np.tile([[poses[0, 0] + 1.011, poses[0, 1] + 0.512]], (5000, 1))
    d_pts = torch.from_numpy(pts).to(dev)
    d_poff = torch.from_numpy(np.array([0, 5000], dtype=np.int64)).to(dev)
    try:
        dev_entry.update_dev(1, d_o1.data_ptr(), d_pts.data_ptr(), d_poff.data_ptr(), 5000, stream.cuda_stream)
        with pytest.raises(RuntimeError, match="4095"):
            gpu._dev.last_stats()
    except RuntimeError as e:                         # the ordered path reports it at once
        assert "4095" in str(e)
    gpu.reset()
    gpu.update_scan(o1[0], pts[:4095])
    ref2 = oo.GridOracleC(*bounds, **GKW)
    ref2.update_scan(o1[0], pts[:4095], fast=True)
    assert_same(gpu, ref2, "after the overflow report")


def test_rebuild_map_in_one_call_matches_the_reference_and_the_oracle():
    """`_rebuild_map` (slam.py:271-277) with `transform_points_2d` (slam.py:46-50) fused in --
    OccupancyGrid2D.rebuild(scan_history) / icpb200_grid_rebuild.  Bit-exact against the reference's own maps
    (tests/golden/rebuild.npz) and, on a larger random history, against the oracle's scan-by-scan rebuild."""
    from utilities import OccupancyGrid2D
    g = load_golden("rebuild.npz")
    off = g["scan_off"]
    grid = OccupancyGrid2D(*g["bounds"], **GKW)
    assert grid.log_odds.shape == tuple(g["grid_shape"])
    grid.update_scan(np.zeros(2), g["scans"][off[0]:off[1]])           # something to clear
    for variant in (0, 1, 0):
        history = [(g["scans"][off[s]:off[s + 1]], g[f"poses_{variant}"][s]) for s in range(len(off) - 1)]
        grid.rebuild(history)
        want = np.zeros(grid.log_odds.size, dtype=np.float32)
        want[g[f"nz_index_{variant}"]] = g[f"nz_value_{variant}"]
        assert grid.log_odds.ravel().tobytes() == want.tobytes(), variant
    grid.rebuild([])
    assert not grid.log_odds.any()
    # 150 scans of the campus world with perturbed poses
    scans, poses = synth.make_sequence(150, world="campus", seed=9)
    rng = np.random.default_rng(4)
    poses = poses + rng.normal(0, [0.1, 0.1, 0.02], size=poses.shape)
    mats = [np.array([[np.cos(t), -np.sin(t), x], [np.sin(t), np.cos(t), y], [0.0, 0.0, 1.0]]) for x, y, t in poses]
    bounds = (poses[:, 0].mean() - 51.2, poses[:, 0].mean() + 51.2, poses[:, 1].mean() - 51.2, poses[:, 1].mean() + 51.2)
    gpu, ref = make_pair(bounds, **GKW)
    history = list(zip(scans, mats))
    gpu.rebuild(history)
    oo.rebuild_map(ref, history)
    assert_same(gpu, ref, "150-scan rebuild")


def test_device_views_and_touched_tile_read_out():
    """mapping.py:150-160 on the device and the read-out of touched tiles only: log-odds bit-exact (through the shim's
    persistent mirror, across updates, a reset and a rebuild), probability / display within 1e-6 of numpy's float32
    evaluation, and far fewer tiles copied than the grid has."""
    scans, poses = synth.make_sequence(30, world="room", seed=9)
    hits = [synth.to_world_frame(s, p) for s, p in zip(scans, poses)]
    bounds = (-51.2, 51.2, -51.2, 51.2)                       # 2048 x 2048 cells, the room covers a corner of it
    gpu, ref = make_pair(bounds, **GKW)
    for k in range(10):
        gpu.update_scan(poses[k, :2], hits[k]); ref.update_scan(poses[k, :2], hits[k])
    assert_same(gpu, ref, "first read (mirror created)")
    tiles_total = ((gpu.nx + 63) // 64) * ((gpu.ny + 63) // 64)
    if os.environ.get("ICPB200_OCC_PATH") != "ordered":      # (the ordered replay does not track touched tiles: it reads everything)
        assert 0 < gpu._dev.last_tiles_copied < tiles_total // 2, gpu._dev.last_tiles_copied
    for k in range(10, 20):
        gpu.update_scan(poses[k, :2], hits[k]); ref.update_scan(poses[k, :2], hits[k])
    assert_same(gpu, ref, "second read (same mirror)")
    lo = ref.log_odds
    want_p = 1.0 / (1.0 + np.exp(-lo))                       # float32, as the reference computes it
    got_p = gpu.to_probability()
    assert got_p.dtype == np.float32 and got_p.shape == lo.shape and np.abs(got_p - want_p).max() < 1e-6
    want_d = 1.0 - want_p
    want_d[lo == 0.0] = 1.0
    want_d[lo < 0.0] = 0.85
    got_d = gpu.to_display()
    assert got_d.dtype == np.float32 and np.abs(got_d - want_d).max() < 1e-6
    assert np.array_equal(got_d == 1.0, want_d == 1.0) and np.array_equal(got_d == np.float32(0.85), want_d == np.float32(0.85))
    got_d[:] = 7.0                                            # a caller may scribble on what it got (the reference returns a new array)
    assert np.abs(gpu.to_display() - want_d).max() < 1e-6
    # whole-map read-out of the views equals the touched-tile one
    full = gpu._dev.read_view("probability")
    assert np.array_equal(full, gpu.to_probability())
    # reset: tiles that were touched must read as unexplored again, in every mirror
    gpu.reset(); ref.reset()
    gpu.update_scan(poses[25, :2] + 30.0, hits[25] + 30.0); ref.update_scan(poses[25, :2] + 30.0, hits[25] + 30.0)
    assert_same(gpu, ref, "after reset")
    assert np.abs(gpu.to_probability() - 1.0 / (1.0 + np.exp(-ref.log_odds))).max() < 1e-6
    # rebuild (slam.py:271-277) clears the touched set as well
    from oracle import occupancy_oracle
    mats = [np.array([[np.cos(t), -np.sin(t), x], [np.sin(t), np.cos(t), y], [0.0, 0.0, 1.0]]) for x, y, t in poses[:6]]
    history = [(scans[k], mats[k]) for k in range(6)]
    gpu.rebuild(history)
    occupancy_oracle.rebuild_map(ref, history)
    assert_same(gpu, ref, "after rebuild")
