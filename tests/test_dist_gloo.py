"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: pair sharding +
all_gather, tile ownership + all_reduce reassembly.  The compute stand-ins are
the CPU oracle, so the gathered results are checked against unsharded runs."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _icp_oracle_pairs(points, cloud_off, src_idx, tgt_idx, *args, R_init=None, t_init=None, **kw):
    from oracle import icp_oracle
    n = len(src_idx)
    dim = points.shape[1]
    out = dict(R=np.zeros((n, dim, dim)), t=np.zeros((n, dim)), error=np.zeros(n), prev_error=np.zeros(n),
               iters=np.zeros(n, np.int32), status=np.zeros(n, np.int32))
    for p, (i, j) in enumerate(zip(src_idx, tgt_idx)):
        extra = dict(kw)
        if R_init is not None:
            extra.update(R_init=R_init[p], t_init=t_init[p])
        R, t, err, iters, status = icp_oracle.register(points[cloud_off[i]:cloud_off[i + 1]],
                                                       points[cloud_off[j]:cloud_off[j + 1]], *args, **extra)
        out["R"][p], out["t"][p], out["error"][p], out["iters"][p], out["status"][p] = R, t, err, iters, status
    return out


def _worker(rank, size, port, tmp):
    for p in (PKG, ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from icp_b200 import dist as icpd, synth
    from oracle import occupancy_oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=size)
    try:
        # ---- pairs: 7 pairs over 2 ranks (uneven split), with initial guesses
        scans, poses = synth.make_sequence(8, world="room", seed=1)
        flat, off = synth.pack_ragged(scans)
        si, ti = np.arange(7, dtype=np.int32), np.arange(1, 8, dtype=np.int32)
        R0 = np.tile(np.eye(2), (7, 1, 1))
        t0 = np.tile([0.01, -0.01], (7, 1))
        cfg = dict(error_threshold=1e-7, max_iterations=20, voxel_size=0.3, method="point_to_point")
        got = icpd.icp_pairs_sharded(flat, off, si, ti, compute=_icp_oracle_pairs, R_init=R0, t_init=t0, **cfg)
        plan = icpd.plan_pair_shards(si, ti, size)
        assert sorted(np.concatenate(plan).tolist()) == list(range(7)) and abs(len(plan[0]) - len(plan[1])) <= 1
        want = _icp_oracle_pairs(flat, off, si, ti, R_init=R0, t_init=t0, **cfg)
        for key in ("R", "t", "error", "iters", "status"):
            assert np.array_equal(got[key], want[key]), key
        # ---- occupancy: each rank keeps only its tiles of the replay; the sum is the whole map
        gkw = dict(resolution=0.05, p_hit=0.85, p_miss=0.42, log_odds_min=-8.0, log_odds_max=8.0)
        full = occupancy_oracle.GridOracleC(-12.8, 12.8, -9.6, 9.6, **gkw)      # 512 x 384 cells
        for s in range(6):
            full.update_scan(poses[s, :2] * 0.4, synth.to_world_frame(scans[s], poses[s]) * 0.4)
        mask = icpd.owned_tile_mask(full.nx, full.ny, rank, size)
        other = icpd.owned_tile_mask(full.nx, full.ny, 1 - rank, size)
        assert not (mask & other).any() and (mask | other).all()
        rows = icpd.owned_rows(full.ny, rank, size)                # bands of 64 rows dealt round-robin
        assert np.array_equal(rows, (np.arange(full.ny) // icpd.TILE) % 2 == rank) and mask.sum() == rows.sum() * full.nx
        partial = np.where(mask, full.log_odds, np.float32(0))
        total = icpd.grid_gather_host(partial)
        assert total.dtype == np.float32 and total.tobytes() == full.log_odds.tobytes()
        # a second round WITHOUT a reset: the rows a rank does not own now hold the previous map (not zeros); the
        # reassembly is a gather, so they are simply overwritten (a sum over ranks would double them)
        for s in range(6, 8):
            full.update_scan(poses[s, :2] * 0.4, synth.to_world_frame(scans[s], poses[s]) * 0.4)
        stale = total.copy()
        stale[rows] = full.log_odds[rows]                           # this rank updated only its own rows
        again = icpd.grid_gather_host(stale)
        assert again.tobytes() == full.log_odds.tobytes()
        # a grid whose bands do not divide by the ranks (5 bands over 2 ranks, the last one short)
        odd = np.arange(300 * 7, dtype=np.float32).reshape(300, 7)
        rows_odd = icpd.owned_rows(300, rank, size)
        part = np.where(rows_odd[:, None], odd, np.float32(-1))
        assert icpd.grid_gather_host(part).tobytes() == odd.tobytes()
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_shard_ranges_cover_everything():
    from icp_b200 import dist as icpd
    for n in (0, 1, 7, 8, 1999, 8192):
        for size in (1, 2, 3, 4, 8):
            blocks = [icpd.shard_range(n, r, size) for r in range(size)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            lens = [b - a for a, b in blocks]
            assert max(lens) - min(lens) <= 1


def test_pair_shard_plan_is_a_partition_with_locality():
    """Every pair lands on exactly one rank, shares differ by a few pairs, and a rank's pairs name a small part of the
    scan history (so the upload and the per-cloud kernels shard with the pairs)."""
    from icp_b200 import dist as icpd, synth
    poses = synth.room_trajectory(2000)
    pairs = synth.loop_closure_pairs(poses, 8192, seed=0, max_dist=3.0)
    si, ti = pairs[:, 0], pairs[:, 1]
    for size in (1, 2, 3, 8):
        plan = icpd.plan_pair_shards(si, ti, size)
        allp = np.concatenate(plan)
        assert len(allp) == 8192 and np.array_equal(np.sort(allp), np.arange(8192))
        lens = [len(p) for p in plan]
        assert max(lens) - min(lens) <= 4 * size
        if size == 8:
            for p in plan:
                used = np.unique(np.concatenate([si[p], ti[p]]))
                assert len(used) < 0.45 * 2000, len(used)


def test_result_block_round_trip():
    from icp_b200 import dist as icpd
    rng = np.random.default_rng(0)
    n, dim, size = 37, 2, 3
    si = rng.integers(0, 50, n); ti = rng.integers(0, 50, n)
    plan = icpd.plan_pair_shards(si, ti, size)
    want = dict(R=rng.normal(size=(n, 2, 2)), t=rng.normal(size=(n, 2)), error=rng.normal(size=n), prev_error=rng.normal(size=n),
                iters=rng.integers(0, 150, n).astype(np.int32), status=rng.integers(0, 3, n).astype(np.int32))
    layout, nbytes = icpd.result_layout(max(len(p) for p in plan), dim)
    blocks = np.zeros((size, nbytes), dtype=np.uint8)
    for r, mine in enumerate(plan):
        for name, (at, dt, width) in layout.items():
            vals = np.ascontiguousarray(want[name][mine], dtype=dt).reshape(-1)
            blocks[r, at:at + vals.nbytes] = vals.view(np.uint8)
    got = icpd._unpack_results(blocks, plan, dim)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
