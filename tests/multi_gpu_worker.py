"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), NCCL.  Checks, on real devices,
* occupancy sharded by bands + the tile push over NVLink peer stores: every rank ends with the oracle's map, bit for bit,
  also after a second round without a reset and on a grid whose bands do not divide by the ranks; the band gather
  (grid_gather_device) gives the same map;
* the pair batch partitioned over the ranks (icp_pairs_sharded on host buffers, DevicePairShard on device buffers) equals
  the single-rank call bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "iterative-closest-point-avmi_b200"), ROOT]
from icp_b200 import api, synth  # noqa: E402
from icp_b200 import dist as icpd  # noqa: E402
from oracle import occupancy_oracle  # noqa: E402


def main():
    rank, size, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api.init(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    try:
        gkw = dict(resolution=0.05, p_hit=0.85, p_miss=0.42, log_odds_min=-8.0, log_odds_max=8.0)
        scans, poses = synth.make_sequence(60, world="room", seed=2)
        hits = [synth.to_world_frame(s, p) for s, p in zip(scans, poses)]
        from utilities import OccupancyGrid2D
        for bounds in ((-25.6, 25.6, -25.6, 25.6), (-23.0, 24.1, -14.3, 15.05)):        # 1024 x 1024, and 942 x 587 (ragged bands)
            ref = occupancy_oracle.GridOracleC(*bounds, **gkw)
            grid = OccupancyGrid2D(*bounds, **gkw)
            grid._dev.set_shard(rank, size)
            icpd.grid_share_setup(grid._dev)
            twin = OccupancyGrid2D(*bounds, **gkw)                                        # same shards, reassembled by the band gather
            twin._dev.set_shard(rank, size)
            for lo, hi in ((0, 25), (25, 60)):                                            # two rounds, no reset in between
                flat, off = synth.pack_ragged(hits[lo:hi])
                org = poses[lo:hi, :2].copy()
                ref.update_many(org, flat, off)
                grid._dev.update(org, flat, off)
                icpd.grid_push_device(grid._dev, stream.cuda_stream)
                torch.cuda.synchronize()
                got = grid._dev.read()
                assert got.tobytes() == ref.log_odds.tobytes(), f"rank {rank}: push, scans {lo}:{hi}, {np.count_nonzero(got != ref.log_odds)} cells differ"
                mirror = np.zeros_like(got)
                grid._dev.read_dirty(mirror)                                              # the peers' tiles are marked touched here too
                assert mirror.tobytes() == ref.log_odds.tobytes(), f"rank {rank}: touched-tile read-out after a push"
                twin._dev.update(org, flat, off)
                icpd.grid_gather_device(twin._dev)
                assert twin._dev.read().tobytes() == ref.log_odds.tobytes(), f"rank {rank}: band gather, scans {lo}:{hi}"
            dist.barrier()
            grid.reset(); twin.reset()
            dist.barrier()
            flat, off = synth.pack_ragged(hits[:5])
            grid._dev.update(poses[:5, :2].copy(), flat, off)
            icpd.grid_push_device(grid._dev, stream.cuda_stream)
            torch.cuda.synchronize()
            ref.reset(); ref.update_many(poses[:5, :2].copy(), flat, off)
            assert grid._dev.read().tobytes() == ref.log_odds.tobytes(), f"rank {rank}: after a reset"
            dist.barrier()
            grid._dev.close(); twin._dev.close()
        # ---- pairs
        cfg = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, method="point_to_line", normal_k=12)
        flat, off = synth.pack_ragged(scans)
        # 350 pairs per rank: every call, the whole batch and each rank's share, takes the two-phase schedule (>= 2 pairs per
        # SM), so the comparison can be bit for bit; batches on either side of that threshold run different kernel variants
        # and agree to rounding only
        pairs = synth.loop_closure_pairs(poses, 350 * size, seed=3, max_dist=2.0).astype(np.int32)
        si, ti = pairs[:, 0].copy(), pairs[:, 1].copy()
        whole = api.icp_pairs(flat, off, si, ti, **cfg)
        sharded = icpd.icp_pairs_sharded(flat, off, si, ti, **cfg)
        shard = icpd.DevicePairShard(flat, off, si, ti, torch.device("cuda", local))
        shard.enqueue(stream, **cfg)
        resident = shard.results()
        for key in ("R", "t", "error", "prev_error", "iters", "status"):
            assert sharded[key].tobytes() == whole[key].tobytes(), (rank, "host buffers", key)
            assert resident[key].tobytes() == whole[key].tobytes(), (rank, "device buffers", key)
        dist.barrier()
        print(f"rank {rank} of {size}: ok", flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
