#!/usr/bin/env python
"""Where does the multi-GPU step lose time?  (bench.py at 8 ranks: 2.48 ms per C2 batch against 1.76 ms on one GPU.)

Run under torchrun with N ranks.  Every rank registers the same C2 batch (as bench.py does) under a list of
variants and rank 0 prints, per variant, the step time (CUDA events, max over ranks), the library's own kernel
times (K1 / K2 / K3 events) and the host time of enqueueing one step:

    sampler   none | smi_spawn (bench.py up to 2dafed2: one nvidia-smi process per 0.2 s, started with the timed
              region) | nvml (in-process NVML thread)
    gather    none | async (all_gather of step k on a side stream under step k + 1) | sync (same stream, inside the step)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        profiles/scale_probe.py --steps 20
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (also puts the package on sys.path)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--scans", type=int, default=2000)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from icp_b200 import _lib, api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    api.init(local_rank)
    lib = _lib.load()
    dev = torch.device("cuda", local_rank)
    scans, poses, flat, off, si, ti = bench.build_c2(args.scans, seed=0)
    n_pairs = len(si)
    max_pts = int(np.max(np.diff(off)))
    cfg = bench.ICP_CFG
    d_pts, d_off = torch.from_numpy(flat).to(dev), torch.from_numpy(off).to(dev)
    d_si, d_ti = torch.from_numpy(si).to(dev), torch.from_numpy(ti).to(dev)
    d_R = torch.empty((n_pairs, 2, 2), dtype=torch.float64, device=dev)
    d_t = torch.empty((n_pairs, 2), dtype=torch.float64, device=dev)
    d_err = torch.empty(n_pairs, dtype=torch.float64, device=dev)
    d_prev = torch.empty(n_pairs, dtype=torch.float64, device=dev)
    d_it = torch.empty(n_pairs, dtype=torch.int32, device=dev)
    d_st = torch.empty(n_pairs, dtype=torch.int32, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    gather_buf = [[torch.empty((n_pairs, 3), dtype=torch.float64, device=dev) for _ in range(world)] for _ in range(2)]
    mine_buf = [torch.empty((n_pairs, 3), dtype=torch.float64, device=dev) for _ in range(2)]
    stream = torch.cuda.Stream(device=dev)
    comm_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    pending = [None, None]
    step_no = [0]

    def compute():
        rc = lib.icpb200_icp_pairs_dev(
            len(off) - 1, 2, d_pts.data_ptr(), d_off.data_ptr(), max_pts, n_pairs, d_si.data_ptr(), d_ti.data_ptr(),
            None, None, cfg["error_threshold"], cfg["max_iterations"], cfg["voxel_size"],
            _lib.POINT_TO_LINE, cfg["normal_k"], -1.0, _lib.NN_AUTO,
            d_R.data_ptr(), d_t.data_ptr(), d_err.data_ptr(), d_prev.data_ptr(), d_it.data_ptr(), d_st.data_ptr(),
            stream.cuda_stream)
        _lib.check(rc, "icpb200_icp_pairs_dev")

    def step(gather):
        compute()
        if world == 1 or gather == "none":
            return
        k = step_no[0] & 1
        step_no[0] += 1
        if pending[k] is not None:
            pending[k].wait()
        mine = mine_buf[k]
        torch.atan2(d_R[:, 1, 0], d_R[:, 0, 0], out=mine[:, 0])
        mine[:, 1:].copy_(d_t)
        if gather == "sync":
            dist.all_gather(gather_buf[k], mine)
        else:
            comm_stream.wait_stream(stream)
            with torch.cuda.stream(comm_stream):
                pending[k] = dist.all_gather(gather_buf[k], mine, async_op=True)

    def drain():
        for k in (0, 1):
            if pending[k] is not None:
                pending[k].wait()
                pending[k] = None
        stream.wait_stream(comm_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class NoSampler:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            pass

        def summary(self):
            return {}

    def sampler(kind):
        if kind == "smi_spawn":
            return bench.SmiSpawnSampler(local_rank)
        if kind == "nvml":
            return bench.ClockSampler(local_rank)
        return NoSampler()

    variants = [("none", "none"), ("none", "async"), ("none", "sync"), ("nvml", "async"), ("smi_spawn", "async"),
                ("nvml", "none"), ("none", "none")]
    for samp, gather in variants:
        for _ in range(3):
            step(gather)
        drain()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        host = []
        with sampler(samp) as clk:
            barrier()
            for a, b in ev:
                flush.zero_()
                a.record(stream)
                t0 = time.perf_counter()
                step(gather)
                host.append(time.perf_counter() - t0)
                b.record(stream)
            ta, tb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ta.record(stream)
            drain()
            tb.record(stream)
            barrier()
        ms = np.array([a.elapsed_time(b) for a, b in ev])
        tail = ta.elapsed_time(tb)
        ks = api.icp_last_stats()
        kern = (ks["voxel_kernel_ns"] + ks["normals_kernel_ns"] + ks["pair_kernel_ns"]) / 1e6
        row = torch.tensor([ms.mean(), np.median(ms), ms.max(), tail, kern, ks["pair_kernel_ns"] / 1e6,
                            1e3 * float(np.mean(host))], dtype=torch.float64, device=dev)
        rows = [torch.empty_like(row) for _ in range(world)]
        if world > 1:
            dist.all_gather(rows, row)
        else:
            rows = [row]
        if rank == 0:
            r = torch.stack(rows).cpu().numpy()
            print(f"sampler={samp:9s} gather={gather:5s} | step mean max-over-ranks {r[:, 0].max():.3f} ms "
                  f"(min rank {r[:, 0].min():.3f}) median {r[:, 1].max():.3f} worst step {r[:, 2].max():.3f} "
                  f"tail {r[:, 3].max():.3f} | kernels K1+K2+K3 {r[:, 4].max():.3f} (K3 {r[:, 5].min():.3f}..{r[:, 5].max():.3f}) "
                  f"| host enqueue {r[:, 6].max():.3f} ms | clocks {clk.summary()}", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
