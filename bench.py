#!/usr/bin/env python
"""Benchmark of the B200-native ICP registration + occupancy raycast hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One GPU (the BENCH line): the headline workload is BASELINE.json configs[1] ("C2"): 2000 synthetic 1080-beam 2-D
scans -> 1999 consecutive scan-to-scan point_to_line registrations with the reference's config.yaml parameters.  One
"step" = one pass of the hot path over that whole batch.  The line also carries `c5` (configs[4], the 8192-pair
loop-closure batch on this one GPU: the N = 1 point of the strong-scaling curve), `occupancy` (configs[3], C4),
`online` (single calls through the drop-in shim), `verify` (the timed outputs against the oracle) and `extras`
(C1 teapot, C3 scan -> submap, rotation search, map rebuild).

Several GPUs (torchrun, the SCALE lines): the headline is configs[4] ("C5"): ONE batch of 8192 loop-closure
candidate pairs partitioned across the ranks (strong scaling; icp_b200.dist: locality-sorted chunks dealt round-robin,
every rank uploads and preprocesses only the clouds its pairs reference, one NCCL all_gather of packed result blocks).
`c2_weak` keeps round 1's weak-scaling figure (every rank registers the whole C2 batch) as an extra.

Keys of a line (rank 0 prints ONE JSON line):
  value      registrations/s with the clouds already resident in HBM, CUDA events on the launching stream, max over ranks
  e2e        the same batch through the host-buffer call (icpb200_icp_pairs / dist.icp_pairs_sharded): H2D of the clouds
             and D2H of the poses inside the timed region; `unpinned` = the same from pageable memory
  roofline   FP32-FMA pipe: EXECUTED fp32 pair evaluations (a device counter) x 5 flop / kernel time against
             SMs x 128 lanes x clock / 4; `algorithmic_equivalent` (the brute-force count of BASELINE.md section 4,
             which the kernel does not execute) and `latency_model` are reported next to it, clearly named
  cpu_baseline  the reference's CPU path timed on this box's host cores on a bounded sample of the same pairs:
             the UNMODIFIED reference from oracle/_ref (kind "reference") when the archive is present, else the
             oracle port (kind "port")

--impl reference times that CPU implementation alone on all host cores.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "iterative-closest-point-avmi_b200")
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

ICP_CFG = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04,
               method="point_to_line", normal_k=12)          # /root/reference/config.yaml:19-24
GRID_CFG = dict(resolution=0.05, p_hit=0.85, p_miss=0.42, log_odds_min=-8.0, log_odds_max=8.0)  # config.yaml:85-90
GRID_BOUNDS = (-102.4, 102.4, -102.4, 102.4)                  # 4096 x 4096 cells
FMA_LANES_PER_SM = 128
FLOP_PER_PAIR_EVAL_2D = 5
FMA_INSTR_PER_PAIR_EVAL_2D = 4
_REAL_STDOUT = None


# --------------------------------------------------------------------------- helpers
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


class _SamplerBase:
    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except ValueError:
                continue
            for name, flag in zip(names, r[3:7]):
                if str(flag).lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx or None,
                    reasons=sorted(reasons), samples=len(sm), source=self.SOURCE)


class SmiSpawnSampler(_SamplerBase):
    """One nvidia-smi process per sample (0.2 s apart).  Kept as the fallback when NVML cannot be loaded and for
    profiles/scale_probe.py: a process that attaches to every GPU of the box while the timed region runs is not free."""
    SOURCE = "nvidia-smi"
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)


class NvmlSampler(_SamplerBase):
    """Samples SM clock, power and the clock-event (throttle) reasons of this rank's GPU during the timed region.

    The same NVML counters nvidia-smi prints (clocks.sm, clocks.max.sm, power.draw, clocks_event_reasons.*), read
    in-process every 5 ms by a thread that was attached to the device BEFORE the timed region starts: the region is
    a few milliseconds long, so a `nvidia-smi -lms 200` loop would rarely land a sample inside it, and starting one
    nvidia-smi process per rank at the start of the region (what this class did before) attaches to every GPU of the
    box while the kernels are being launched."""
    SOURCE = "nvml"
    PERIOD_S = 0.005

    def __init__(self, index):
        import pynvml as nv
        self.nv = nv
        nv.nvmlInit()
        self.h = None
        try:                                        # CUDA index -> NVML handle through the UUID (CUDA_VISIBLE_DEVICES)
            import torch
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            self.h = nv.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        except Exception:
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
        self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv, h = self.nv, self.h
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = (0x8, 0x40, 0x20, 0x4)               # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mask = int(get_reasons(h))
                try:
                    power = nv.nvmlDeviceGetPowerUsage(h) / 1e3
                except Exception:
                    power = float("nan")
                self.rows.append([sm, self.max_sm, power] + ["Active" if mask & b else "Not Active" for b in bits])
            except Exception:
                pass
            self._stop.wait(self.PERIOD_S)


def ClockSampler(index):
    """NVML in-process when it loads, one nvidia-smi process per sample otherwise."""
    try:
        return NvmlSampler(index)
    except Exception:
        return SmiSpawnSampler(index)


def voxel_count(cloud, voxel):
    """Occupied voxels of a cloud (sizes the roofline's pair-evaluation count)."""
    cell = np.floor((cloud - cloud.min(axis=0)) / voxel).astype(np.int64)
    return len(np.unique(cell[:, 0] * (cell[:, 1].max() + 1) + cell[:, 1]))


def build_c2(n_scans, seed):
    from icp_b200 import synth
    scans, poses = synth.make_sequence(n_scans, world="room", seed=seed)
    flat, off = synth.pack_ragged(scans)
    idx = np.arange(n_scans - 1, dtype=np.int32)
    return scans, poses, flat, off, idx, idx + 1


def build_c5(n_scans, n_pairs, seed):
    from icp_b200 import synth
    scans, poses = synth.make_sequence(n_scans, world="room", seed=seed)
    flat, off = synth.pack_ragged(scans)
    pairs = synth.loop_closure_pairs(poses, n_pairs, seed=seed, max_dist=3.0).astype(np.int32)
    return scans, poses, flat, off, pairs[:, 0].copy(), pairs[:, 1].copy()


def build_c4(n_scans, seed):
    from icp_b200 import synth
    scans, poses = synth.make_sequence(n_scans, world="campus", seed=seed)
    hits = [synth.to_world_frame(s, p) for s, p in zip(scans, poses)]
    flat, off = synth.pack_ragged(hits)
    return poses[:, :2].copy(), flat, off


# --------------------------------------------------------------------------- CPU baseline (reference or oracle port)
def _reference_kind():
    """("reference", description) when oracle/_ref holds the packed, unmodified reference, else ("port", ...)."""
    from oracle import ref_loader
    if ref_loader.reference_root() is not None:
        return "reference", "the UNMODIFIED reference utilities/icp.py::ICP from oracle/_ref (numpy + scipy KDTree)"
    return "port", "oracle/icp_oracle.py (numpy + scipy KDTree restatement, same calls as the reference)"


_REF_ICP = None


def _cpu_icp_one(args):
    """One registration on one core: (seconds, R, t, error, iters or -1, status or -1)."""
    global _REF_ICP
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    src, tgt, kind = args
    if kind == "reference":
        if _REF_ICP is None:
            from oracle import ref_loader
            _REF_ICP = ref_loader.import_reference("utilities.icp").ICP
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            t0 = time.perf_counter()
            R, t, err = _REF_ICP(src, tgt, ICP_CFG["error_threshold"], ICP_CFG["max_iterations"], ICP_CFG["voxel_size"],
                                 method=ICP_CFG["method"], normal_k=ICP_CFG["normal_k"])
            dt = time.perf_counter() - t0
        return dt, R, t, float(err), -1, -1
    from oracle import icp_oracle
    t0 = time.perf_counter()
    R, t, err, iters, status = icp_oracle.register(src, tgt, **ICP_CFG)
    return time.perf_counter() - t0, R, t, float(err), int(iters), int(status)


def cpu_icp_run(scans, src_idx, tgt_idx, sel, cores, kind):
    """Register the pairs `sel` on `cores` processes; returns (wall seconds of the map alone, list of results)."""
    jobs = [(scans[src_idx[i]], scans[tgt_idx[i]], kind) for i in sel]
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"     # inherited by the workers
    if cores > 1:
        # spawn: the parent may hold a CUDA context, which must not be forked
        with mp.get_context("spawn").Pool(cores) as pool:
            pool.map(_cpu_icp_one, jobs[:cores])                 # start-up + imports, untimed
            t0 = time.perf_counter()
            res = pool.map(_cpu_icp_one, jobs, chunksize=1)
            wall = time.perf_counter() - t0
    else:
        t0 = time.perf_counter()
        res = [_cpu_icp_one(j) for j in jobs]
        wall = time.perf_counter() - t0
    return wall, res


def cpu_icp_baseline(scans, src_idx, tgt_idx, n_sample, cores):
    """The reference's CPU registration timed on `cores` processes over n_sample pairs spread over the batch."""
    kind, what = _reference_kind()
    sel = np.linspace(0, len(src_idx) - 1, n_sample).astype(int)
    wall, res = cpu_icp_run(scans, src_idx, tgt_idx, sel, cores, kind)
    per_call = float(np.mean([r[0] for r in res]))
    return dict(value=n_sample / wall, unit="registrations/s", cores=cores, kind=kind,
                sample=f"{n_sample} of {len(src_idx)} pairs, {what}, {cores} processes x 1 thread; "
                       f"mean {per_call * 1e3:.1f} ms/registration/core",
                one_core_value=1.0 / per_call)


def cpu_raycast_baseline(origins, flat, off, n_sample):
    from oracle import occupancy_oracle
    g = occupancy_oracle.GridOracleC(*GRID_BOUNDS, **GRID_CFG)
    sel_off = off[:n_sample + 1]
    t0 = time.perf_counter()
    g.update_many(origins[:n_sample], flat[:sel_off[-1]], sel_off, fast=False)
    wall = time.perf_counter() - t0
    return dict(value=float(sel_off[-1]) / wall, unit="rays/s", cores=1, kind="port",
                sample=f"first {n_sample} of {len(off) - 1} scans, oracle/occupancy_oracle.c (plain C restatement, "
                       f"whole-grid clip per scan as the reference), 1 thread; the reference's pure-Python "
                       f"update_scan measured 2.8k rays/s/core in BASELINE.md")


# --------------------------------------------------------------------------- verification of the timed outputs
def verify_icp(scans, si, ti, res, n_spread, max_slow, cores):
    """Compare the registrations the bench timed with the oracle port (bit-pinned against the live reference,
    oracle/pin_against_reference.py) on a sample: n_spread pairs spread over the batch plus up to max_slow pairs that
    hit the iteration limit.  Poses within north_star's 1e-4 m / 1e-5 rad, iteration counts and exit status equal."""
    n = len(si)
    sel = set(np.linspace(0, n - 1, min(n_spread, n)).astype(int).tolist())
    slow = np.flatnonzero(res["status"] == 1)
    sel |= set(slow[:max_slow].tolist())
    sel = np.array(sorted(sel))
    wall, ref = cpu_icp_run(scans, si, ti, sel, cores, "port")
    dt = dr = de = 0.0
    it_bad = st_bad = 0
    diverged = 0
    for k, i in enumerate(sel):
        _, R, t, err, iters, status = ref[k]
        if np.isfinite(err) and err > 1e6:      # a registration that diverges in the reference itself: chaotic
            diverged += 1
            continue
        dt = max(dt, float(np.max(np.abs(res["t"][i] - t))))
        rel = res["R"][i] @ R.T
        dr = max(dr, abs(float(np.arctan2(rel[1, 0], rel[0, 0]))))
        if np.isfinite(err) or np.isfinite(res["error"][i]):
            de = max(de, abs(float(res["error"][i]) - err))
        it_bad += int(res["iters"][i] != iters)
        st_bad += int(res["status"][i] != status)
    ok = dt < 1e-4 and dr < 1e-5 and st_bad == 0
    out = dict(pairs_checked=int(len(sel)), iteration_limit_pairs_checked=int(min(len(slow), max_slow)),
               max_translation_diff_m=dt, max_rotation_diff_rad=dr, max_error_diff=de,
               iteration_count_mismatches=it_bad, status_mismatches=st_bad, diverged_in_the_reference_skipped=diverged, ok=bool(ok),
               checker="oracle/icp_oracle.py on the host cores, outside every timed region", seconds=wall)
    if not ok:
        raise SystemExit(f"bench.py: timed ICP outputs differ from the oracle: {out}")
    return out


# dram__bytes_read.sum + dram__bytes_write.sum per launch come from profiles/r02_ncu_traffic.json, written by
# profiles/ncu_traffic.py from an `ncu --set full` capture of this command and keyed by a hash of the kernel sources: a
# figure measured on other sources is reported as null, not silently reused.
def source_hash():
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(PKG, "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel):
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if not os.path.exists(path):
        return None, "no ncu capture committed for this round yet"
    with open(path) as f:
        d = json.load(f)
    if d.get("source_hash") != source_hash():
        return None, f"profiles/r02_ncu_traffic.json was captured on other kernel sources ({d.get('source_hash')}): stale, not reported"
    v = d.get("dram_bytes_per_launch", {}).get(kernel)
    return v, f"ncu --set full, profiles/r02_ncu_traffic.json ({d.get('command', '')})"


# --------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind, what = _reference_kind()
    if args.gpus > 1:
        scans, poses, flat, off, si, ti = build_c5(args.scans, args.pairs, seed=0)
        wl = f"C5 loop-closure candidate batch ({len(si)} pairs; bounded sample per step)"
    else:
        scans, poses, flat, off, si, ti = build_c2(args.scans, seed=0)
        wl = "C2 scan-to-scan point_to_line ICP, 1080-beam 2-D scans (bounded sample per step)"
    n_sample = min(len(si), max(cores * 4, 128))
    for _ in range(max(args.warmup, 0)):
        cpu_icp_baseline(scans, si, ti, min(cores, 8), cores)
    t_all, last = [], None
    for _ in range(args.steps):
        last = cpu_icp_baseline(scans, si, ti, n_sample, cores)
        t_all.append(n_sample / last["value"])           # the pool's map() alone: worker start-up and imports are not timed
    value = n_sample * args.steps / sum(t_all)
    line = dict(metric="icp_registrations_per_s", value=value, unit="registrations/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * float(np.mean(t_all)),
                higher_is_better=True, scaling="strong" if args.gpus > 1 else "weak", vs_baseline=None, dtype="f64",
                data="synthetic", impl="reference",
                config=dict(workload=wl, pairs_per_step=n_sample, **{k: v for k, v in ICP_CFG.items()}),
                cpu_baseline=dict(value=value, unit="registrations/s", cores=cores, kind=kind, sample=last["sample"]),
                e2e=dict(value=value, unit="registrations/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


# --------------------------------------------------------------------------- our arm
class Env:
    """What every leg needs: rank / world, device, streams, the library."""

    def __init__(self, args, rank, world, local_rank):
        import torch
        from icp_b200 import _lib, api
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        self.torch, self.api, self.lib = torch, api, _lib.load()
        self.dev = torch.device("cuda", local_rank)
        self.pk = peaks()
        self.stream = torch.cuda.Stream(device=self.dev)          # explicit stream: the library launches on it too
        torch.cuda.set_stream(self.stream)
        self.flush = torch.empty(512 << 20, dtype=torch.uint8, device=self.dev)        # > 126 MB L2
        self.sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return list(vals)
        import torch.distributed as dist
        red = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        return [float(v) for v in red]


def bench_icp(env, name, data, mode, steps, warmup, want_cpu, want_unpinned, verify):
    """One ICP workload.  mode: "single" (one GPU), "strong" (one batch partitioned over the ranks) or "weak"
    (every rank registers the whole batch).  Returns the result object (rank 0) or None."""
    import torch
    from icp_b200 import dist as icpd
    api = env.api
    scans, poses, flat, off, si, ti = data
    n_pairs = len(si)
    shard = icpd.DevicePairShard(flat, off, si, ti, env.dev, replicated=(mode != "strong"))
    kw = {k: v for k, v in ICP_CFG.items()}

    def step():
        shard.enqueue(env.stream, **kw, gather=env.world > 1 or mode == "single")

    for _ in range(max(warmup, 3)):
        step()
    env.barrier()
    launches0 = api.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with ClockSampler(env.local_rank) as clk:
        env.barrier()
        wall0 = time.perf_counter()
        for a, b in ev:
            env.flush.zero_()                   # evict L2 between timed iterations (outside the events)
            a.record(env.stream)
            step()
            b.record(env.stream)
        env.barrier()
        wall = time.perf_counter() - wall0
    launches = api.launch_count() - launches0
    kstats = api.icp_last_stats()
    phase = api.icp_phase_profile()
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    t_local = sum(ms_steps) / 1e3
    clocks = clk.summary()
    res = shard.results()                       # every pair of the batch, the caller's order (all ranks hold it)

    # ---- e2e: host buffers through the public call (H2D from page-locked memory + kernels + D2H per step)
    def e2e_call():
        if mode == "strong" and env.world > 1:
            return icpd.icp_pairs_sharded(flat, off, si, ti, **kw)
        out = api.icp_pairs(flat, off, si, ti, **kw)
        if mode == "weak" and env.world > 1:       # the path's only exchange: every rank's poses to every rank
            icpd.gather_result_blocks(out, 2)
        return out

    def e2e_time(reps):
        ts = []
        for k in range(2 + reps):
            env.barrier()
            t0 = time.perf_counter()
            out = e2e_call()
            if k >= 2:
                ts.append(time.perf_counter() - t0)
        return float(np.mean(ts)), out

    e2e_unpinned = None
    if want_unpinned:
        e2e_unpinned, _ = e2e_time(max(2, steps // 2))
    pin = api.pinned(flat, off, si, ti)
    e2e_s, out = e2e_time(steps)
    pin.release()
    assert np.array_equal(out["iters"], res["iters"]) and np.array_equal(out["status"], res["status"]), \
        "host-buffer and device-resident paths disagree"
    h2d = shard.h2d_bytes if mode == "strong" else flat.nbytes + off.nbytes + si.nbytes + ti.nbytes
    d2h = sum(out[k].nbytes for k in ("R", "t", "error", "prev_error", "iters", "status"))
    t_max, e2e_max, unp_max = env.max_over_ranks(t_local, e2e_s, e2e_unpinned or 0.0)
    checked = None
    if verify and env.rank == 0:
        checked = verify_icp(scans, si, ti, res, *verify, os.cpu_count() or 1)
    if env.rank != 0:
        return None

    # ---- roofline of the per-pair kernel K3 (this rank's share, last timed step)
    iters, status = res["iters"].astype(np.int64), res["status"]
    n_ds = np.array([voxel_count(s, ICP_CFG["voxel_size"]) for s in scans], dtype=np.int64)
    mine = shard.plan[env.rank] if mode == "strong" else np.arange(n_pairs)
    ns, nt = n_ds[si[mine]], n_ds[ti[mine]]
    alg_evals = float(np.sum(iters[mine] * ns * nt + nt * nt))     # brute force, + Mt^2 for the normals
    exe_evals = float(kstats["sweep_pair_evals"])                   # fp32 evaluations the kernel really issued (device counter)
    kernel_s = max(kstats["pair_kernel_ns"], 1) / 1e9               # K3 alone, CUDA events on its stream
    clk_mhz = clocks["sm_mhz"] or env.pk["sm_max_mhz"]
    peak_evals = env.sm_count * FMA_LANES_PER_SM * clk_mhz * 1e6 / FMA_INSTR_PER_PAIR_EVAL_2D
    peak_tf = peak_evals * FLOP_PER_PAIR_EVAL_2D / 1e12
    executed_tf = exe_evals * FLOP_PER_PAIR_EVAL_2D / kernel_s / 1e12
    traffic, traffic_src = ncu_traffic("icp_pairs_kernel")
    # latency model: a registration is a chain of dependent iterations; the kernel cannot finish before the busiest
    # CTA slot has run its share of them, nor before the longest single chain has run
    it_cycles = phase["cycles_per_iteration"]
    slots = env.sm_count * 2
    it_s = it_cycles / (clk_mhz * 1e6) if it_cycles else 0.0
    lb_throughput = float(iters[mine].sum()) * it_s / slots
    lb_chain = float(iters[mine].max()) * it_s if len(mine) else 0.0
    roofline = dict(
        bound="fp32_fma", achieved=executed_tf, peak=peak_tf, unit="TFLOP/s", frac=executed_tf / peak_tf,
        traffic=traffic, traffic_source=traffic_src, kernel="icp_pairs_kernel<2> (bulk + hand-over launches)",
        kernel_ms=kernel_s * 1e3, kernel_share_of_step=kernel_s * 1e3 / float(np.mean(ms_steps)),
        voxel_kernel_ms=kstats["voxel_kernel_ns"] / 1e6, normals_kernel_ms=kstats["normals_kernel_ns"] / 1e6,
        executed_pair_evals_per_launch=exe_evals, points_swept=kstats["points_swept"],
        points_carried=kstats["points_carried"], fp64_rescans=kstats["fp64_rescans"],
        note="achieved = fp32 pair evaluations the kernel EXECUTED (device counter) x 5 flop / kernel time: exact carry-over "
             "and the voxel-ordered slab sweep leave a few percent of the brute-force evaluations, so the FMA pipe is not "
             "what bounds this kernel -- the latency of its dependent per-iteration phases is (latency_model)",
        algorithmic_equivalent=dict(
            pair_evals_per_launch=alg_evals, tflops=alg_evals * FLOP_PER_PAIR_EVAL_2D / kernel_s / 1e12,
            times_fma_ceiling=alg_evals * FLOP_PER_PAIR_EVAL_2D / kernel_s / 1e12 / peak_tf,
            note="brute-force count of BASELINE.md section 4 (sum iters*Ns*Mt + Mt^2) the kernel would have to execute "
                 "without pruning; NOT a roofline fraction (the work is not executed)"),
        latency_model=dict(
            cycles_per_converged_iteration=it_cycles, phase_cycles=phase["phases"], iterations=int(iters[mine].sum()),
            longest_chain_iterations=int(iters[mine].max()) if len(mine) else 0, cta_slots=slots,
            lower_bound_ms=1e3 * max(lb_throughput, lb_chain), frac_of_kernel_time=max(lb_throughput, lb_chain) / kernel_s,
            note="lower bound = max(sum of iterations x measured converged-iteration time / CTA slots, longest chain x "
                 "that time); early iterations (full sweeps) cost more, so the fraction stays below 1"),
        peak_basis=f"{env.sm_count} SMs x {FMA_LANES_PER_SM} FP32 lanes x {clk_mhz:.0f} MHz (median SM clock sampled "
                   f"during the timed region) / {FMA_INSTR_PER_PAIR_EVAL_2D} FMA-pipe instr per 2-D pair evaluation x "
                   f"{FLOP_PER_PAIR_EVAL_2D} flop (BASELINE.md section 4)")
    cores = os.cpu_count() or 1
    cpu = cpu_icp_baseline(scans, si, ti, min(n_pairs, max(4 * cores, 256)), cores) if want_cpu else None
    total = n_pairs * (env.world if mode == "weak" else 1)
    sharding = {"single": "one GPU",
                "strong": "ONE batch partitioned over the ranks (strong scaling): pairs sorted by their lower cloud index, cut "
                          "into 4 chunks per rank dealt round-robin; a rank uploads / preprocesses only the clouds its pairs "
                          "reference; no data-path collective; one NCCL all_gather of packed result blocks inside the timed region",
                "weak": "every rank registers the whole batch (weak scaling, per-GPU work fixed); one NCCL all_gather of "
                        "result blocks per step inside the timed region"}[mode]
    return dict(metric="icp_registrations_per_s", value=total * steps / t_max, unit="registrations/s",
                n_gpus=env.world, steps=steps, warmup=max(warmup, 3), ms_per_step=1e3 * t_max / steps,
                higher_is_better=True, scaling="weak" if mode == "weak" else "strong", vs_baseline=None, dtype="f64",
                data="synthetic",
                config=dict(workload=name, pairs=n_pairs, pairs_this_rank=int(len(mine)), clouds_this_rank=int(shard.n_clouds),
                            points_per_cloud_after_voxel=float(n_ds.mean()), mean_iterations=float(iters.mean()),
                            converged=int((status == 0).sum()), max_iter_pairs=int((status == 1).sum()),
                            l2="flushed between timed steps (512 MiB memset)", sharding=sharding, **ICP_CFG),
                e2e=dict(value=total / e2e_max, unit="registrations/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                         api="icp_b200.dist.icp_pairs_sharded -> icpb200_icp_pairs (host buffers, blocking)" if mode == "strong"
                             else "icpb200_icp_pairs (host buffers, blocking)",
                         host_buffers="page-locked by the caller once (icpb200_pin_host), outside the timed region",
                         unpinned=dict(value=total / unp_max, unit="registrations/s",
                                       host_buffers="pageable numpy arrays") if want_unpinned else None),
                gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu, clocks=clocks, verify=checked,
                wall_s_timed_region=wall, peaks_source=env.pk["source"])


def bench_online(env, scans):
    """the slam.py main loop calls ICP() and update_scan() one at a time (slam.py:471-483, 557)."""
    import contextlib
    import io
    from utilities import ICP, OccupancyGrid2D
    from icp_b200 import synth
    n = min(len(scans) - 1, 200)
    lat = []
    with contextlib.redirect_stdout(io.StringIO()):
        for k in range(5):
            ICP(scans[k], scans[k + 1], **ICP_CFG)
        for k in range(n):
            t0 = time.perf_counter()
            ICP(scans[k], scans[k + 1], **ICP_CFG)
            lat.append(time.perf_counter() - t0)
    c4_scans, c4_poses = synth.make_sequence(120, world="campus", seed=0)
    grid = OccupancyGrid2D(*GRID_BOUNDS, **GRID_CFG)
    up = []
    for k in range(120):
        hits = synth.to_world_frame(c4_scans[k], c4_poses[k])
        t0 = time.perf_counter()
        grid.update_scan(c4_poses[k, :2], hits)
        if k >= 20:
            up.append(time.perf_counter() - t0)
    grid._dev.close()
    lat, up = np.array(lat) * 1e3, np.array(up) * 1e3
    return dict(icp_call_ms=dict(median=float(np.median(lat)), p90=float(np.percentile(lat, 90)), calls=int(n)),
                icp_calls_per_s=float(1e3 / np.mean(lat)),
                update_scan_ms=dict(median=float(np.median(up)), p90=float(np.percentile(up, 90)), calls=int(len(up))),
                note="sequential single calls through the drop-in shim (utilities.ICP / OccupancyGrid2D.update_scan), "
                     "host numpy arrays in and out, wall clock per call")


def run_ours(args, rank, world, local_rank):
    import torch
    from icp_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libicp_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    api.init(local_rank)
    env = Env(args, rank, world, local_rank)
    cpu_ok = not args.no_cpu and world == 1        # the CPU legs run at N = 1 only (measurement contract)
    c2 = build_c2(args.scans, seed=0)
    c2_name = f"C2 scan-to-scan point_to_line ICP: {len(c2[4])} consecutive pairs of {args.scans} synthetic 1080-beam 2-D scans"
    c5 = c5_name = None
    if not args.no_c5 or world > 1:
        c5 = build_c5(args.scans, args.pairs, seed=0)
        c5_name = (f"C5 loop-closure candidate batch: {len(c5[4])} scan-pair point_to_line registrations "
                   f"(pairs within 3 m) from {args.scans} scans")
    v = not args.no_verify and not args.no_cpu
    if world == 1:
        line = bench_icp(env, c2_name, c2, "single", args.steps, args.warmup, cpu_ok, True, (200, 128) if v else None)
        if c5 is not None:
            line["c5"] = bench_icp(env, c5_name, c5, "single", max(3, args.steps // 2), 3, False, False, (256, 128) if v else None)
    else:
        line = bench_icp(env, c5_name, c5, "strong", args.steps, args.warmup, False, False, (128, 64) if v else None)
        weak = bench_icp(env, c2_name, c2, "weak", max(3, args.steps // 2), 3, False, False, None)
        if rank == 0:
            line["c2_weak"] = weak
    occ = None
    if not args.no_raycast:
        occ = bench_raycast(args, env.lib, api, env.dev, local_rank, env.pk, rank, world, verify=v)      # all ranks take part
    if rank != 0:
        return
    if occ is not None:
        line["occupancy"] = occ
    if world == 1 and not args.no_extras:
        line["online"] = bench_online(env, c2[0])
        line["extras"] = bench_extras(args, api)
    if args.no_icp_line:
        line = line["occupancy"]
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def bench_raycast(args, lib, api, dev, local_rank, pk, rank=0, world=1, verify=False):
    """C4: 2000 scans x 1080 rays into a 4096 x 4096 grid @ 5 cm.

    Multi-GPU: the grid is cut into bands of 64 rows dealt round-robin over the ranks;
    every rank replays every scan clipped to its own bands, then one NCCL all_gather over
    the device grids reassembles the map (strong scaling: the job is fixed)."""
    import torch
    import torch.distributed as dist
    from icp_b200 import dist as icpd
    from utilities import OccupancyGrid2D
    origins, flat, off = build_c4(args.scans, seed=0)
    grid = OccupancyGrid2D(*GRID_BOUNDS, **GRID_CFG)
    if world > 1:
        grid._dev.set_shard(rank, world)
        icpd.grid_share_setup(grid._dev)            # CUDA IPC handles of every rank's grid, once
    d_org = torch.from_numpy(origins).to(dev)
    d_hits = torch.from_numpy(flat).to(dev)
    d_off = torch.from_numpy(off).to(dev)
    n_rays = int(off[-1])
    stream = torch.cuda.current_stream()            # the explicit stream set by run_ours
    assert stream.cuda_stream != 0
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    ms, ms_update, launches0 = [], [], api.launch_count()
    steps = max(args.steps, 1)
    with ClockSampler(local_rank) as clk:
        for k in range(3 + steps):
            grid.reset()
            flush.zero_()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b, m = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            if k == 3:
                launches0 = api.launch_count()
            a.record(stream)
            grid._dev.update_dev(len(off) - 1, d_org.data_ptr(), d_hits.data_ptr(), d_off.data_ptr(), n_rays,
                                 stream.cuda_stream)
            m.record(stream)
            if world > 1:
                icpd.grid_push_device(grid._dev, stream.cuda_stream)   # touched tiles -> every peer's grid (NVLink peer stores) + barrier
            b.record(stream)
            torch.cuda.synchronize()
            if k >= 3:
                ms.append(a.elapsed_time(b))
                ms_update.append(a.elapsed_time(m))
    launches = api.launch_count() - launches0
    st = grid._dev.last_stats()
    sec = float(np.mean(ms)) / 1e3
    sec_update = float(np.mean(ms_update)) / 1e3       # this rank's update alone (the map still sharded across the GPUs)
    e2e_t = []
    host_out = np.empty((grid.ny, grid.nx), dtype=np.float32)
    pin = api.pinned(host_out, flat, origins, off)
    for k in range(1 + steps):
        grid.reset()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        grid._dev.update(origins, flat, off)
        if world > 1:
            icpd.grid_push_device(grid._dev, stream.cuda_stream)
            torch.cuda.synchronize()
        if rank == 0:                                   # one consumer reads the map: only the tiles the scans touched
            if k == 0:
                host_out[...] = 0.0                     # the caller's mirror; untouched tiles are never written again
            grid._dev.read_dirty(host_out)
        if k >= 1:
            e2e_t.append(time.perf_counter() - t0)
    e2e_s = float(np.mean(e2e_t))
    pin.release()
    checked = None
    if verify and rank == 0:
        # the map the e2e leg just read back (reassembled from every rank's share at N > 1) against the C oracle of
        # mapping.py:103-141 run over the SAME 2000-scan batch: bit for bit
        from oracle import occupancy_oracle
        t0 = time.perf_counter()
        ref = occupancy_oracle.GridOracleC(*GRID_BOUNDS, **GRID_CFG)
        ref.update_many(origins, flat, off, fast=True)
        same = host_out.tobytes() == ref.log_odds.tobytes()
        checked = dict(bit_exact=bool(same), cells=int(host_out.size), nonzero_cells=int(np.count_nonzero(ref.log_odds)),
                       differing_cells=int(np.count_nonzero(host_out != ref.log_odds)), scans=int(len(off) - 1),
                       checker="oracle/occupancy_oracle.c over the whole batch, outside every timed region",
                       seconds=time.perf_counter() - t0)
        if not same:
            raise SystemExit(f"bench.py: the timed occupancy map differs from the oracle: {checked}")
    cells, hits_in = float(st["traversed"]), float(st["hits"])
    if world > 1:
        red = torch.tensor([sec, e2e_s, sec_update], dtype=torch.float64, device=dev)
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        sec_update = float(red[2])
        tot = torch.tensor([cells, hits_in], dtype=torch.float64, device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        sec, e2e_s, cells, hits_in = float(red[0]), float(red[1]), float(tot[0]), float(tot[1])
    if rank != 0:
        return None
    # algorithmic bytes: 16 B endpoint + 8 B per traversed cell + 8 B hit-cell RMW
    alg_bytes = 16.0 * n_rays + 8.0 * cells + 8.0 * hits_in
    cpu = cpu_raycast_baseline(origins, flat, off, min(len(off) - 1, 60)) if not args.no_cpu and world == 1 else None
    return dict(metric="occupancy_rays_per_s", value=n_rays / sec, unit="rays/s", ms_per_step=sec * 1e3,
                n_gpus=world, scaling="strong", update_only_ms=sec_update * 1e3, gather_ms=(sec - sec_update) * 1e3,
                config=dict(workload=f"C4 occupancy log-odds raycast: {len(off) - 1} scans, {n_rays} rays, "
                                     f"{grid.nx}x{grid.ny} grid @ 0.05 m, campus world", **GRID_CFG,
                            cells_per_ray=cells / n_rays, tile_runs=st["runs"],
                            sharding="bands of 64 rows dealt round-robin over the GPUs; every rank walks every ray clipped to its own bands, in scan order; then every rank stores the tiles it touched into every peer's grid (NVLink peer stores through CUDA IPC mappings, icpb200_grid_push_tiles) and a one-element all_reduce is the barrier -- both inside the timed region (update_only_ms = the slowest rank's update without them, gather_ms = the rest); e2e: one rank reads the touched tiles of the reassembled map"),
                e2e=dict(value=n_rays / e2e_s, unit="rays/s",
                         h2d_bytes_per_step=int(origins.nbytes + flat.nbytes + off.nbytes),
                         d2h_bytes_per_step=int(grid._dev.last_tiles_copied * 64 * 64 * 4),
                         api="icpb200_grid_update (+ icpb200_grid_push_tiles and a barrier at N > 1) + icpb200_grid_read_view(log-odds, "
                             "touched tiles only) into the caller's page-locked mirror",
                         tiles_copied=int(grid._dev.last_tiles_copied)),
                gpu_launches=int(launches),
                verify=checked,
                roofline=dict(bound="hbm", achieved=alg_bytes / sec / 1e9 / world, peak=pk["hbm_gbs"], unit="GB/s",
                              frac=alg_bytes / sec / 1e9 / world / pk["hbm_gbs"], traffic=ncu_traffic("occupancy_update")[0], traffic_source=ncu_traffic("occupancy_update")[1],
                              kernel="occ_fast_tiles + binning passes + hit-cell replay (the whole update is timed); per GPU",
                              algorithmic_bytes=alg_bytes, peak_basis=f"{pk['source']} HBM copy bandwidth"),
                cpu_baseline=cpu, clocks=clk.summary())


def c3_roofline(ks, ex, n_target_raw):
    """SURVEY 8(d) grid-NN roofline: bytes per query = 8 (query xy) + 9 x 8 (cell ranges of the 3 x 3 block) + 8 x P
    (candidate coordinates; P measured: target points evaluated per query) + 4 (index out); one grid build pass per
    registration call, M_t x (8 read + 8 write + 4 key).  Against the measured HBM copy bandwidth; the 0.75 MB target
    is L2-resident, so this is an L2-side figure (lts__t_bytes in profiles/)."""
    pk = peaks()
    q = max(ex["grid_queries"], 1)
    p_mean = ex["grid_candidates"] / q
    cells = ex["grid_cells"] / q
    bytes_nn = q * (8.0 + 9 * 8.0 + 16.0 * p_mean + 4.0)         # fp64 xy per candidate here (16 B, not the survey's float2)
    bytes_build = n_target_raw * 20.0
    t = max(ks["pair_kernel_ns"], 1) / 1e9
    return dict(bound="hbm", achieved=bytes_nn / t / 1e9, peak=pk["hbm_gbs"], unit="GB/s", frac=bytes_nn / t / 1e9 / pk["hbm_gbs"],
                traffic=ncu_traffic("icp_pairs_kernel_grid")[0], traffic_source=ncu_traffic("icp_pairs_kernel_grid")[1],
                kernel="icp_pairs_kernel<2, grid> (64 pairs, one CTA each)", queries=int(q), candidates_per_query=p_mean,
                cells_visited_per_query=cells, algorithmic_bytes=bytes_nn, grid_build_bytes=bytes_build,
                grid_build_ms=ks["normals_kernel_ns"] / 1e6, voxel_ms=ks["voxel_kernel_ns"] / 1e6,
                note="64 CTAs on 148 SMs, each a dependent chain of ~20 iterations: latency-bound, far below the memory "
                     "roofline; carried-over correspondences (most queries after the first iterations) issue no grid query")


def build_c3(n_sources=64):
    """C3 inputs (BASELINE.json configs[2]): one ~52k-point submap and n_sources 1080-point scans cut out of it, each with
    an initial guess perturbed by (+0.05 m, -0.04 m, +0.01 rad)."""
    from icp_b200 import synth
    target = synth.submap_cloud(n_raw=52000, seed=3)
    rng = np.random.default_rng(103)
    clouds, R0, t0s = [target], [], []
    while len(clouds) < 1 + n_sources:
        c = target[rng.integers(len(target))]
        near = target[np.hypot(*(target - c).T) < 12.0]
        if len(near) < 1500:
            continue
        pts = near[rng.choice(len(near), 1080, replace=False)] + rng.normal(0, 0.01, size=(1080, 2))
        th = rng.uniform(-0.3, 0.3)
        rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        shift = rng.uniform(-5, 5, size=2)
        clouds.append((pts - shift) @ rot)
        a = th + 0.01
        R0.append([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        t0s.append(shift + [0.05, -0.04])
    flat, off = synth.pack_ragged(clouds)
    return target, flat, off, R0, t0s


def bench_extras(args, api):
    """Smaller configs of BASELINE.json on one GPU (rank 0): C1 teapot, C3 scan -> 50k submap."""
    out = {}
    golden = os.path.join(ROOT, "tests", "golden", "teapot.npz")
    if os.path.exists(golden):                                  # C1: demos/teapot_icp_demo.py:58-65
        g = np.load(golden)
        kw = dict(error_threshold=1e-12, max_iterations=300, voxel_size=0.005, method="point_to_point")
        api.icp_batch([g["moved"]], [g["teapot"]], **kw)
        t0 = time.perf_counter()
        for _ in range(20):
            one = api.icp_batch([g["moved"]], [g["teapot"]], **kw)
        lat = (time.perf_counter() - t0) / 20
        n = 2048
        rng = np.random.default_rng(0)
        srcs = []
        for _ in range(n):                                      # random rigid perturbations
            ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
            ang = rng.uniform(-0.4, 0.4)
            K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            rot = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
            srcs.append(g["teapot"] @ rot.T + rng.uniform(-0.2, 0.2, size=3))
        api.icp_batch(srcs, [g["teapot"]] * n, **kw)
        t0 = time.perf_counter()
        res = api.icp_batch(srcs, [g["teapot"]] * n, **kw)
        dt = time.perf_counter() - t0
        out["C1_teapot_p2p_3d"] = dict(single_call_ms=lat * 1e3, iters=int(one["iters"][0]),
                                       batch_pairs=n, batch_e2e_registrations_per_s=n / dt,
                                       batch_mean_iters=float(res["iters"].mean()))
    # C3: scans against one ~47k-voxel submap, p2p + max_corr_dist (slam.py:217-225)
    from icp_b200 import synth
    target, flat, off, R0, t0s = build_c3()
    kw = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, method="point_to_point", max_corr_dist=1.5,
              R_init=np.asarray(R0), t_init=np.asarray(t0s))
    si, ti = np.arange(1, 65, dtype=np.int32), np.zeros(64, dtype=np.int32)
    api.icp_pairs(flat, off, si, ti, **kw)
    t0 = time.perf_counter()
    res = api.icp_pairs(flat, off, si, ti, **kw)
    dt = time.perf_counter() - t0
    ks = api.icp_last_stats()
    ex = api.icp_extra_stats()
    one_t = time.perf_counter()
    api.icp_pairs(flat, off, si[:1], ti[:1], error_threshold=1e-10, max_iterations=150, voxel_size=0.04,
                  method="point_to_point", max_corr_dist=1.5, R_init=np.asarray(R0)[:1], t_init=np.asarray(t0s)[:1])
    one_t = time.perf_counter() - one_t
    out["C3_scan_to_submap"] = dict(target_raw_points=len(target), pairs=64, e2e_registrations_per_s=64 / dt,
                                    single_call_ms=one_t * 1e3, mean_iters=float(res["iters"].mean()),
                                    voxel_kernel_ms=ks["voxel_kernel_ns"] / 1e6, grid_kernel_ms=ks["normals_kernel_ns"] / 1e6,
                                    pair_kernel_ms=ks["pair_kernel_ns"] / 1e6,
                                    roofline=c3_roofline(ks, ex, len(target)),
                                    note="the 52k-point target is re-downsampled and re-gridded inside every call, "
                                         "as ICP() does (icp.py:150-151)")
    # F2: the same target as a device-resident window (40 pushes), single calls against it
    from utilities import DeviceSubmap
    import contextlib
    import io
    sm = DeviceSubmap(40)
    step = -(-len(target) // 40)
    for k in range(40):
        sm.append(target[k * step:(k + 1) * step])
    R1, t1 = np.asarray(R0)[0], np.asarray(t0s)[0]
    src1 = flat[off[1]:off[2]]
    with contextlib.redirect_stdout(io.StringIO()):
        t0 = time.perf_counter()
        first = sm.ICP(src1, 1e-4, 1e-10, 150, 0.04, R_init=R1, t_init=t1, method="point_to_point", max_corr_dist=1.5)
        first_ms = (time.perf_counter() - t0) * 1e3
        lat = []
        for k in range(20):
            t0 = time.perf_counter()
            again = sm.ICP(src1, 1e-4, 1e-10, 150, 0.04, R_init=R1, t_init=t1, method="point_to_point", max_corr_dist=1.5)
            lat.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        sm.append(target[:step] + 1e-3)
        moved = sm.ICP(src1, 1e-4, 1e-10, 150, 0.04, R_init=R1, t_init=t1, method="point_to_point", max_corr_dist=1.5)
        push_ms = (time.perf_counter() - t0) * 1e3
    out["F2_device_resident_submap"] = dict(window_scans=40, window_points=int(sm.size()[1]),
                                            call_ms_window_unchanged=float(np.median(lat)) * 1e3, first_call_ms=first_ms,
                                            push_plus_call_ms=push_ms, single_call_ms_host_target=one_t * 1e3,
                                            note="ICP(scan, submap) with the window resident on the device (submap voxel 1e-4: the window "
                                                 "itself is the target, as in C3): unchanged window -> cached downsample + hash grid, the call "
                                                 "is source upload + pair kernel; after a push both downsamples and the grid are redone on the "
                                                 "device (no 52k-point upload)")
    sm.close()
    # F3: _rebuild_map (slam.py:271-277) -- the C4 scans in their local frames + poses, one call
    from utilities import OccupancyGrid2D
    c4_scans, c4_poses = synth.make_sequence(args.scans, world="campus", seed=0)
    mats = np.array([[[np.cos(t), -np.sin(t), x], [np.sin(t), np.cos(t), y], [0.0, 0.0, 1.0]] for x, y, t in c4_poses])
    grid = OccupancyGrid2D(*GRID_BOUNDS, **GRID_CFG)
    lflat, loff = synth.pack_ragged(c4_scans)
    pin = api.pinned(lflat, loff, mats)
    rb = []
    for k in range(4):
        t0 = time.perf_counter()
        grid._dev.rebuild(mats, lflat, loff)
        rb.append(time.perf_counter() - t0)
    pin.release()
    out["F3_rebuild_map"] = dict(scans=len(c4_scans), rays=int(loff[-1]), call_ms=float(np.min(rb[1:])) * 1e3,
                                 rays_per_s=float(loff[-1]) / float(np.min(rb[1:])),
                                 note="reset + transform_points_2d on the device + raycast of the whole history, host buffers in "
                                      "(the map stays on the device); reference: reset + update_scan per scan, about 0.39 s per scan")
    grid._dev.close()
    # F1: rotation-search pre-alignment, the reference's config values
    # (config.yaml:37-39: voxel 0.15, coarse 1.5 deg = 240 angles, fine 0.1 deg = 30 angles)
    from utilities import rotation_search
    import contextlib
    import io
    scans, _ = synth.make_sequence(258, world="room", seed=0)
    with contextlib.redirect_stdout(io.StringIO()):
        rotation_search(scans[0], scans[1], voxel_size=0.15, angle_step_coarse=1.5, angle_step_fine=0.1)     # warm
        t0 = time.perf_counter()
        for k in range(20):
            rotation_search(scans[k], scans[k + 1], voxel_size=0.15, angle_step_coarse=1.5, angle_step_fine=0.1)
        call_ms = (time.perf_counter() - t0) / 20 * 1e3
    n = 256
    ds = [api.voxel_downsample(scans[k], 0.15) for k in range(n + 1)]
    srcs = [d - d.mean(axis=0) for d in ds[:n]]
    tgts = ds[1:n + 1]
    angles = np.deg2rad(np.arange(-180, 180, 1.5))
    api.rotation_scores(srcs, tgts, [angles] * n, [t.mean(axis=0) for t in tgts])     # warm: buffers grow to the batch's size
    t0 = time.perf_counter()
    sc = api.rotation_scores(srcs, tgts, [angles] * n, [t.mean(axis=0) for t in tgts])
    dt = time.perf_counter() - t0
    evals = float(sum(len(a) * len(b) for a, b in zip(srcs, tgts))) * len(angles)
    out["F1_rotation_search"] = dict(single_call_ms=call_ms, angles_per_call=270, batch_problems=n,
                                     batch_coarse_sweeps_per_s=n / dt, batch_pair_evals_per_s=evals / dt,
                                     points_after_voxel=float(np.mean([len(d) for d in ds])),
                                     best_angles_deg=[float(np.degrees(angles[int(np.argmin(s))])) for s in sc[:4]])
    out["F4_pose_graph"] = pose_graph_extra()
    return out


def pose_graph_extra():
    """PoseGraph2D.optimize on a closed 2000-node trajectory with 200 loop closures (host solver of the
    library), and on 500 nodes beside the oracle's dense solve (the reference's algorithm: O(n^3) per iteration)."""
    import contextlib
    import io
    from utilities.pose_graph import PoseGraph2D
    from oracle import pose_graph_oracle

    def graph(n, loops, seed):
        rng = np.random.default_rng(seed)
        th = np.cumsum(np.full(n, 2 * np.pi / n))
        truth = np.column_stack([np.cumsum(0.3 * np.cos(th)), np.cumsum(0.3 * np.sin(th)), pose_graph_oracle.wrap(th)])

        def rel(a, b):
            c, s_ = np.cos(a[2]), np.sin(a[2])
            d = b[:2] - a[:2]
            return np.array([c * d[0] + s_ * d[1], -s_ * d[0] + c * d[1], pose_graph_oracle.wrap(b[2] - a[2])])
        est = truth + rng.normal(0, [0.05, 0.05, 0.01], size=truth.shape)
        est[0] = truth[0]
        edges = [(k - 1, k, rel(truth[k - 1], truth[k]) + rng.normal(0, 0.002, 3), np.diag([100.0, 100.0, 400.0])) for k in range(1, n)]
        for _ in range(loops):
            a, b = rng.choice(n, size=2, replace=False)
            edges.append((int(max(a, b)), int(min(a, b)), rel(truth[max(a, b)], truth[min(a, b)]) + rng.normal(0, 0.002, 3),
                          np.diag([200.0, 200.0, 800.0])))
        return est, edges

    def ours(est, edges):
        pg = PoseGraph2D()
        for p_ in est:
            pg.add_node(p_)
        for e in edges:
            pg.add_edge(*e)
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()) as log:
            pg.optimize(n_iterations=20)
        return time.perf_counter() - t0, np.array(pg.nodes), log.getvalue().strip()
    est, edges = graph(2000, 200, 0)
    t_big, _, line = ours(est, edges)
    est, edges = graph(500, 50, 1)
    t_small, nodes, _ = ours(est, edges)
    t0 = time.perf_counter()
    want, it, _, _ = pose_graph_oracle.optimize(est, edges, n_iterations=20)
    t_ref = time.perf_counter() - t0
    d = nodes - want
    d[:, 2] = pose_graph_oracle.wrap(d[:, 2])
    return dict(nodes=2000, loop_closures=200, optimize_ms=t_big * 1e3, result=line,
                nodes_500_ms=t_small * 1e3, nodes_500_reference_algorithm_ms=t_ref * 1e3, max_pose_diff_vs_oracle=float(np.abs(d).max()),
                note="host block-skyline Cholesky (icpb200_pose_graph_optimize) against the reference's dense np.linalg.solve per "
                     "iteration (oracle/pose_graph_oracle.py, bit-identical to the reference); the reference at 2000 nodes solves a "
                     "6000 x 6000 system per iteration")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scans", type=int, default=2000)
    ap.add_argument("--pairs", type=int, default=8192)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-raycast", action="store_true", help="skip the occupancy (C4) leg")
    ap.add_argument("--no-icp-line", action="store_true", help="print only the occupancy object (profiling aid)")
    ap.add_argument("--no-extras", action="store_true", help="skip the online / C1 / C3 extras")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 object of the one-GPU line")
    ap.add_argument("--no-verify", action="store_true", help="skip the comparison of the timed outputs with the oracle")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: anything libraries print (NCCL's version banner, the
    # shim's per-call ICP line) goes to stderr instead
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
