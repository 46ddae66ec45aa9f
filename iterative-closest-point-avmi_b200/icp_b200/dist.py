"""One-process-per-GPU sharding of the two batch seams (torch.distributed).

* registrations: pairs are independent, so every rank registers its share of the
  pair list (plan_pair_shards: locality-sorted chunks dealt round-robin) with no
  data-path collective and uploads only the clouds its pairs reference; one
  all_gather of packed result blocks returns (R, t, error, iters, status) of
  every pair to all ranks.
* occupancy replay: the grid is cut into bands of 64 rows dealt round-robin -- rank
  r owns the bands b with b % world == r (the rule libicp_b200 applies in
  icpb200_grid_set_shard); every rank replays every scan clipped to its own bands,
  in scan order, so the clamp order matches the reference.  One all_gather of the
  packed bands reassembles the map bit-exactly (a gather, never a sum: rows a rank
  does not own are overwritten).

NCCL (GPU tensors) on the B200 box, gloo (CPU tensors) in the CPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

TILE = 64          # must equal kOccOwnTile in csrc/occupancy.h


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, size):
    """Contiguous block [lo, hi) of n items for `rank`; blocks differ by at most one item."""
    base, extra = divmod(n, size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _comm_device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def plan_pair_shards(src_idx, tgt_idx, size, chunks_per_rank=4):
    """Which pairs each of ``size`` ranks registers: a list of index arrays into the pair list.

    Pairs are sorted by the lower of their two cloud indices and cut into ``size * chunks_per_rank`` contiguous
    chunks that are dealt round-robin.  A rank's pairs then name a few contiguous ranges of the scan history, so
    the upload and the per-cloud kernels (voxel means, normals) shard with the pairs instead of being repeated on
    every rank, while a stretch of slow pairs (the reference's limit cycles) is spread over all ranks.
    Every rank computes the same plan from the same arguments -- nothing is exchanged."""
    src_idx = np.asarray(src_idx)
    tgt_idx = np.asarray(tgt_idx)
    n = len(src_idx)
    key = np.minimum(src_idx, tgt_idx)
    # (a stable sort either way; numpy sorts 16-bit keys by radix, ten times faster than its merge sort of int32:
    # 0.06 against 0.7 ms for 8192 pairs -- this runs inside every sharded call)
    if n and 0 <= int(key.min()) and int(key.max()) < 65536:
        order = np.argsort(key.astype(np.uint16), kind="stable")
    else:
        order = np.argsort(key, kind="stable")
    n_chunks = max(1, size * chunks_per_rank)
    bounds = (np.arange(n_chunks + 1, dtype=np.int64) * n) // n_chunks
    plan = []
    for r in range(size):
        parts = [order[bounds[c]:bounds[c + 1]] for c in range(r, n_chunks, size)]
        plan.append(np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64))
    return plan


_FIELDS = (("R", np.float64, lambda d: d * d), ("t", np.float64, lambda d: d), ("error", np.float64, lambda d: 1),
           ("prev_error", np.float64, lambda d: 1), ("iters", np.int32, lambda d: 1), ("status", np.int32, lambda d: 1))


def result_layout(n_max, dim):
    """Byte layout of one rank's result block (every field padded to n_max pairs): {name: (offset, dtype, width)}, total."""
    at, out = 0, {}
    for name, dt, width in _FIELDS:
        out[name] = (at, dt, width(dim))
        at += (n_max * width(dim) * np.dtype(dt).itemsize + 63) // 64 * 64
    return out, at


def _take_rows(plan):
    """For a plan: the row of the gathered (size * n_max)-row table that holds pair i, for every i in the caller's order."""
    n_max = max(len(p) for p in plan)
    where = np.concatenate([r * n_max + np.arange(len(p), dtype=np.int64) for r, p in enumerate(plan)])
    perm = np.concatenate(plan)
    take = np.empty(len(perm), dtype=np.int64)
    take[perm] = where
    return take


def _unpack_results(blocks, plan, dim, take=None):
    """blocks: (size, bytes) uint8 array of the gathered result blocks -> full result dict in the caller's pair order.
    One gather per field through ``take`` (:func:`_take_rows`)."""
    n_max = max(len(p) for p in plan)
    layout, _ = result_layout(n_max, dim)
    if take is None:
        take = _take_rows(plan)
    blocks = np.ascontiguousarray(blocks)
    full = {}
    for name, (at, dt, width) in layout.items():
        nb = n_max * width * np.dtype(dt).itemsize
        table = np.ascontiguousarray(blocks[:, at:at + nb]).view(dt).reshape(len(plan) * n_max, width)
        rows = np.take(table, take, axis=0)               # (10x faster than table[take] on rows)
        full[name] = rows.reshape((len(take), dim, dim)) if name == "R" else rows.reshape(len(take), width) if width > 1 else rows.reshape(-1)
    return full


def icp_pairs_sharded(points, cloud_off, src_idx, tgt_idx, *args, compute=None, R_init=None, t_init=None,
                      chunks_per_rank=4, **kw):
    """Register all pairs, each rank doing its share (:func:`plan_pair_shards`); every rank returns the full result
    dict in the caller's pair order.  Host buffers in, host arrays out: the batch seam of slam.py:575-579.

    A rank hands the library the whole scan history and its own pairs; ``icpb200_icp_pairs`` uploads only the clouds
    those pairs reference.  One all_gather of the packed result blocks is the only exchange.
    ``compute`` defaults to :func:`icp_b200.api.icp_pairs` (the gloo tests inject a CPU stand-in)."""
    if compute is None:
        from .api import icp_pairs as compute
    rank, size = world()
    src_idx = np.ascontiguousarray(src_idx, dtype=np.int32)
    tgt_idx = np.ascontiguousarray(tgt_idx, dtype=np.int32)
    if size == 1:
        if R_init is not None and t_init is not None:
            kw = dict(kw, R_init=R_init, t_init=t_init)
        return compute(points, cloud_off, src_idx, tgt_idx, *args, **kw)
    plan = plan_pair_shards(src_idx, tgt_idx, size, chunks_per_rank)
    take = _take_rows(plan)
    mine = plan[rank]
    if R_init is not None and t_init is not None:
        kw = dict(kw, R_init=np.asarray(R_init)[mine], t_init=np.asarray(t_init)[mine])
    part = compute(points, cloud_off, src_idx[mine], tgt_idx[mine], *args, **kw) if len(mine) else None
    dim = int(np.asarray(points).shape[1])
    layout, nbytes = result_layout(max(len(p) for p in plan), dim)
    block = np.zeros(nbytes, dtype=np.uint8)
    if part is not None:
        for name, (at, dt, width) in layout.items():
            vals = np.ascontiguousarray(part[name], dtype=dt).reshape(-1)
            block[at:at + vals.nbytes] = vals.view(np.uint8)
    dev = _comm_device()
    send = torch.from_numpy(block).to(dev)
    if dist.get_backend() == "nccl":
        recv = torch.empty((size, nbytes), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(recv, send)
        blocks = recv.cpu().numpy()
    else:
        bufs = [torch.empty_like(send) for _ in range(size)]
        dist.all_gather(bufs, send)
        blocks = np.stack([b.cpu().numpy() for b in bufs])
    return _unpack_results(blocks, plan, dim, take)


def gather_result_blocks(part, dim):
    """all_gather of one rank's result dict (host arrays) as a packed block; returns the (size, bytes) uint8 array."""
    rank, size = world()
    n = len(part["iters"])
    layout, nbytes = result_layout(n, dim)
    block = np.zeros(nbytes, dtype=np.uint8)
    for name, (at, dt, width) in layout.items():
        vals = np.ascontiguousarray(part[name], dtype=dt).reshape(-1)
        block[at:at + vals.nbytes] = vals.view(np.uint8)
    if size == 1:
        return block[None]
    dev = _comm_device()
    send = torch.from_numpy(block).to(dev)
    bufs = [torch.empty_like(send) for _ in range(size)]
    dist.all_gather(bufs, send)
    return np.stack([b.cpu().numpy() for b in bufs])


class DevicePairShard:
    """This rank's share of a pair batch, resident in HBM: the clouds its pairs reference (compacted), its pair
    lists and one packed result block.  ``enqueue()`` runs the registration stream-ordered (icpb200_icp_pairs_dev)
    and gathers every rank's result block with one NCCL all_gather; ``results()`` unpacks them on the host."""

    def __init__(self, points, cloud_off, src_idx, tgt_idx, device, chunks_per_rank=4, replicated=False):
        """replicated = True: every rank registers the WHOLE batch (weak scaling); the gather then moves every
        rank's full result block."""
        from . import _lib
        self._lib_mod = _lib
        self.lib = _lib.load()
        self.rank, self.size = world()
        points = np.asarray(points, dtype=np.float64)
        off = np.asarray(cloud_off, dtype=np.int64)
        src_idx = np.asarray(src_idx, dtype=np.int32)
        tgt_idx = np.asarray(tgt_idx, dtype=np.int32)
        self.dim = int(points.shape[1])
        self.replicated = bool(replicated) or self.size == 1
        if self.replicated:
            self.plan = [np.arange(len(src_idx), dtype=np.int64)] * self.size
        else:
            self.plan = plan_pair_shards(src_idx, tgt_idx, self.size, chunks_per_rank)
        mine = self.plan[self.rank]
        self.n_mine = len(mine)
        used = np.unique(np.concatenate([src_idx[mine], tgt_idx[mine]])) if self.n_mine else np.zeros(0, dtype=np.int64)
        remap = np.full(len(off) - 1, -1, dtype=np.int32)
        remap[used] = np.arange(len(used), dtype=np.int32)
        lens = off[used + 1] - off[used]
        new_off = np.zeros(len(used) + 1, dtype=np.int64)
        new_off[1:] = np.cumsum(lens)
        flat = np.concatenate([points[off[c]:off[c + 1]] for c in used]) if len(used) else np.zeros((0, self.dim))
        self.n_clouds, self.max_pts = len(used), int(lens.max()) if len(used) else 0
        self.h2d_bytes = flat.nbytes + new_off.nbytes + 2 * 4 * self.n_mine
        self.d_pts = torch.from_numpy(flat).to(device)
        self.d_off = torch.from_numpy(new_off).to(device)
        self.d_si = torch.from_numpy(remap[src_idx[mine]]).to(device)
        self.d_ti = torch.from_numpy(remap[tgt_idx[mine]]).to(device)
        self.layout, self.nbytes = result_layout(max(len(p) for p in self.plan), self.dim)
        self.block = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.gathered = torch.zeros((self.size, self.nbytes), dtype=torch.uint8, device=device)
        self._ptr = {name: self.block.data_ptr() + at for name, (at, _, _) in self.layout.items()}

    def enqueue(self, stream, error_threshold, max_iterations, voxel_size, method="point_to_point", normal_k=10,
                max_corr_dist=None, gather=True):
        """Stream-ordered on ``stream`` (a torch.cuda.Stream that is also torch's current stream)."""
        L = self._lib_mod
        if self.n_mine:
            rc = self.lib.icpb200_icp_pairs_dev(
                self.n_clouds, self.dim, self.d_pts.data_ptr(), self.d_off.data_ptr(), self.max_pts, self.n_mine,
                self.d_si.data_ptr(), self.d_ti.data_ptr(), None, None, float(error_threshold), int(max_iterations),
                float(voxel_size), L.POINT_TO_LINE if method == "point_to_line" else L.POINT_TO_POINT, int(normal_k),
                -1.0 if max_corr_dist is None else float(max_corr_dist), L.NN_AUTO,
                self._ptr["R"], self._ptr["t"], self._ptr["error"], self._ptr["prev_error"], self._ptr["iters"],
                self._ptr["status"], stream.cuda_stream)
            L.check(rc, "icpb200_icp_pairs_dev")
        if gather and self.size > 1:
            dist.all_gather_into_tensor(self.gathered, self.block)

    def results(self):
        """Full result dict in the caller's pair order (host arrays); call after a gathered enqueue()."""
        torch.cuda.synchronize()
        if self.replicated:
            return _unpack_results(self.block.cpu().numpy()[None], self.plan[:1], self.dim)
        return _unpack_results(self.gathered.cpu().numpy(), self.plan, self.dim)


def owned_rows(ny, rank, size):
    """(ny,) bool: the grid rows `rank` owns -- bands of TILE rows dealt round-robin (occ_owner in csrc/occupancy.h)."""
    return (np.arange(ny) // TILE) % size == rank


def owned_tile_mask(nx, ny, rank, size):
    """(ny, nx) bool mask of the cells this rank owns."""
    return np.repeat(owned_rows(ny, rank, size)[:, None], nx, axis=1)


def grid_gather_host(log_odds):
    """Reassemble the sharded map from per-rank host arrays (gloo / tests): every rank contributes the rows it owns.
    A gather, not a sum: what a rank holds in rows it does not own (zeros, or a previously reassembled map) is ignored."""
    rank, size = world()
    if size == 1:
        return log_odds
    ny, nx = log_odds.shape
    mine = np.ascontiguousarray(log_odds[owned_rows(ny, rank, size)])
    counts = [int(owned_rows(ny, r, size).sum()) for r in range(size)]
    send = torch.zeros((max(counts), nx), dtype=torch.float32)
    send[:len(mine)] = torch.from_numpy(mine)
    send = send.to(_comm_device())
    bufs = [torch.empty_like(send) for _ in range(size)]
    dist.all_gather(bufs, send)
    out = np.array(log_odds, copy=True)
    for r in range(size):
        out[owned_rows(ny, r, size)] = bufs[r][:counts[r]].cpu().numpy()
    return out


class _DevArray:
    """Minimal __cuda_array_interface__ holder so torch can wrap a raw device pointer."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = dict(shape=shape, typestr="<f4", data=(int(ptr), False), version=3)


def grid_device_tensor(device_grid):
    """Zero-copy torch view (ny, nx) float32 of a DeviceGrid's HBM buffer."""
    return torch.as_tensor(_DevArray(device_grid.device_ptr(), (device_grid.ny, device_grid.nx)),
                           device=torch.device("cuda", torch.cuda.current_device()))


def grid_gather_device(device_grid, sync=True):
    """Reassemble the sharded grid on every rank, in place on the device (NCCL, B200 box).

    Rank r's bands (rows [64 b, 64 b + 64) with b % size == r) are packed into one contiguous block, one
    all_gather_into_tensor moves every rank's block to every rank, and the blocks are scattered back to their rows.
    A gather, never a sum: rows a rank does not own are overwritten, so update -> gather -> update -> gather works
    without a reset in between.  sync=False leaves everything stream-ordered on torch's current stream (the caller ran
    the update on that stream -- icpb200_grid_update_dev -- and reads the result on it): no host wait on either side."""
    rank, size = world()
    if size == 1:
        return
    t = grid_device_tensor(device_grid)
    ny, nx = t.shape
    per = -(-ny // (TILE * size))                      # bands per rank, the last ones possibly short or missing
    if sync:
        torch.cuda.synchronize()
    stage = getattr(device_grid, "_gather_stage", None)
    if stage is None or stage[0].shape != (per, TILE, nx):
        stage = (torch.empty((per, TILE, nx), dtype=torch.float32, device=t.device),
                 torch.empty((size, per, TILE, nx), dtype=torch.float32, device=t.device),
                 None if ny == per * size * TILE else torch.zeros((per * size * TILE, nx), dtype=torch.float32, device=t.device))
        device_grid._gather_stage = stage
    send, recv, padded = stage
    full = t
    if padded is not None:                            # ny is not a multiple of 64 * size: work on a padded copy
        padded[:ny].copy_(t)
        full = padded
    bands = full.view(per, size, TILE, nx)
    send.copy_(bands[:, rank])
    dist.all_gather_into_tensor(recv, send)
    bands.copy_(recv.permute(1, 0, 2, 3))
    if padded is not None:
        t.copy_(padded[:ny])
    if sync:
        torch.cuda.synchronize()


def grid_share_setup(device_grid):
    """Once per grid, after ``set_shard``: exchange the CUDA IPC handles of every rank's grid (one all_gather of 128 bytes
    per rank) and map the peers' grids into this process, so that ``grid_push_device`` can store tiles into them."""
    rank, size = world()
    if size == 1:
        return
    mine = torch.frombuffer(bytearray(device_grid.ipc_export()), dtype=torch.uint8).to(_comm_device())
    bufs = [torch.empty_like(mine) for _ in range(size)]
    dist.all_gather(bufs, mine)
    device_grid.ipc_attach(size, rank, b"".join(bytes(b.cpu().numpy().tobytes()) for b in bufs))
    dist.barrier()


_TOKEN = {}


def grid_push_device(device_grid, stream=0):
    """Make this rank's touched tiles part of every peer's map: one kernel of NVLink peer stores (icpb200_grid_push_tiles)
    followed by a one-element all_reduce as the cross-rank barrier -- both stream-ordered on torch's current stream (pass
    its handle as ``stream``).  Moves the explored part of the map only (C4: 8 MB in all, against 64 MiB for a gather of
    the bands); needs ``grid_share_setup`` once."""
    rank, size = world()
    if size == 1:
        return
    device_grid.push_tiles(stream)
    dev = torch.cuda.current_device()
    if dev not in _TOKEN:
        _TOKEN[dev] = torch.zeros(1, dtype=torch.float32, device=torch.device("cuda", dev))
    dist.all_reduce(_TOKEN[dev])
