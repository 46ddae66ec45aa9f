"""One-process-per-GPU sharding of the two batch seams (torch.distributed).

* registrations: pairs are independent, so rank r registers a contiguous block
  of the pair list with no data-path collective; one all_gather returns every
  rank's (R, t, error, iters, status) to all ranks.
* occupancy replay: rank r owns a horizontal strip of the grid -- the 64-cell tile
  rows [r*T/world, (r+1)*T/world) (the rule libicp_b200 applies in
  icpb200_grid_set_shard); every rank replays every scan clipped to its own strip,
  in scan order, so the clamp order matches the reference.  Cells a rank does not
  own stay exactly 0: one all_gather of the strips (equally tall strips) or one
  all_reduce(SUM) reassembles the map bit-exactly.

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


def icp_pairs_sharded(points, cloud_off, src_idx, tgt_idx, *args, compute=None, R_init=None, t_init=None, **kw):
    """Register all pairs, each rank doing its block; every rank returns the full result dict.

    ``compute`` defaults to :func:`icp_b200.api.icp_pairs` (tests inject a CPU stand-in)."""
    if compute is None:
        from .api import icp_pairs as compute
    rank, size = world()
    src_idx = np.asarray(src_idx, dtype=np.int32)
    tgt_idx = np.asarray(tgt_idx, dtype=np.int32)
    n = len(src_idx)
    lo, hi = shard_range(n, rank, size)
    if R_init is not None and t_init is not None:
        kw = dict(kw, R_init=np.asarray(R_init)[lo:hi], t_init=np.asarray(t_init)[lo:hi])
    part = compute(points, cloud_off, src_idx[lo:hi], tgt_idx[lo:hi], *args, **kw) if hi > lo else None
    if size == 1:
        return part
    dim = int(np.asarray(points).shape[1])
    width = dim * dim + dim + 2 + 2                       # R | t | error, prev_error | iters, status
    counts = [shard_range(n, r, size)[1] - shard_range(n, r, size)[0] for r in range(size)]
    dev = _comm_device()
    mine = torch.zeros((max(counts), width), dtype=torch.float64, device=dev)
    if part is not None:
        packed = np.concatenate([part["R"].reshape(hi - lo, -1), part["t"], part["error"][:, None],
                                 part["prev_error"][:, None], part["iters"][:, None].astype(np.float64),
                                 part["status"][:, None].astype(np.float64)], axis=1)
        mine[:hi - lo] = torch.from_numpy(packed).to(dev)
    bufs = [torch.empty_like(mine) for _ in range(size)]
    dist.all_gather(bufs, mine)
    full = np.concatenate([b[:c].cpu().numpy() for b, c in zip(bufs, counts)], axis=0)
    dd = dim * dim
    return dict(R=full[:, :dd].reshape(n, dim, dim), t=full[:, dd:dd + dim], error=full[:, dd + dim],
                prev_error=full[:, dd + dim + 1], iters=full[:, dd + dim + 2].astype(np.int32),
                status=full[:, dd + dim + 3].astype(np.int32))


def strip_rows(ny, rank, size):
    """Rows [lo, hi) of the grid owned by `rank` (occ_strip_begin in csrc/occupancy.h)."""
    tiles_y = (ny + TILE - 1) // TILE
    lo = min((tiles_y * rank // size) * TILE, ny)
    hi = min((tiles_y * (rank + 1) // size) * TILE, ny)
    return lo, hi


def owned_tile_mask(nx, ny, rank, size):
    """(ny, nx) bool mask of the cells this rank owns."""
    lo, hi = strip_rows(ny, rank, size)
    mask = np.zeros((ny, nx), dtype=bool)
    mask[lo:hi] = True
    return mask


def grid_allreduce_host(log_odds):
    """Sum the per-rank partial maps held as host arrays (gloo / tests)."""
    rank, size = world()
    if size == 1:
        return log_odds
    t = torch.from_numpy(np.ascontiguousarray(log_odds)).to(_comm_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


class _DevArray:
    """Minimal __cuda_array_interface__ holder so torch can wrap a raw device pointer."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = dict(shape=shape, typestr="<f4", data=(int(ptr), False), version=3)


def grid_device_tensor(device_grid):
    """Zero-copy torch view (ny, nx) float32 of a DeviceGrid's HBM buffer."""
    return torch.as_tensor(_DevArray(device_grid.device_ptr(), (device_grid.ny, device_grid.nx)),
                           device=torch.device("cuda", torch.cuda.current_device()))


def grid_allreduce_device(device_grid, sync=True):
    """Reassemble the sharded grid on every rank, in place on the device (NCCL, B200 box): an all_gather of the
    strips when they are equally tall (each rank sends only its own rows), else an all_reduce(SUM).

    sync=False leaves everything stream-ordered on torch's current stream (the caller ran the update on that
    stream -- icpb200_grid_update_dev -- and reads the result on it): no host wait on either side."""
    rank, size = world()
    if size > 1:
        t = grid_device_tensor(device_grid)
        ny = t.shape[0]
        rows = [strip_rows(ny, r, size) for r in range(size)]
        if sync:
            torch.cuda.synchronize()
        if len({hi - lo for lo, hi in rows}) == 1 and rows[-1][1] == ny:
            lo, hi = rows[rank]
            dist.all_gather_into_tensor(t, t[lo:hi])
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if sync:
            torch.cuda.synchronize()
