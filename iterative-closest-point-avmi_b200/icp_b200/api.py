"""numpy-facing wrappers over the C ABI: batched registration, voxel
downsample, occupancy-grid handle.  Host arrays in, host arrays out; all
compute happens in libicp_b200.so on the GPU."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import (c_double_p, c_float_p, c_int32_p, c_int64_p, check)

METHODS = {"point_to_point": _lib.POINT_TO_POINT, "point_to_line": _lib.POINT_TO_LINE}
NN_MODES = {"auto": _lib.NN_AUTO, "brute": _lib.NN_BRUTE, "grid": _lib.NN_GRID}


def init(device=-1):
    check(_lib.load().icpb200_init(int(device)), "icpb200_init")


def shutdown():
    _lib.load().icpb200_shutdown()


def launch_count():
    return int(_lib.load().icpb200_launch_count())


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


_ONE = {}


def _ptr(a, typ):
    """The array as a ctypes argument of pointer type ``typ``.  A one-element ctypes array laid over the buffer converts to
    the pointer and costs 1.2 us; ``a.ctypes.data_as`` costs 4.5 us, and a call marshals a dozen arrays (a fifth of the
    Python side of one ICP() call).  Read-only or empty arrays take the slow way."""
    if a is None:
        return None
    try:
        one = _ONE.get(typ)
        if one is None:
            one = _ONE[typ] = typ._type_ * 1
        return one.from_buffer(a)
    except (TypeError, ValueError, BufferError):
        return a.ctypes.data_as(typ)


def _method_code(method):
    # the reference compares the string (icp.py:162); anything that is not
    # "point_to_line" takes the point-to-point branch
    return _lib.POINT_TO_LINE if method == "point_to_line" else _lib.POINT_TO_POINT


def _init_arrays(R_init, t_init, n_pairs, dim):
    if R_init is None or t_init is None:          # icp.py:153: both or neither
        return None, None
    r = _f64(R_init).reshape(n_pairs, dim, dim)
    t = _f64(t_init).reshape(n_pairs, dim)
    return r, t


def _alloc_out(n_pairs, dim):
    return (np.empty((n_pairs, dim, dim)), np.empty((n_pairs, dim)), np.empty(n_pairs), np.empty(n_pairs),
            np.empty(n_pairs, dtype=np.int32), np.empty(n_pairs, dtype=np.int32))


def icp_batch(sources, targets, error_threshold, max_iterations, voxel_size, R_init=None, t_init=None,
              method="point_to_point", normal_k=10, max_corr_dist=None, nn_mode="auto"):
    """Register sources[p] onto targets[p] for every p.  Returns dict of arrays
    R (n,d,d), t (n,d), error (n,), iters (n,), status (n,)."""
    n_pairs = len(sources)
    assert len(targets) == n_pairs
    dim = int(np.asarray(sources[0]).shape[1])
    from .synth import pack_ragged
    src, src_off = pack_ragged([_f64(s) for s in sources], dim)
    tgt, tgt_off = pack_ragged([_f64(t) for t in targets], dim)
    r0, t0 = _init_arrays(R_init, t_init, n_pairs, dim)
    R, t, err, prev, iters, status = _alloc_out(n_pairs, dim)
    rc = _lib.load().icpb200_icp_batch(
        n_pairs, dim, _ptr(src, c_double_p), _ptr(src_off, c_int64_p), _ptr(tgt, c_double_p), _ptr(tgt_off, c_int64_p),
        _ptr(r0, c_double_p), _ptr(t0, c_double_p), float(error_threshold), int(max_iterations), float(voxel_size),
        _method_code(method), int(normal_k), -1.0 if max_corr_dist is None else float(max_corr_dist), NN_MODES[nn_mode],
        _ptr(R, c_double_p), _ptr(t, c_double_p), _ptr(err, c_double_p), _ptr(prev, c_double_p), _ptr(iters, c_int32_p), _ptr(status, c_int32_p))
    check(rc, "icpb200_icp_batch")
    return dict(R=R, t=t, error=err, prev_error=prev, iters=iters, status=status)


def icp_pairs(points, cloud_off, src_idx, tgt_idx, error_threshold, max_iterations, voxel_size, R_init=None,
              t_init=None, method="point_to_point", normal_k=10, max_corr_dist=None, nn_mode="auto"):
    """Register cloud src_idx[p] onto cloud tgt_idx[p]; clouds are rows
    cloud_off[c]:cloud_off[c+1] of ``points``."""
    pts = _f64(points)
    dim = int(pts.shape[1])
    off = np.ascontiguousarray(cloud_off, dtype=np.int64)
    si = np.ascontiguousarray(src_idx, dtype=np.int32)
    ti = np.ascontiguousarray(tgt_idx, dtype=np.int32)
    n_pairs = len(si)
    r0, t0 = _init_arrays(R_init, t_init, n_pairs, dim)
    R, t, err, prev, iters, status = _alloc_out(n_pairs, dim)
    rc = _lib.load().icpb200_icp_pairs(
        len(off) - 1, dim, _ptr(pts, c_double_p), _ptr(off, c_int64_p), n_pairs, _ptr(si, c_int32_p), _ptr(ti, c_int32_p),
        _ptr(r0, c_double_p), _ptr(t0, c_double_p), float(error_threshold), int(max_iterations), float(voxel_size),
        _method_code(method), int(normal_k), -1.0 if max_corr_dist is None else float(max_corr_dist), NN_MODES[nn_mode],
        _ptr(R, c_double_p), _ptr(t, c_double_p), _ptr(err, c_double_p), _ptr(prev, c_double_p), _ptr(iters, c_int32_p), _ptr(status, c_int32_p))
    check(rc, "icpb200_icp_pairs")
    return dict(R=R, t=t, error=err, prev_error=prev, iters=iters, status=status)


def icp_trace(source, target, error_threshold, max_iterations, voxel_size, R_init=None, t_init=None,
              method="point_to_point", normal_k=10, max_corr_dist=None, nn_mode="auto", trace_iters=4):
    """One registration plus its intermediate state (parity tests)."""
    src, tgt = _f64(source), _f64(target)
    dim = int(src.shape[1])
    r0, t0 = _init_arrays(R_init, t_init, 1, dim)
    R, t, err, prev, iters, status = _alloc_out(1, dim)
    src_ds, tgt_ds = np.empty_like(src), np.empty_like(tgt)
    normals = np.zeros((len(tgt), 2))
    matches = np.full((max(trace_iters, 1), len(src)), -1, dtype=np.int32)
    n_s, n_t = ctypes.c_int64(0), ctypes.c_int64(0)
    rc = _lib.load().icpb200_icp_trace(
        dim, _ptr(src, c_double_p), len(src), _ptr(tgt, c_double_p), len(tgt), _ptr(r0, c_double_p), _ptr(t0, c_double_p),
        float(error_threshold), int(max_iterations), float(voxel_size), _method_code(method), int(normal_k),
        -1.0 if max_corr_dist is None else float(max_corr_dist), NN_MODES[nn_mode],
        _ptr(R, c_double_p), _ptr(t, c_double_p), _ptr(err, c_double_p), _ptr(prev, c_double_p), _ptr(iters, c_int32_p), _ptr(status, c_int32_p),
        _ptr(src_ds, c_double_p), ctypes.byref(n_s), _ptr(tgt_ds, c_double_p), ctypes.byref(n_t),
        _ptr(normals, c_double_p), _ptr(matches, c_int32_p), int(trace_iters))
    check(rc, "icpb200_icp_trace")
    ns, nt = int(n_s.value), int(n_t.value)
    return dict(R=R[0], t=t[0], error=float(err[0]), prev_error=float(prev[0]), iters=int(iters[0]), status=int(status[0]),
                src=src_ds[:ns].copy(), tgt=tgt_ds[:nt].copy(), normals=normals[:nt].copy(),
                matches=matches[:trace_iters, :ns].copy())


def icp_last_stats():
    """Work counters + per-kernel device times of the most recent registration call."""
    st = np.zeros(8, dtype=np.int64)
    check(_lib.load().icpb200_icp_last_stats(_ptr(st, c_int64_p)), "icpb200_icp_last_stats")
    return dict(sweep_pair_evals=int(st[0]), fp64_rescans=int(st[1]), iterations=int(st[2]), points_swept=int(st[3]),
                points_carried=int(st[4]), voxel_kernel_ns=int(st[5]), normals_kernel_ns=int(st[6]),
                pair_kernel_ns=int(st[7]))


def icp_phase_profile():
    """Thread 0's SM cycles per phase of a pair-kernel iteration, averaged over iterations >= 8 of the last call."""
    st = np.zeros(8, dtype=np.int64)
    check(_lib.load().icpb200_icp_phase_profile(_ptr(st, c_int64_p)), "icpb200_icp_phase_profile")
    n = max(int(st[5]), 1)
    names = ("classify", "nearest_neighbour", "accumulate", "solve", "apply_error")
    phases = {k: float(st[i]) / n for i, k in enumerate(names)}
    return dict(phases=phases, iterations_profiled=int(st[5]), cycles_per_iteration=float(sum(phases.values())) if st[5] else 0.0)


def icp_extra_stats():
    st = np.zeros(8, dtype=np.int64)
    check(_lib.load().icpb200_icp_extra_stats(_ptr(st, c_int64_p)), "icpb200_icp_extra_stats")
    return dict(far_field_iterations=int(st[0]), grid_queries=int(st[1]), grid_candidates=int(st[2]), grid_cells=int(st[3]),
                handed_over_by_class=[int(st[4]), int(st[5]), int(st[6])], helper_joins=int(st[7]))


def icp_pair_profile(n_pairs=0):
    """Per-pair counters of the last registration call: (n, 8) int64 = cycles, points swept, fp64 fallbacks, iterations,
    then cycles (iterations >= 8 only) of the classify phase, the nearest-neighbour phase and the rest, one spare column.
    The first call only switches the counters on (returns an empty array)."""
    out = np.zeros((max(int(n_pairs), 1), 8), dtype=np.int64)
    n = _lib.load().icpb200_icp_pair_profile(_ptr(out, c_int64_p), int(n_pairs))
    if n < 0:
        check(n, "icpb200_icp_pair_profile")
    return out[:n]


def voxel_downsample(points, voxel_size):
    pts = _f64(points)
    out = np.empty_like(pts)
    n_out = ctypes.c_int64(0)
    rc = _lib.load().icpb200_voxel_downsample(_ptr(pts, c_double_p), pts.shape[0], pts.shape[1], float(voxel_size),
                                              _ptr(out, c_double_p), ctypes.byref(n_out))
    check(rc, "icpb200_voxel_downsample")
    return out[:int(n_out.value)].copy()


def rotation_scores(sources, targets, angle_lists, shifts, want_nn=False):
    """score(angle) = mean_i min_j |R(angle) s_i + shift - t_j|^2 for every angle of every problem
    (features.py:205-211, slam.py:138-143); one launch for all problems and angles.

    Returns a list of score arrays, plus (dist, idx) lists when ``want_nn`` (one angle per problem)."""
    from .synth import pack_ragged
    src, so = pack_ragged([_f64(s).reshape(-1, 2) for s in sources])
    tgt, to = pack_ragged([_f64(t).reshape(-1, 2) for t in targets])
    angs = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in angle_lists]
    ao = np.zeros(len(angs) + 1, dtype=np.int64)
    ao[1:] = np.cumsum([len(a) for a in angs])
    ang = np.concatenate(angs) if angs else np.zeros(0)
    sh = _f64(shifts).reshape(-1, 2)
    if not (len(sources) == len(targets) == len(angs) == len(sh)):
        raise ValueError("rotation_scores: one target, angle list and shift per source")
    scores = np.empty(int(ao[-1]), dtype=np.float64)
    nn_d = np.empty(int(so[-1]), dtype=np.float64) if want_nn else None
    nn_i = np.empty(int(so[-1]), dtype=np.int32) if want_nn else None
    check(_lib.load().icpb200_rotation_scores(len(angs), _ptr(src, c_double_p), _ptr(so, c_int64_p), _ptr(tgt, c_double_p),
                                              _ptr(to, c_int64_p), _ptr(ang, c_double_p), _ptr(ao, c_int64_p),
                                              _ptr(sh, c_double_p), _ptr(scores, c_double_p),
                                              _ptr(nn_d, c_double_p) if want_nn else None,
                                              _ptr(nn_i, c_int32_p) if want_nn else None), "icpb200_rotation_scores")
    out = [scores[ao[k]:ao[k + 1]] for k in range(len(angs))]
    if not want_nn:
        return out
    return out, [nn_d[so[k]:so[k + 1]] for k in range(len(angs))], [nn_i[so[k]:so[k + 1]] for k in range(len(angs))]


class pinned:
    """Page-lock numpy arrays the caller reuses from call to call (full-rate PCIe copies).

    ``with pinned(flat, out): ...`` or keep the object and call ``release()``; the arrays must stay
    alive and must not be resized while pinned."""

    def __init__(self, *arrays):
        self._lib = _lib.load()
        self._ptrs = []
        for a in arrays:
            if not isinstance(a, np.ndarray) or not a.flags["C_CONTIGUOUS"] or a.nbytes == 0:
                continue
            check(self._lib.icpb200_pin_host(ctypes.c_void_p(a.ctypes.data), a.nbytes), "icpb200_pin_host")
            self._ptrs.append(a.ctypes.data)

    def release(self):
        for p in self._ptrs:
            self._lib.icpb200_unpin_host(ctypes.c_void_p(p))
        self._ptrs = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.release()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class DeviceGrid:
    """Owner of one device-resident occupancy grid handle."""

    def __init__(self, nx, ny, min_x, min_y, resolution, l_hit, l_miss, lo_min, lo_max):
        self.nx, self.ny = int(nx), int(ny)
        self.sharded = False
        self.peers_attached = False
        self.last_tiles_copied = 0
        self._lib = _lib.load()
        self._h = self._lib.icpb200_grid_create(self.nx, self.ny, float(min_x), float(min_y), float(resolution),
                                                float(l_hit), float(l_miss), float(lo_min), float(lo_max))
        if not self._h:
            raise RuntimeError(f"icpb200_grid_create failed: {_lib.last_error()}")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.icpb200_grid_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_shard(self, rank, world):
        self.sharded = int(world) > 1
        check(self._lib.icpb200_grid_set_shard(self._h, int(rank), int(world)), "icpb200_grid_set_shard")

    def update(self, origins, hits, hit_off):
        org = _f64(origins).reshape(-1, 2)
        pts = _f64(hits).reshape(-1, 2)
        off = np.ascontiguousarray(hit_off, dtype=np.int64)
        check(self._lib.icpb200_grid_update(self._h, len(off) - 1, _ptr(org, c_double_p), _ptr(pts, c_double_p),
                                            _ptr(off, c_int64_p)), "icpb200_grid_update")

    def rebuild(self, poses, local_points, off):
        """slam.py:271-277 in one call: clear, then replay scans given in their local frames with 3x3 poses."""
        P = _f64(poses).reshape(-1, 9)
        pts = _f64(local_points).reshape(-1, 2)
        o = np.ascontiguousarray(off, dtype=np.int64)
        if len(o) - 1 != len(P):
            raise ValueError("rebuild: one pose per scan")
        check(self._lib.icpb200_grid_rebuild(self._h, len(P), _ptr(P, c_double_p), _ptr(pts, c_double_p), _ptr(o, c_int64_p)),
              "icpb200_grid_rebuild")

    def update_dev(self, n_scans, d_origins, d_hits, d_hit_off, total_hits, stream=0):
        check(self._lib.icpb200_grid_update_dev(self._h, int(n_scans), d_origins, d_hits, d_hit_off, int(total_hits),
                                                stream), "icpb200_grid_update_dev")

    def read(self, out=None):
        if out is None:
            out = np.empty((self.ny, self.nx), dtype=np.float32)
        check(self._lib.icpb200_grid_read(self._h, _ptr(out, c_float_p)), "icpb200_grid_read")
        return out

    VIEWS = {"log_odds": 0, "probability": 1, "display": 2}
    UNTOUCHED = {"log_odds": 0.0, "probability": 0.5, "display": 1.0}      # the view of an unexplored cell

    def read_view(self, view="log_odds", out=None, dirty_only=False):
        """(ny, nx) float32 view of the map computed on the device (mapping.py:150-160).  With ``dirty_only`` only the
        tiles touched since the last reset are copied into ``out``, which must hold ``UNTOUCHED[view]`` elsewhere."""
        if out is None:
            out = np.full((self.ny, self.nx), self.UNTOUCHED[view], dtype=np.float32)
        n = ctypes.c_int32(0)
        check(self._lib.icpb200_grid_read_view(self._h, self.VIEWS[view], 1 if dirty_only else 0, _ptr(out, c_float_p),
                                               ctypes.byref(n)), "icpb200_grid_read_view")
        self.last_tiles_copied = int(n.value)
        return out

    def read_dirty(self, out):
        """Log-odds of the touched tiles into ``out`` (zeros elsewhere, kept by the caller)."""
        return self.read_view("log_odds", out, dirty_only=True)

    def reset(self):
        check(self._lib.icpb200_grid_reset(self._h), "icpb200_grid_reset")

    def ipc_export(self):
        buf = ctypes.create_string_buffer(128)
        check(self._lib.icpb200_grid_ipc_export(self._h, buf), "icpb200_grid_ipc_export")
        return buf.raw

    def ipc_attach(self, world, rank, handles):
        assert len(handles) == 128 * int(world)
        check(self._lib.icpb200_grid_ipc_attach(self._h, int(world), int(rank), bytes(handles)), "icpb200_grid_ipc_attach")
        self.peers_attached = True

    def push_tiles(self, stream=0):
        check(self._lib.icpb200_grid_push_tiles(self._h, stream), "icpb200_grid_push_tiles")

    def device_ptr(self):
        return int(self._lib.icpb200_grid_device_ptr(self._h) or 0)

    def last_stats(self):
        st = np.zeros(4, dtype=np.int64)
        check(self._lib.icpb200_grid_last_stats(self._h, _ptr(st, c_int64_p)), "icpb200_grid_last_stats")
        return dict(rays=int(st[0]), traversed=int(st[1]), hits=int(st[2]), runs=int(st[3]))


class DeviceSubmap:
    """The rolling window of global-frame scans (slam.py:559-562), resident on the device.

    ``append`` / ``pop`` / ``clear`` / ``len`` mirror the list the reference keeps (``submap_buffer``); ``build(voxel)`` is
    ``_build_submap`` (slam.py:103-108: vstack + voxel_downsample) and ``icp(...)`` is ``ICP(scan, build(voxel), ...)``
    (slam.py:217-225) without the 52k-point target ever crossing PCIe again: the window's two downsamples and its hash
    grid are computed on the device and cached until the window changes."""

    def __init__(self, capacity_scans=40, dim=2):
        self._lib = _lib.load()
        self.dim, self.capacity = int(dim), int(capacity_scans)
        self._h = self._lib.icpb200_submap_create(self.dim, self.capacity)
        if not self._h:
            raise RuntimeError(f"icpb200_submap_create failed: {_lib.last_error()}")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.icpb200_submap_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def append(self, points):
        pts = _f64(points).reshape(-1, self.dim)
        check(self._lib.icpb200_submap_push(self._h, _ptr(pts, c_double_p), pts.shape[0]), "icpb200_submap_push")

    def clear(self):
        check(self._lib.icpb200_submap_clear(self._h), "icpb200_submap_clear")

    def size(self):
        a, b = ctypes.c_int64(0), ctypes.c_int64(0)
        check(self._lib.icpb200_submap_size(self._h, ctypes.byref(a), ctypes.byref(b)), "icpb200_submap_size")
        return int(a.value), int(b.value)

    def __len__(self):
        return self.size()[0]

    def build(self, voxel_size, fetch=True):
        """``voxel_downsample(np.vstack(window), voxel_size)``; ``fetch=False`` only returns the row count."""
        n = ctypes.c_int64(0)
        cap = self.size()[1]
        out = np.empty((cap, self.dim)) if fetch else None
        check(self._lib.icpb200_submap_build(self._h, float(voxel_size), _ptr(out, c_double_p) if fetch else None, cap,
                                             ctypes.byref(n)), "icpb200_submap_build")
        return out[:int(n.value)].copy() if fetch else int(n.value)

    def icp(self, sources, submap_voxel, error_threshold, max_iterations, voxel_size, R_init=None, t_init=None,
            method="point_to_point", normal_k=10, max_corr_dist=None, nn_mode="auto"):
        """Register every cloud of ``sources`` (a list, or one (N, dim) array) onto the window's submap."""
        single = isinstance(sources, np.ndarray) and sources.ndim == 2
        clouds = [sources] if single else list(sources)
        from .synth import pack_ragged
        src, off = pack_ragged([_f64(s) for s in clouds], self.dim)
        n = len(clouds)
        r0, t0 = _init_arrays(R_init, t_init, n, self.dim)
        R, t, err, prev, iters, status = _alloc_out(n, self.dim)
        check(self._lib.icpb200_submap_icp(
            self._h, float(submap_voxel), n, _ptr(src, c_double_p), _ptr(off, c_int64_p), _ptr(r0, c_double_p), _ptr(t0, c_double_p),
            float(error_threshold), int(max_iterations), float(voxel_size), _method_code(method), int(normal_k),
            -1.0 if max_corr_dist is None else float(max_corr_dist), NN_MODES[nn_mode],
            _ptr(R, c_double_p), _ptr(t, c_double_p), _ptr(err, c_double_p), _ptr(prev, c_double_p), _ptr(iters, c_int32_p),
            _ptr(status, c_int32_p)), "icpb200_submap_icp")
        return dict(R=R, t=t, error=err, prev_error=prev, iters=iters, status=status)
