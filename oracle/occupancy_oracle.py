"""CPU oracle for the occupancy-grid update -- TEST INFRASTRUCTURE ONLY.

Two restatements of /root/reference/utilities/mapping.py:28-145:

* ``GridOraclePy`` -- literal numpy / pure-Python loops (small cases only);
* ``GridOracleC``  -- ctypes front end of ``occupancy_oracle.c`` (fast; used
  for full-size parity and as the ``cpu_baseline`` "port").

Both are pinned bit-exact against the live reference by
``oracle/pin_against_reference.py``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module; the product path never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboccupancy_oracle.so")


def build(force=False):
    """Compile occupancy_oracle.c with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "occupancy_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboccupancy_oracle.so"])
    return _SO


def grid_shape(min_x, max_x, min_y, max_y, resolution):
    """mapping.py:44-45."""
    nx = int(np.ceil((float(max_x) - float(min_x)) / float(resolution)))
    ny = int(np.ceil((float(max_y) - float(min_y)) / float(resolution)))
    return nx, ny


def log_odds(p):
    """mapping.py:49-50."""
    return np.log(p / (1.0 - p))


def line_cells_py(x0, y0, x1, y1):
    """mapping.py:68-89: cells from (x0,y0) to (x1,y1), endpoint excluded."""
    out = []
    dx, dy = abs(x1 - x0), abs(y1 - y0)
    sx = 1 if x0 < x1 else -1
    sy = 1 if y0 < y1 else -1
    err = dx - dy
    x, y = x0, y0
    while not (x == x1 and y == y1):
        out.append((x, y))
        e2 = 2 * err
        if e2 > -dy:
            err -= dy
            x += sx
        if e2 < dx:
            err += dx
            y += sy
    return out


class _GridBase:
    def __init__(self, min_x, max_x, min_y, max_y, resolution=0.1, p_hit=0.7,
                 p_miss=0.4, log_odds_min=-5.0, log_odds_max=5.0):
        # mapping.py:28-52
        self.min_x, self.max_x = float(min_x), float(max_x)
        self.min_y, self.max_y = float(min_y), float(max_y)
        self.resolution = float(resolution)
        self.nx, self.ny = grid_shape(min_x, max_x, min_y, max_y, resolution)
        self.log_odds = np.zeros((self.ny, self.nx), dtype=np.float32)
        self.l_hit = log_odds(p_hit)
        self.l_miss = log_odds(p_miss)
        self.log_odds_min = float(log_odds_min)
        self.log_odds_max = float(log_odds_max)

    def reset(self):
        self.log_odds[:] = 0.0                                  # mapping.py:143-145


class GridOraclePy(_GridBase):
    """Literal restatement (numpy + Python loops) of mapping.py:103-141."""

    def update_scan(self, origin_xy, hit_points):
        hit_points = np.asarray(hit_points, dtype=np.float64)
        if hit_points.size == 0:
            return
        ox = int(np.floor((origin_xy[0] - self.min_x) / self.resolution))
        oy = int(np.floor((origin_xy[1] - self.min_y) / self.resolution))
        hx = np.floor((hit_points[:, 0] - self.min_x) / self.resolution).astype(int)
        hy = np.floor((hit_points[:, 1] - self.min_y) / self.resolution).astype(int)
        ok = (hx >= 0) & (hx < self.nx) & (hy >= 0) & (hy < self.ny)
        if ok.any():
            np.add.at(self.log_odds, (hy[ok], hx[ok]), self.l_hit)
        for i in range(len(hit_points)):
            for fx, fy in line_cells_py(ox, oy, int(hx[i]), int(hy[i])):
                if 0 <= fx < self.nx and 0 <= fy < self.ny:
                    self.log_odds[fy, fx] += self.l_miss
        np.clip(self.log_odds, self.log_odds_min, self.log_odds_max, out=self.log_odds)


class GridOracleC(_GridBase):
    """ctypes front end of occupancy_oracle.c."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            lib = ctypes.CDLL(build())
            dp = ctypes.POINTER(ctypes.c_double)
            i64 = ctypes.c_int64
            lib.occ_oracle_update.restype = i64
            lib.occ_oracle_update.argtypes = [
                ctypes.POINTER(ctypes.c_float), i64, i64,
                ctypes.c_double, ctypes.c_double, ctypes.c_double,
                ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                dp, dp, i64, ctypes.c_int]
            lib.occ_oracle_update_many.restype = i64
            lib.occ_oracle_update_many.argtypes = [
                ctypes.POINTER(ctypes.c_float), i64, i64,
                ctypes.c_double, ctypes.c_double, ctypes.c_double,
                ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                i64, dp, dp, ctypes.POINTER(i64), ctypes.c_int]
            lib.occ_oracle_ray_cells.restype = i64
            lib.occ_oracle_ray_cells.argtypes = [i64, i64, i64, i64,
                                                 ctypes.POINTER(i64), i64]
            cls._lib = lib
        return cls._lib

    def _clip_mode(self, fast):
        # the bounding-box clip is only equivalent when untouched cells are
        # already inside the clamp interval (true once 0 lies inside it)
        return 1 if (fast and self.log_odds_min <= 0.0 <= self.log_odds_max) else 0

    def _common(self):
        return (self.log_odds.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                self.nx, self.ny, self.min_x, self.min_y, self.resolution,
                float(self.l_hit), float(self.l_miss),
                self.log_odds_min, self.log_odds_max)

    def update_scan(self, origin_xy, hit_points, fast=False):
        pts = np.ascontiguousarray(hit_points, dtype=np.float64)
        org = np.ascontiguousarray(origin_xy, dtype=np.float64)
        dp = ctypes.POINTER(ctypes.c_double)
        return self.lib().occ_oracle_update(
            *self._common(), org.ctypes.data_as(dp), pts.ctypes.data_as(dp),
            pts.shape[0] if pts.size else 0, self._clip_mode(fast))

    def update_many(self, origins, hits_flat, hit_off, fast=True):
        org = np.ascontiguousarray(origins, dtype=np.float64)
        pts = np.ascontiguousarray(hits_flat, dtype=np.float64)
        off = np.ascontiguousarray(hit_off, dtype=np.int64)
        dp = ctypes.POINTER(ctypes.c_double)
        return self.lib().occ_oracle_update_many(
            *self._common(), len(off) - 1, org.ctypes.data_as(dp),
            pts.ctypes.data_as(dp), off.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
            self._clip_mode(fast))


def line_cells_c(x0, y0, x1, y1):
    n = max(abs(x1 - x0), abs(y1 - y0))
    buf = np.empty((max(n, 1), 2), dtype=np.int64)
    got = GridOracleC.lib().occ_oracle_ray_cells(
        x0, y0, x1, y1, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), len(buf))
    assert got == n
    return buf[:n]


# ---- map rebuild: /root/reference/slam.py:46-50 and 271-277 ----------------
def transform_points_2d(points_2d, pose):
    """slam.py:46-50 -- `points_2d @ R.T + t` with R, t taken from a 3x3 homogeneous pose, written as the reference
    writes it (numpy matmul, then the broadcast add)."""
    R = pose[:2, :2]
    t = pose[:2, 2]
    return points_2d @ R.T + t


def rebuild_map(grid, scan_history):
    """slam.py:271-277 `_rebuild_map`: clear the grid, replay every (local points, pose) of the history in order,
    origin = the pose's translation.  `grid` is any of the oracle grids above."""
    grid.reset()
    for pts, pose in scan_history:
        origin = pose[:2, 2]
        grid.update_scan(origin, transform_points_2d(pts, pose))
    return grid
