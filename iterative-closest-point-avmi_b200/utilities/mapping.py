"""GPU-backed ``utilities.mapping.OccupancyGrid2D`` -- same call surface as
the reference (/root/reference/utilities/mapping.py:13-187).

The log-odds grid lives on the GPU; ``update_scan`` / ``reset`` enqueue work
in libicp_b200.so and ``log_odds`` copies the (ny, nx) float32 array back on
first access after a change.  pyvista is imported only when a display method
is called (the reference imports it at module top, mapping.py:2, which breaks
headless use).
"""
import os
import sys

import numpy as np

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)

from icp_b200 import api as _api          # noqa: E402

UNEXPLORED = 0.0


class OccupancyGrid2D:
    """2-D probabilistic occupancy grid, log-odds ray tracing on the GPU.

    ``log_odds[iy, ix]`` has shape (ny, nx): > 0 occupied, < 0 free, 0 unexplored.
    """

    def __init__(self, min_x, max_x, min_y, max_y, resolution=0.1, p_hit=0.7,
                 p_miss=0.4, log_odds_min=-5.0, log_odds_max=5.0):
        self.min_x, self.max_x = float(min_x), float(max_x)
        self.min_y, self.max_y = float(min_y), float(max_y)
        self.resolution = float(resolution)
        # mapping.py:44-45, 49-52 -- computed on the host exactly as the reference does
        self.nx = int(np.ceil((self.max_x - self.min_x) / self.resolution))
        self.ny = int(np.ceil((self.max_y - self.min_y) / self.resolution))
        self.l_hit = np.log(p_hit / (1.0 - p_hit))
        self.l_miss = np.log(p_miss / (1.0 - p_miss))
        self.log_odds_min = float(log_odds_min)
        self.log_odds_max = float(log_odds_max)
        self._dev = _api.DeviceGrid(self.nx, self.ny, self.min_x, self.min_y, self.resolution,
                                    self.l_hit, self.l_miss, self.log_odds_min, self.log_odds_max)
        self._views = {}                  # view name -> page-locked host mirror + its state

    # ---- device <-> host -------------------------------------------------
    # One page-locked mirror per view (log-odds, probability, display value), kept between calls: a read-back copies only
    # the 64 x 64-cell tiles some update has touched since the last reset() -- a map is mostly unexplored -- and everything
    # else already holds the value of an untouched cell.
    def _read(self, view):
        entry = self._views.get(view)
        fill = _api.DeviceGrid.UNTOUCHED[view]
        if entry is None:
            arr = np.full((self.ny, self.nx), fill, dtype=np.float32)
            entry = dict(arr=arr, valid=False, refill=False, pin=_api.pinned(arr))
            self._views[view] = entry
        if not entry["valid"]:
            if entry["refill"]:                       # a reset un-touched tiles this mirror still shows
                entry["arr"][...] = fill
                entry["refill"] = False
            self._dev.read_view(view, entry["arr"], dirty_only=not self._dev.sharded)
            entry["valid"] = True
        return entry["arr"]

    def _invalidate(self, reset=False):
        for entry in self._views.values():
            entry["valid"] = False
            entry["refill"] = entry["refill"] or reset

    @property
    def log_odds(self):
        return self._read("log_odds")

    # ---- update ----------------------------------------------------------
    def update_scan(self, origin_xy, hit_points):
        """Trace rays from ``origin_xy`` (2,) to every row of ``hit_points`` (N, 2)."""
        pts = np.ascontiguousarray(hit_points, dtype=np.float64)
        if pts.size == 0:
            return
        org = np.ascontiguousarray(origin_xy, dtype=np.float64).reshape(1, 2)
        self._dev.update(org, pts.reshape(-1, 2), np.array([0, pts.shape[0]], dtype=np.int64))
        self._invalidate()

    def update_scans(self, origins, hit_clouds):
        """Batch form (the _rebuild_map replay, slam.py:271-277): scans applied in order."""
        from icp_b200.synth import pack_ragged
        flat, off = pack_ragged([np.asarray(h, dtype=np.float64).reshape(-1, 2) for h in hit_clouds])
        self._dev.update(np.asarray(origins, dtype=np.float64), flat, off)
        self._invalidate()

    def rebuild(self, scan_history):
        """The reference's `_rebuild_map(mapper, scan_history)` (slam.py:271-277) as one device call: clear the grid and
        replay every `(local_points (N, 2), pose (3, 3))` of the history in order; `transform_points_2d` (slam.py:46-50)
        runs on the device with numpy's roundings.  Same map as `reset()` + `update_scan(pose[:2, 2], pts @ R.T + t)` per
        scan."""
        from icp_b200.synth import pack_ragged
        scans = [np.asarray(pts, dtype=np.float64).reshape(-1, 2) for pts, _ in scan_history]
        poses = np.asarray([np.asarray(pose, dtype=np.float64) for _, pose in scan_history], dtype=np.float64).reshape(-1, 3, 3)
        if len(scans) == 0:
            self.reset()
            return
        flat, off = pack_ragged(scans)
        self._dev.rebuild(poses, flat, off)
        self._invalidate(reset=True)

    def reset(self):
        """Back to unexplored (all zeros)."""
        self._dev.reset()
        self._invalidate(reset=True)

    # ---- probability / display ---------------------------------------------
    def to_probability(self):
        """mapping.py:150-153 evaluated on the device in float32, as numpy evaluates it (the device exp differs from
        numpy's float32 exp by at most 2 ulp: values agree to 1e-6).  A fresh array, as the reference returns."""
        return self._read("probability").copy()

    def to_display(self):
        """mapping.py:155-160 on the device: occupied 1 - p, unexplored 1 (white), free 0.85 (light grey)."""
        return self._read("display").copy()

    def _flat_cell_data(self):
        return self.to_display().ravel(order="C")

    def create_pyvista_grid(self):
        import pyvista as pv
        grid = pv.ImageData(dimensions=(self.nx + 1, self.ny + 1, 1),
                            spacing=(self.resolution, self.resolution, 1e-6),
                            origin=(self.min_x, self.min_y, 0.0))
        grid.cell_data["occ"] = self._flat_cell_data()
        return grid

    def update_pyvista_grid(self, grid):
        grid.cell_data["occ"] = self._flat_cell_data()

    # ---- export --------------------------------------------------------------
    def save_csv(self, file_path):
        np.savetxt(file_path, self.to_probability(), delimiter=",")

    def save_npy(self, file_path):
        np.save(file_path, self.to_probability())
