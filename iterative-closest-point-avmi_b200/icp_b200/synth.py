"""Deterministic synthetic 2-D lidar worlds, trajectories and scans.

The reference ships no lidar data (its ``data/`` directory is git-ignored,
``config.yaml:6``), so every benchmark and parity input is generated here from
a seed.  Shapes follow SURVEY.md §8(d): 1080 beams over a 270° field of view,
30 m maximum range, N(0, 0.01 m) range noise, ≈0.08 m / ≈0.015 rad of motion
per scan.  Beams with no return are dropped, as the reference's reader drops
all-zero rows (``services/lidar_service.py:17-18``).

Pure numpy; runs identically in the CPU container and on the GPU box.
"""
from __future__ import annotations

import numpy as np

N_BEAMS = 1080
FOV_DEG = 270.0
MAX_RANGE = 30.0
RANGE_SIGMA = 0.01


# ---------------------------------------------------------------------------
# worlds: arrays of wall segments, shape (S, 2, 2) = [segment][endpoint][xy]
# ---------------------------------------------------------------------------
def _box(x0, y0, x1, y1):
    c = [(x0, y0), (x1, y0), (x1, y1), (x0, y1)]
    return [(c[i], c[(i + 1) % 4]) for i in range(4)]


def room_world():
    """40 m x 24 m room with four box obstacles (the survey's probe world)."""
    segs = _box(-20.0, -12.0, 20.0, 12.0)
    segs += _box(-12.0, -6.0, -8.0, -2.5)
    segs += _box(6.0, 3.0, 11.0, 6.5)
    segs += _box(-3.0, 4.5, 1.5, 8.0)
    segs += _box(9.0, -8.5, 13.0, -5.0)
    return np.asarray(segs, dtype=np.float64)


def campus_world():
    """180 m x 180 m block of 4 x 4 buildings (30 m) separated by 12 m streets."""
    segs = _box(-90.0, -90.0, 90.0, 90.0)
    for i in range(4):
        for j in range(4):
            x0 = -78.0 + 42.0 * i
            y0 = -78.0 + 42.0 * j
            segs += _box(x0, y0, x0 + 30.0, y0 + 30.0)
    return np.asarray(segs, dtype=np.float64)


# ---------------------------------------------------------------------------
# trajectories: (n, 3) arrays of x, y, heading
# ---------------------------------------------------------------------------
def _rounded_rect_path(hx, hy, radius, step, n, start=0.0):
    """Poses along a rounded rectangle (half-extents hx, hy), arc-length step."""
    sx, sy = 2.0 * (hx - radius), 2.0 * (hy - radius)
    arc = 0.5 * np.pi * radius
    pieces = [("line", sx), ("arc", arc), ("line", sy), ("arc", arc),
              ("line", sx), ("arc", arc), ("line", sy), ("arc", arc)]
    perimeter = sum(p[1] for p in pieces)
    out = np.empty((n, 3))
    for k in range(n):
        s = (start + k * step) % perimeter
        # walk the pieces, counter-clockwise starting on the bottom edge
        x, y, th = -(hx - radius), -hy, 0.0
        for kind, length in pieces:
            adv = min(s, length)
            if kind == "line":
                x += adv * np.cos(th)
                y += adv * np.sin(th)
            else:
                cx = x - radius * np.sin(th)
                cy = y + radius * np.cos(th)
                dth = adv / radius
                th2 = th + dth
                x = cx + radius * np.sin(th2)
                y = cy - radius * np.cos(th2)
                th = th2
            s -= adv
            if s <= 0.0:
                break
        out[k] = (x, y, th)
    return out


def room_trajectory(n, seed=7, step=0.08):
    """Smooth loop inside ``room_world`` with a small seeded wobble."""
    rng = np.random.default_rng(seed)
    base = _rounded_rect_path(16.0, 9.6, 5.3, step, n, start=rng.uniform(0, 10))
    # low-frequency lateral / heading wobble keeps consecutive scans from
    # being exact rigid copies of each other
    ph = rng.uniform(0, 2 * np.pi, size=3)
    k = np.arange(n)
    base[:, 0] += 0.35 * np.sin(0.011 * k + ph[0])
    base[:, 1] += 0.30 * np.sin(0.013 * k + ph[1])
    base[:, 2] += 0.10 * np.sin(0.017 * k + ph[2])
    return base


def campus_trajectory(n, seed=7, step=0.08):
    """Loop along the street centre lines around the four middle buildings."""
    rng = np.random.default_rng(seed)
    base = _rounded_rect_path(42.0, 42.0, 4.0, step, n, start=rng.uniform(0, 30))
    ph = rng.uniform(0, 2 * np.pi, size=3)
    k = np.arange(n)
    base[:, 0] += 0.6 * np.sin(0.009 * k + ph[0])
    base[:, 1] += 0.6 * np.sin(0.012 * k + ph[1])
    base[:, 2] += 0.08 * np.sin(0.015 * k + ph[2])
    return base


# ---------------------------------------------------------------------------
# ray casting
# ---------------------------------------------------------------------------
def cast_scan(world, pose, rng, n_beams=N_BEAMS, fov_deg=FOV_DEG,
              max_range=MAX_RANGE, sigma=RANGE_SIGMA):
    """One scan in the sensor frame: (n_returns, 2) float64, beam order kept."""
    x, y, th = pose
    ang = np.deg2rad(np.linspace(-0.5 * fov_deg, 0.5 * fov_deg, n_beams))
    d = np.stack([np.cos(ang + th), np.sin(ang + th)], axis=1)      # (B, 2)
    a = world[:, 0, :]                                              # (S, 2)
    e = world[:, 1, :] - a                                          # (S, 2)
    w = a - np.array([x, y])                                        # (S, 2)
    den = d[:, None, 0] * e[None, :, 1] - d[:, None, 1] * e[None, :, 0]
    with np.errstate(divide="ignore", invalid="ignore"):
        r = (w[None, :, 0] * e[None, :, 1] - w[None, :, 1] * e[None, :, 0]) / den
        u = (w[None, :, 0] * d[:, None, 1] - w[None, :, 1] * d[:, None, 0]) / den
    ok = (np.abs(den) > 1e-12) & (r > 1e-6) & (u >= 0.0) & (u <= 1.0)
    r = np.where(ok, r, np.inf).min(axis=1)
    noise = rng.normal(0.0, sigma, size=n_beams)
    keep = r <= max_range
    r = r[keep] + noise[keep]
    la = ang[keep]
    return np.stack([r * np.cos(la), r * np.sin(la)], axis=1)


def make_sequence(n_scans, world="room", seed=0, traj_seed=7):
    """Return (scans, poses): list of (Ni, 2) sensor-frame clouds and (n, 3) poses."""
    if world == "room":
        segs, poses = room_world(), room_trajectory(n_scans, traj_seed)
    elif world == "campus":
        segs, poses = campus_world(), campus_trajectory(n_scans, traj_seed)
    else:
        raise ValueError(f"unknown world {world!r}")
    rng = np.random.default_rng(seed)
    scans = [cast_scan(segs, poses[i], rng) for i in range(n_scans)]
    return scans, poses


def to_world_frame(scan, pose):
    """Sensor-frame points -> world frame (what slam.py feeds update_scan)."""
    x, y, th = pose
    c, s = np.cos(th), np.sin(th)
    rot = np.array([[c, -s], [s, c]])
    return scan @ rot.T + np.array([x, y])


def loop_closure_pairs(poses, n_pairs, seed=0, max_dist=3.0, min_gap=1):
    """Index pairs (i, j), i != j, with |pos_i - pos_j| < max_dist (config.yaml:70)."""
    rng = np.random.default_rng(seed)
    n = len(poses)
    out = np.empty((n_pairs, 2), dtype=np.int64)
    got = 0
    while got < n_pairs:
        i = rng.integers(0, n, size=4 * (n_pairs - got))
        j = rng.integers(0, n, size=i.size)
        dist = np.hypot(*(poses[i, :2] - poses[j, :2]).T)
        ok = (dist < max_dist) & (np.abs(i - j) >= min_gap)
        take = min(int(ok.sum()), n_pairs - got)
        out[got:got + take, 0] = i[ok][:take]
        out[got:got + take, 1] = j[ok][:take]
        got += take
    return out


def submap_cloud(n_raw=52000, seed=3, extent=60.0, spacing=0.04, sigma=0.01):
    """Random axis-aligned wall segments sampled every ``spacing`` m."""
    rng = np.random.default_rng(seed)
    chunks, total = [], 0
    while total < n_raw:
        length = rng.uniform(3.0, 25.0)
        c = rng.uniform(-extent, extent, size=2)
        s = np.arange(0.0, length, spacing)
        pts = np.zeros((s.size, 2))
        if rng.random() < 0.5:
            pts[:, 0], pts[:, 1] = c[0] + s, c[1]
        else:
            pts[:, 0], pts[:, 1] = c[0], c[1] + s
        pts += rng.normal(0.0, sigma, size=pts.shape)
        chunks.append(pts)
        total += s.size
    return np.vstack(chunks)[:n_raw]


def pack_ragged(clouds, dim=2):
    """Concatenate clouds -> (flat (sum,dim) float64 C-contiguous, offsets int64 (n+1,))."""
    if len(clouds) == 1:                                        # one cloud (every ICP() call): no copy
        c = np.ascontiguousarray(clouds[0], dtype=np.float64)
        if c.ndim == 2 and c.shape[1] == dim:
            return c, np.array([0, len(c)], dtype=np.int64)
    off = np.zeros(len(clouds) + 1, dtype=np.int64)
    for i, c in enumerate(clouds):
        off[i + 1] = off[i] + len(c)
    flat = np.empty((int(off[-1]), dim), dtype=np.float64)
    for i, c in enumerate(clouds):
        flat[off[i]:off[i + 1]] = c
    return flat, off
