// Closed-form Bresenham geometry used by the occupancy kernels.
//
// The reference walks each ray with an error accumulator
// (/root/reference/utilities/mapping.py:68-89).  For a tile-resident kernel a
// thread must be able to enter a ray at an arbitrary step, so the walk is
// restated in closed form.  With dmaj = max(|dx|,|dy|), dmin = min(|dx|,|dy|)
// the reference visits exactly dmaj cells n = 0 .. dmaj-1 (endpoint excluded);
// the major coordinate advances by one every step and after n steps the minor
// coordinate has advanced
//
//        J(n) = floor((2*n*dmin + dmaj - 1) / (2*dmaj))
//
// times (round-half-down of n*dmin/dmaj).  Derivation: the reference steps the
// minor axis at iteration n iff 2*err_n < dmaj (resp. > -dmaj) with
// err_n = dmaj - dmin - n*dmin + j*dmaj, i.e. iff j < ((2n+2)*dmin - dmaj)/(2*dmaj);
// both comparisons are strict, so the rule is symmetric in x and y.
// tests/test_bres_host.py checks J(n) and the tile splitter below, compiled
// for the host, against the oracle's cell lists exhaustively.
//
// Everything here is __host__ __device__ and integer-only.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ICPB_HD __host__ __device__ __forceinline__
#else
#define ICPB_HD inline
#endif

namespace icpb {

// Cell coordinates are saturated to +-2^27 so every product below fits int64,
// every sum of two coordinates fits int32 and the walker's error term
// (|d| <= 2*dmaj + 2*dmin) fits int32.  A hit 2^27 cells away from the grid is
// ~6,700 km at 5 cm; the reference's own int cast is undefined far beyond that.
constexpr int32_t kCellSat = 1 << 27;

ICPB_HD int32_t sat_cell(double c) {
    if (!(c > -(double)kCellSat)) return -kCellSat;      // also catches NaN
    if (c > (double)kCellSat) return kCellSat;
    return (int32_t)c;                                   // c is already floor()ed
}

struct RayGeom {
    int32_t ox, oy;        // origin cell
    int32_t smaj, smin;    // +-1 step along major / minor axis
    int32_t dmaj, dmin;    // |delta| along major / minor axis
    int32_t xmajor;        // 1: x is the major axis
};

ICPB_HD RayGeom make_ray(int32_t ox, int32_t oy, int32_t hx, int32_t hy) {
    RayGeom g;
    g.ox = ox; g.oy = oy;
    int32_t dx = hx > ox ? hx - ox : ox - hx;
    int32_t dy = hy > oy ? hy - oy : oy - hy;
    int32_t sx = ox < hx ? 1 : -1;                       // mapping.py:74
    int32_t sy = oy < hy ? 1 : -1;                       // mapping.py:75
    g.xmajor = dx >= dy;
    if (g.xmajor) { g.dmaj = dx; g.dmin = dy; g.smaj = sx; g.smin = sy; }
    else          { g.dmaj = dy; g.dmin = dx; g.smaj = sy; g.smin = sx; }
    return g;
}

// minor-axis advance after n major steps (0 <= n <= dmaj, dmaj >= 1)
ICPB_HD int32_t minor_steps(const RayGeom& g, int32_t n) {
    int64_t num = 2 * (int64_t)n * g.dmin + g.dmaj - 1;
    return (int32_t)(num / (2 * (int64_t)g.dmaj));
}

// smallest n >= 0 with minor_steps(n) >= j  (j >= 1, dmin >= 1)
ICPB_HD int64_t first_step_reaching(const RayGeom& g, int64_t j) {
    int64_t num = 2 * (int64_t)g.dmaj * j - g.dmaj + 1;        // >= 1
    int64_t den = 2 * (int64_t)g.dmin;
    return (num + den - 1) / den;
}

ICPB_HD void cell_at(const RayGeom& g, int32_t n, int32_t j, int32_t& x, int32_t& y) {
    if (g.xmajor) { x = g.ox + g.smaj * n; y = g.oy + g.smin * j; }
    else          { y = g.oy + g.smaj * n; x = g.ox + g.smin * j; }
}

// Half-open step range [lo, hi) of a coordinate c(n) = o + s*k(n), k
// non-decreasing in n, for which 0 <= c < size; k(n) = n for the major axis.
ICPB_HD void major_range(int32_t o, int32_t s, int32_t size, int64_t& lo, int64_t& hi) {
    if (s > 0) { lo = -(int64_t)o;                 hi = (int64_t)size - o; }
    else       { lo = (int64_t)o - (size - 1);     hi = (int64_t)o + 1; }
    if (lo < 0) lo = 0;
}

// Clip the ray's step range to the grid: returns [na, nb) (possibly empty).
ICPB_HD void clip_to_grid(const RayGeom& g, int32_t nx, int32_t ny, int64_t& na, int64_t& nb) {
    const int32_t size_maj = g.xmajor ? nx : ny, size_min = g.xmajor ? ny : nx;
    const int32_t omaj = g.xmajor ? g.ox : g.oy, omin = g.xmajor ? g.oy : g.ox;
    int64_t lo, hi;
    major_range(omaj, g.smaj, size_maj, lo, hi);
    na = lo; nb = hi < g.dmaj ? hi : g.dmaj;
    // minor axis: need jlo <= J(n) < jhi
    int64_t jlo, jhi;
    major_range(omin, g.smin, size_min, jlo, jhi);
    if (g.dmin == 0) {
        if (!(jlo <= 0 && 0 < jhi)) nb = na;             // never inside
        return;
    }
    if (jlo > 0) { int64_t n = first_step_reaching(g, jlo); if (n > na) na = n; }
    if (jhi <= 0) { nb = na; return; }
    { int64_t n = first_step_reaching(g, jhi); if (n < nb) nb = n; }
    if (nb < na) nb = na;
}

// One maximal run of consecutive ray cells inside a single TS x TS tile.
struct TileRun {
    int32_t tile;     // ty * tiles_x + tx
    int32_t n0;       // first step of the run
    int32_t j0;       // minor_steps(n0)
    int32_t len;      // number of cells, 1 .. TS
};

// Enumerate the tile runs of the in-grid part of a ray, in walk order.
// F is callable as f(const TileRun&).  TS must be a power of two.
template <int TS, class F>
ICPB_HD void for_each_tile_run(const RayGeom& g, int32_t nx, int32_t ny, int32_t tiles_x, F&& f) {
    if (g.dmaj == 0) return;                              // hit in the origin cell: no free cells
    int64_t na, nb;
    clip_to_grid(g, nx, ny, na, nb);
    int64_t n = na;
    while (n < nb) {
        const int32_t j = minor_steps(g, (int32_t)n);
        int32_t x, y;
        cell_at(g, (int32_t)n, j, x, y);
        const int32_t tx = x / TS, ty = y / TS;           // x, y >= 0 here
        const int32_t cmaj = g.xmajor ? x : y, cmin = g.xmajor ? y : x;
        // steps until the major coordinate leaves the tile
        const int32_t inmaj = cmaj & (TS - 1);
        int64_t end = n + (g.smaj > 0 ? (TS - inmaj) : (inmaj + 1));
        // steps until the minor coordinate leaves the tile
        if (g.dmin != 0) {
            const int32_t inmin = cmin & (TS - 1);
            const int64_t jleave = (int64_t)j + (g.smin > 0 ? (TS - inmin) : (inmin + 1));
            const int64_t nleave = first_step_reaching(g, jleave);
            if (nleave < end) end = nleave;
        }
        if (end > nb) end = nb;
        TileRun r;
        r.tile = ty * tiles_x + tx;
        r.n0 = (int32_t)n;
        r.j0 = j;
        r.len = (int32_t)(end - n);
        f(r);
        n = end;
    }
}

// Iterator form of for_each_tile_run (same runs, same order) for code that must
// keep all lanes of a warp converged while each walks its own ray.
template <int TS>
struct TileRunIter {
    RayGeom g;
    int64_t n, nb;
    int32_t tiles_x;
    ICPB_HD void init(const RayGeom& geom, int32_t nx, int32_t ny, int32_t tiles_x_) {
        g = geom; tiles_x = tiles_x_;
        n = 0; nb = 0;
        if (g.dmaj != 0) clip_to_grid(g, nx, ny, n, nb);
    }
    ICPB_HD void init_empty() { n = 0; nb = 0; }
    ICPB_HD bool next(TileRun& r) {
        if (n >= nb) return false;
        const int32_t j = minor_steps(g, (int32_t)n);
        int32_t x, y;
        cell_at(g, (int32_t)n, j, x, y);
        const int32_t tx = x / TS, ty = y / TS;
        const int32_t cmaj = g.xmajor ? x : y, cmin = g.xmajor ? y : x;
        const int32_t inmaj = cmaj & (TS - 1);
        int64_t end = n + (g.smaj > 0 ? (TS - inmaj) : (inmaj + 1));
        if (g.dmin != 0) {
            const int32_t inmin = cmin & (TS - 1);
            const int64_t jleave = (int64_t)j + (g.smin > 0 ? (TS - inmin) : (inmin + 1));
            const int64_t nleave = first_step_reaching(g, jleave);
            if (nleave < end) end = nleave;
        }
        if (end > nb) end = nb;
        r.tile = ty * tiles_x + tx;
        r.n0 = (int32_t)n;
        r.j0 = j;
        r.len = (int32_t)(end - n);
        n = end;
        return true;
    }
};

// Incremental walker used inside a tile: starts at (n0, j0) and reproduces the
// reference's minor-axis decisions without divisions.
struct RunWalker {
    int32_t d;          // (2n+2)*dmin - 2*dmaj*j - dmaj ; minor step taken iff d > 0
    int32_t inc_maj;    // 2*dmin
    int32_t dec_min;    // 2*dmaj
    ICPB_HD void start(const RayGeom& g, int32_t n0, int32_t j0) {
        inc_maj = 2 * g.dmin;
        dec_min = 2 * g.dmaj;
        // the products need 64 bits, the result is bounded by 2*dmaj + 2*dmin
        d = (int32_t)((2 * (int64_t)n0 + 2) * g.dmin - (int64_t)dec_min * j0 - g.dmaj);
    }
    // advance one major step; returns 1 if the minor coordinate also steps
    ICPB_HD int32_t step() {
        const int32_t m = d > 0;
        d += inc_maj - (m ? dec_min : 0);
        return m;
    }
};

}  // namespace icpb
