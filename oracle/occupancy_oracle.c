/*
 * CPU oracle for the occupancy-grid update -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain C restatement of the reference's OccupancyGrid2D.update_scan
 * (/root/reference/utilities/mapping.py:103-141) and its helpers, so that
 * full-size parity (2000 scans into 4096 x 4096) finishes in seconds.  Only
 * tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load
 * it.  It is pinned bit-exact against the live reference by
 * oracle/pin_against_reference.py before the golden fixtures are written.
 *
 * Arithmetic contract restated:
 *   - world -> cell: floor((w - min) / resolution) in float64, then to
 *     integer (mapping.py:57-60, 94-98);
 *   - the grid is float32; every add is x = (float)((double)x + c) with c a
 *     float64 log-odds increment (mapping.py:47-50, 129, 139);
 *   - per scan: all in-bounds hits first, in input order (np.add.at,
 *     mapping.py:124-129), then every ray's free cells (mapping.py:135-139),
 *     then one clip of the whole grid to [lo_min, lo_max] (mapping.py:141).
 *
 * Build: see oracle/Makefile (gcc -O2 -shared -fPIC; no -ffast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* mapping.py:68-89 -- integer line from (x0,y0) to (x1,y1), endpoint excluded.
 * Calls visit(ctx, x, y) for each cell in order; returns the number of cells. */
typedef void (*cell_fn)(void *ctx, int64_t x, int64_t y);

static int64_t walk_line(int64_t x0, int64_t y0, int64_t x1, int64_t y1,
                         cell_fn visit, void *ctx)
{
    int64_t dx = llabs(x1 - x0);
    int64_t dy = llabs(y1 - y0);
    int64_t sx = (x0 < x1) ? 1 : -1;
    int64_t sy = (y0 < y1) ? 1 : -1;
    int64_t err = dx - dy;
    int64_t x = x0, y = y0, n = 0;
    for (;;) {
        if (x == x1 && y == y1)
            break;
        visit(ctx, x, y);
        ++n;
        int64_t e2 = 2 * err;
        if (e2 > -dy) { err -= dy; x += sx; }
        if (e2 < dx)  { err += dx; y += sy; }
    }
    return n;
}

/* ---- single-ray cell listing, for Bresenham cell-set parity tests ---- */
struct cell_sink { int64_t *out; int64_t cap; int64_t n; };

static void sink_cell(void *ctx, int64_t x, int64_t y)
{
    struct cell_sink *s = (struct cell_sink *)ctx;
    if (s->n < s->cap) {
        s->out[2 * s->n] = x;
        s->out[2 * s->n + 1] = y;
    }
    s->n++;
}

/* Writes up to cap (x,y) pairs; returns the true cell count. */
int64_t occ_oracle_ray_cells(int64_t x0, int64_t y0, int64_t x1, int64_t y1,
                             int64_t *out_xy, int64_t cap)
{
    struct cell_sink s = { out_xy, cap, 0 };
    walk_line(x0, y0, x1, y1, sink_cell, &s);
    return s.n;
}

/* ---- grid update ---- */
struct grid_ctx {
    float *g;
    int64_t nx, ny;
    double l_miss;
    int64_t bx0, bx1, by0, by1;   /* bounding box of touched cells */
    int64_t touched;
};

static void touch(struct grid_ctx *c, int64_t x, int64_t y)
{
    if (x < c->bx0) c->bx0 = x;
    if (x > c->bx1) c->bx1 = x;
    if (y < c->by0) c->by0 = y;
    if (y > c->by1) c->by1 = y;
    c->touched++;
}

static void free_cell(void *ctx, int64_t x, int64_t y)
{
    struct grid_ctx *c = (struct grid_ctx *)ctx;
    if (x >= 0 && x < c->nx && y >= 0 && y < c->ny) {       /* mapping.py:138 */
        float *p = &c->g[y * c->nx + x];
        *p = (float)((double)*p + c->l_miss);               /* mapping.py:139 */
        touch(c, x, y);
    }
}

static int64_t to_cell(double w, double lo, double res)
{
    return (int64_t)floor((w - lo) / res);                  /* mapping.py:58-59, 96-97 */
}

/*
 * One update_scan.  clip_mode 0 = clip the whole grid (literal, mapping.py:141);
 * 1 = clip only the bounding box of cells touched by this scan (identical
 * result whenever every cell was already inside [lo_min, lo_max] on entry).
 * Returns the number of in-bounds free-cell updates (traversed cells).
 */
int64_t occ_oracle_update(float *grid, int64_t nx, int64_t ny,
                          double min_x, double min_y, double res,
                          double l_hit, double l_miss,
                          double lo_min, double lo_max,
                          const double *origin_xy,
                          const double *hits_xy, int64_t n_hits,
                          int clip_mode)
{
    if (n_hits == 0)                                        /* mapping.py:113-114 */
        return 0;
    struct grid_ctx c = { grid, nx, ny, l_miss, nx, -1, ny, -1, 0 };
    int64_t ox = to_cell(origin_xy[0], min_x, res);         /* mapping.py:116 */
    int64_t oy = to_cell(origin_xy[1], min_y, res);

    for (int64_t i = 0; i < n_hits; ++i) {                  /* mapping.py:124-129 */
        int64_t hx = to_cell(hits_xy[2 * i], min_x, res);
        int64_t hy = to_cell(hits_xy[2 * i + 1], min_y, res);
        if (hx >= 0 && hx < nx && hy >= 0 && hy < ny) {
            float *p = &grid[hy * nx + hx];
            *p = (float)((double)*p + l_hit);
            touch(&c, hx, hy);
        }
    }
    int64_t traversed = 0;
    for (int64_t i = 0; i < n_hits; ++i) {                  /* mapping.py:135-139 */
        int64_t hx = to_cell(hits_xy[2 * i], min_x, res);
        int64_t hy = to_cell(hits_xy[2 * i + 1], min_y, res);
        int64_t before = c.touched;
        walk_line(ox, oy, hx, hy, free_cell, &c);
        traversed += c.touched - before;
    }
    /* mapping.py:141 -- bounds are Python floats, so numpy clips in float32 */
    const float lo = (float)lo_min, hi = (float)lo_max;
    int64_t x0 = 0, x1 = nx - 1, y0 = 0, y1 = ny - 1;
    if (clip_mode == 1) { x0 = c.bx0; x1 = c.bx1; y0 = c.by0; y1 = c.by1; }
    for (int64_t y = y0; y <= y1; ++y) {
        float *row = grid + y * nx;
        for (int64_t x = x0; x <= x1; ++x) {
            float v = row[x];
            if (v < lo) v = lo;
            if (v > hi) v = hi;
            row[x] = v;
        }
    }
    return traversed;
}

/* Many scans in order (the _rebuild_map replay, /root/reference/slam.py:271-277). */
int64_t occ_oracle_update_many(float *grid, int64_t nx, int64_t ny,
                               double min_x, double min_y, double res,
                               double l_hit, double l_miss,
                               double lo_min, double lo_max,
                               int64_t n_scans, const double *origins_xy,
                               const double *hits_xy, const int64_t *hit_off,
                               int clip_mode)
{
    int64_t total = 0;
    for (int64_t s = 0; s < n_scans; ++s)
        total += occ_oracle_update(grid, nx, ny, min_x, min_y, res, l_hit, l_miss,
                                   lo_min, lo_max, origins_xy + 2 * s,
                                   hits_xy + 2 * hit_off[s],
                                   hit_off[s + 1] - hit_off[s], clip_mode);
    return total;
}
