// Host build of the integer / small-matrix device helpers, for CPU unit tests.
// The same headers are compiled into the CUDA kernels; here g++ compiles them
// as plain C++ so tests/test_bres_host.py and tests/test_linalg_host.py can
// check them against the oracle and numpy without a GPU.
#include <math.h>
#include <stdint.h>
#include <vector>

#include "bres.cuh"
#include "linalg_small.cuh"

using namespace icpb;

extern "C" {

// Walk a ray through for_each_tile_run + RunWalker exactly as the tile kernel
// does; emit (x, y, tile) for every in-grid free cell.  Returns the count.
int64_t harness_ray_cells(int ts, int ox, int oy, int hx, int hy, int nx, int ny,
                          int32_t* out_xyt, int64_t cap) {
    const RayGeom g = make_ray(ox, oy, hx, hy);
    int64_t n = 0;
    auto emit = [&](const TileRun& r) {
        int x, y;
        cell_at(g, r.n0, r.j0, x, y);
        RunWalker wk;
        wk.start(g, r.n0, r.j0);
        const int smaj = g.smaj, smin = g.smin;
        for (int c = 0; c < r.len; ++c) {
            if (n < cap) { out_xyt[3 * n] = x; out_xyt[3 * n + 1] = y; out_xyt[3 * n + 2] = r.tile; }
            ++n;
            const int m = wk.step();
            if (g.xmajor) { x += smaj; if (m) y += smin; }
            else          { y += smaj; if (m) x += smin; }
        }
    };
    const int tiles_x = (nx + ts - 1) / ts;
    if (ts == 64) for_each_tile_run<64>(g, nx, ny, tiles_x, emit);
    else if (ts == 8) for_each_tile_run<8>(g, nx, ny, tiles_x, emit);
    else if (ts == 4) for_each_tile_run<4>(g, nx, ny, tiles_x, emit);
    else return -1;
    return n;
}


// Per-tile load statistics of a batch of scans (mirrors occ_bin): for each
// tile the number of runs, traversed cells and distinct scans touching it.
// cells_out / runs_out / scans_out have tiles_x * tiles_y entries.
void harness_tile_stats(int n_scans, const double* origins, const double* hits, const int64_t* hit_off,
                        double min_x, double min_y, double res, int nx, int ny,
                        int64_t* cells_out, int64_t* runs_out, int64_t* scans_out, int64_t* maxrun_out) {
    const int tiles_x = (nx + 63) / 64, tiles_y = (ny + 63) / 64;
    std::vector<int> last(tiles_x * tiles_y, -1);
    std::vector<int64_t> cur(tiles_x * tiles_y, 0);
    for (int s = 0; s < n_scans; ++s) {
        const int ox = sat_cell(floor((origins[2 * s] - min_x) / res));
        const int oy = sat_cell(floor((origins[2 * s + 1] - min_y) / res));
        for (int64_t r = hit_off[s]; r < hit_off[s + 1]; ++r) {
            const int hx = sat_cell(floor((hits[2 * r] - min_x) / res));
            const int hy = sat_cell(floor((hits[2 * r + 1] - min_y) / res));
            const RayGeom g = make_ray(ox, oy, hx, hy);
            for_each_tile_run<64>(g, nx, ny, tiles_x, [&](const TileRun& t) {
                cells_out[t.tile] += t.len;
                runs_out[t.tile] += 1;
                if (last[t.tile] != s) { last[t.tile] = s; scans_out[t.tile] += 1; cur[t.tile] = 0; }
                cur[t.tile] += 1;
                if (cur[t.tile] > maxrun_out[t.tile]) maxrun_out[t.tile] = cur[t.tile];
            });
        }
    }
}

// Same walk through the iterator form (TileRunIter) used by the binning kernel.
int64_t harness_ray_runs_iter(int ox, int oy, int hx, int hy, int nx, int ny, int32_t* out4, int64_t cap) {
    TileRunIter<64> it;
    it.init(make_ray(ox, oy, hx, hy), nx, ny, (nx + 63) / 64);
    TileRun r;
    int64_t n = 0;
    while (it.next(r)) {
        if (n < cap) { out4[4 * n] = r.tile; out4[4 * n + 1] = r.n0; out4[4 * n + 2] = r.j0; out4[4 * n + 3] = r.len; }
        ++n;
    }
    return n;
}

int64_t harness_ray_runs_cb(int ox, int oy, int hx, int hy, int nx, int ny, int32_t* out4, int64_t cap) {
    int64_t n = 0;
    for_each_tile_run<64>(make_ray(ox, oy, hx, hy), nx, ny, (nx + 63) / 64, [&](const TileRun& r) {
        if (n < cap) { out4[4 * n] = r.tile; out4[4 * n + 1] = r.n0; out4[4 * n + 2] = r.j0; out4[4 * n + 3] = r.len; }
        ++n;
    });
    return n;
}

int harness_minor_steps(int ox, int oy, int hx, int hy, int n) {
    const RayGeom g = make_ray(ox, oy, hx, hy);
    return minor_steps(g, n);
}

int harness_sat_cell(double c) { return sat_cell(c); }

int harness_solve3(const double* a, const double* b, double* x) { return solve3_lu(a, b, x); }
void harness_kabsch2(const double* w, double* r) { kabsch2(w, r); }
void harness_kabsch3(const double* w, double* r) { kabsch3(w, r); }
void harness_eigvec2(double a, double b, double c, double* n) { sym2_min_eigvec(a, b, c, n); }

}  // extern "C"
