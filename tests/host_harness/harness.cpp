// Host build of the integer / small-matrix device helpers, for CPU unit tests.
// The same headers are compiled into the CUDA kernels; here g++ compiles them
// as plain C++ so tests/test_bres_host.py and tests/test_linalg_host.py can
// check them against the oracle and numpy without a GPU.
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
