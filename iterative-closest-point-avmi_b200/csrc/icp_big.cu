// Kernels for clouds that do not fit one CTA's shared memory (scan-to-submap
// registration against a ~50k-point target, SURVEY.md section 2 kernels K1/K4):
//
//   big_voxel_kernel   voxel-grid mean of one cloud per CTA with the working
//                      set in global memory: fp64 cell keys -> stable LSD radix
//                      sort (8-bit digits, only as many passes as the key range
//                      needs) -> run heads -> input-order means.  Same result,
//                      bit for bit, as the shared-memory version
//                      (/root/reference/utilities/icp.py:117-129).
//   big_grid_kernel    uniform hash grid over one downsampled target per CTA:
//                      bucket histogram -> exclusive scan -> scatter
//                      (replaces the KDTree build, icp.py:173).
//   big_normals_kernel exact (k+1)-NN + 2x2 PCA normals on that grid, one thread
//                      per target point, many CTAs per cloud (icp.py:51-76).
//
// The per-pair loop for these targets is icp_pairs_kernel<2, true> (grid mode)
// in icp_kernel.cu.
#include <algorithm>

#include "icp_b200.h"
#include "common.cuh"
#include "icp_kernel.h"
#include "linalg_small.cuh"

namespace icpb {

constexpr int kBT = 1024;                 // threads per CTA for the per-cloud kernels
constexpr int kBW = kBT / 32;

struct BigShared {
    double red[kBW][8];
    unsigned hist[256];
    unsigned base[256];
    unsigned short wcnt[kBW][256];
    unsigned scan_w[kBW];
    unsigned carry;
    int n_out;
};

__device__ __forceinline__ void big_reduce_min(double* v, int nv, BigShared& sh) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    for (int k = 0; k < nv; ++k)
        for (int o = 16; o > 0; o >>= 1) v[k] = fmin(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
    __syncthreads();
    if (l == 0) for (int k = 0; k < nv; ++k) sh.red[w][k] = v[k];
    __syncthreads();
    for (int k = 0; k < nv; ++k) {
        double s = sh.red[0][k];
        for (int ww = 1; ww < kBW; ++ww) s = fmin(s, sh.red[ww][k]);
        v[k] = s;
    }
}

// exclusive scan of one unsigned per thread over the CTA; returns the prefix and the total
__device__ __forceinline__ unsigned big_scan(unsigned v, BigShared& sh, unsigned& total) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    unsigned inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (l >= o) inc += t;
    }
    __syncthreads();
    if (l == 31) sh.scan_w[w] = inc;
    __syncthreads();
    unsigned base = 0, tot = 0;
    for (int ww = 0; ww < kBW; ++ww) { const unsigned c = sh.scan_w[ww]; if (ww < w) base += c; tot += c; }
    total = tot;
    return base + inc - v;
}

// ---- voxel-grid mean, one CTA per cloud, global-memory scratch ---------------------
// scratch layout per cloud (at the cloud's raw offset): keys[2][n] (u64), idx[2][n] (u32)
template <int DIM>
__global__ void __launch_bounds__(kBT) big_voxel_kernel(const CloudSet cs, double voxel,
                                                        unsigned long long* key_buf, unsigned* idx_buf,
                                                        long long total_points) {
    __shared__ BigShared sh;
    const int c = blockIdx.x;
    if (cs.used && !cs.used[c]) return;
    const long long beg = cs.off[c];
    const int n = (int)(cs.off[c + 1] - beg);
    if (n <= 0) { if (threadIdx.x == 0) cs.ds_n[c] = -1; return; }
    const double* raw = cs.raw + beg * DIM;
    double* out = cs.ds + beg * DIM;
    unsigned long long* keys[2] = {key_buf + beg, key_buf + total_points + beg};
    unsigned* idx[2] = {idx_buf + beg, idx_buf + total_points + beg};
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // bounding box (icp.py:119)
    double b[2 * DIM];
    for (int a = 0; a < 2 * DIM; ++a) b[a] = INFINITY;
    for (int i = tid; i < n; i += kBT)
        for (int a = 0; a < DIM; ++a) {
            const double v = raw[(size_t)i * DIM + a];
            b[a] = fmin(b[a], v);
            b[DIM + a] = fmin(b[DIM + a], -v);
        }
    big_reduce_min(b, 2 * DIM, sh);
    double lo[DIM], hi[DIM];
    unsigned long long ext[DIM];
    double span = 1.0;
    for (int a = 0; a < DIM; ++a) {
        lo[a] = b[a]; hi[a] = -b[DIM + a];
        const double top = floor((hi[a] - lo[a]) / voxel);
        ext[a] = (unsigned long long)top + 1ull;
        span *= (top + 1.0);
    }
    if (!(span < 4.0e18)) { if (tid == 0) cs.ds_n[c] = -1; return; }
    unsigned long long kmax = 1ull;
    for (int a = 0; a < DIM; ++a) kmax *= ext[a];
    kmax -= 1ull;
    const int bits = kmax ? 64 - __clzll((long long)kmax) : 1;
    const int passes = (bits + 7) / 8;

    // cell keys (icp.py:120), lexicographic in (ix, iy[, iz])
    for (int i = tid; i < n; i += kBT) {
        unsigned long long key = 0ull;
        for (int a = 0; a < DIM; ++a) {
            const double cc = floor((raw[(size_t)i * DIM + a] - lo[a]) / voxel);
            key = key * ext[a] + (unsigned long long)cc;
        }
        keys[0][i] = key;
        idx[0][i] = (unsigned)i;
    }
    __syncthreads();

    // stable LSD radix sort: equal keys keep ascending input index
    int cur = 0;
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = 8 * pass;
        const unsigned long long* kin = keys[cur];
        const unsigned* iin = idx[cur];
        unsigned long long* kout = keys[cur ^ 1];
        unsigned* iout = idx[cur ^ 1];
        if (tid < 256) sh.hist[tid] = 0u;
        __syncthreads();
        for (int i = tid; i < n; i += kBT) atomicAdd(&sh.hist[(unsigned)(kin[i] >> shift) & 255u], 1u);
        __syncthreads();
        {
            unsigned tot;
            const unsigned ex = big_scan(tid < 256 ? sh.hist[tid] : 0u, sh, tot);
            if (tid < 256) sh.base[tid] = ex;
        }
        __syncthreads();
        for (int c0 = 0; c0 < n; c0 += kBT) {
            for (int k = tid; k < kBW * 256; k += kBT) (&sh.wcnt[0][0])[k] = 0;
            __syncthreads();
            const int i = c0 + tid;
            const bool ok = i < n;
            unsigned long long key = 0ull;
            unsigned id = 0u, d = 256u + (unsigned)lane;          // inactive lanes: unique pseudo-digits
            if (ok) { key = kin[i]; id = iin[i]; d = (unsigned)(key >> shift) & 255u; }
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const unsigned rank_in_warp = __popc(peers & ((1u << lane) - 1u));
            if (ok && rank_in_warp == 0) sh.wcnt[warp][d] = (unsigned short)__popc(peers);
            __syncthreads();
            if (ok) {
                unsigned before = 0;
                for (int w = 0; w < warp; ++w) before += sh.wcnt[w][d];
                const unsigned pos = sh.base[d] + before + rank_in_warp;
                kout[pos] = key;
                iout[pos] = id;
            }
            __syncthreads();
            if (tid < 256) {
                unsigned add = 0;
                for (int w = 0; w < kBW; ++w) add += sh.wcnt[w][tid];
                sh.base[tid] += add;
            }
            __syncthreads();
        }
        cur ^= 1;
    }
    const unsigned long long* ks = keys[cur];
    const unsigned* is = idx[cur];

    // run heads -> rows; means are input-order sums (icp.py:123-128)
    if (tid == 0) sh.carry = 0u;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += kBT) {
        const int i = c0 + tid;
        const bool head = i < n && (i == 0 || ks[i] != ks[i - 1]);
        unsigned tot;
        const unsigned ex = big_scan(head ? 1u : 0u, sh, tot);
        const unsigned carry = sh.carry;
        if (head) {
            const unsigned long long key = ks[i];
            double s[DIM];
            for (int a = 0; a < DIM; ++a) s[a] = 0.0;
            int cnt = 0;
            for (int m = i; m < n && ks[m] == key; ++m) {
                const size_t src = (size_t)is[m] * DIM;
                for (int a = 0; a < DIM; ++a) s[a] += raw[src + a];
                ++cnt;
            }
            const size_t pos = carry + ex;
            for (int a = 0; a < DIM; ++a) out[pos * DIM + a] = s[a] / (double)cnt;
        }
        __syncthreads();
        if (tid == 0) sh.carry = carry + tot;
        __syncthreads();
    }
    if (tid == 0) {
        cs.ds_n[c] = (int)sh.carry;
        for (int k = 0; k < 3; ++k) {
            cs.box[(size_t)c * 6 + k] = k < DIM ? lo[k] : 0.0;
            cs.box[(size_t)c * 6 + 3 + k] = k < DIM ? hi[k] : 0.0;
        }
    }
}

// ---- hash grid over one downsampled target (2-D), one CTA per cloud ----------------------
// start/items/cell scratch of cloud c begins at grid_off[c] (start has buckets+1 ints there).
__global__ void __launch_bounds__(kBT) big_grid_kernel(const CloudSet cs, double cell_size, const long long* grid_off,
                                                       const int* buckets_of, int* start_buf, int* items_buf,
                                                       int2* cell_buf, BigGrid* grids) {
    __shared__ BigShared sh;
    const int c = blockIdx.x;
    if (!cs.is_tgt[c]) return;
    const int m = cs.ds_n[c];
    if (m <= 0) return;
    const int tid = threadIdx.x;
    const long long beg = cs.off[c];
    const double* pts = cs.ds + beg * 2;
    const int buckets = buckets_of[c];
    int* start = start_buf + grid_off[c];
    int* items = items_buf + beg;
    int2* cell = cell_buf + beg;
    const double lox = cs.box[(size_t)c * 6], loy = cs.box[(size_t)c * 6 + 1];
    const double w = cs.box[(size_t)c * 6 + 3] - lox, hgt = cs.box[(size_t)c * 6 + 4] - loy;
    double h = fmax(cell_size, fmax(w, hgt) / 1.0e9);
    if (!(h > 0.0)) h = 1.0;
    const int gx = (int)fmin(w / h, 2.0e9) + 1, gy = (int)fmin(hgt / h, 2.0e9) + 1;
    const unsigned mask = (unsigned)buckets - 1u;
    for (int b = tid; b <= buckets; b += kBT) start[b] = 0;
    __syncthreads();
    for (int i = tid; i < m; i += kBT) {
        const int cx = min(gx - 1, max(0, (int)((pts[2 * i] - lox) / h)));
        const int cy = min(gy - 1, max(0, (int)((pts[2 * i + 1] - loy) / h)));
        atomicAdd(&start[big_cell_hash(cx, cy) & mask], 1);
    }
    __syncthreads();
    if (tid == 0) sh.carry = 0u;
    __syncthreads();
    for (int b0 = 0; b0 < buckets; b0 += kBT) {           // exclusive scan, in place
        const int b = b0 + tid;
        const unsigned v = b < buckets ? (unsigned)start[b] : 0u;
        unsigned tot;
        const unsigned ex = big_scan(v, sh, tot);
        const unsigned carry = sh.carry;
        if (b < buckets) start[b] = (int)(carry + ex);
        __syncthreads();
        if (tid == 0) sh.carry = carry + tot;
        __syncthreads();
    }
    for (int i = tid; i < m; i += kBT) {
        const int cx = min(gx - 1, max(0, (int)((pts[2 * i] - lox) / h)));
        const int cy = min(gy - 1, max(0, (int)((pts[2 * i + 1] - loy) / h)));
        const int pos = atomicAdd(&start[big_cell_hash(cx, cy) & mask], 1);
        items[pos] = i;
        cell[pos] = make_int2(cx, cy);
    }
    __syncthreads();
    if (tid == 0) {
        BigGrid g;
        g.start = start; g.items = items; g.cell = cell;
        g.h = h; g.lox = lox; g.loy = loy; g.nx = gx; g.ny = gy; g.mask = mask;
        grids[c] = g;
    }
}

// ---- normals on the global hash grid: one thread per target point -----------------------------
constexpr int kBigKnn = 64;
constexpr int kBigMaxRing = 24;

__global__ void __launch_bounds__(256) big_normals_kernel(const CloudSet cs, const BigGrid* grids, int normal_k,
                                                          int blocks_per_cloud) {
    const int c = blockIdx.x / blocks_per_cloud;
    if (!cs.is_tgt[c]) return;
    const int m = cs.ds_n[c];
    if (m <= 0) return;
    const BigGrid G = grids[c];
    const long long beg = cs.off[c];
    const double2* pts = reinterpret_cast<const double2*>(cs.ds + beg * 2);
    double* nrm = cs.nrm + beg * 2;
    const int K = min(normal_k, m - 1) + 1;
    for (int i = (blockIdx.x % blocks_per_cloud) * 256 + threadIdx.x; i < m; i += blocks_per_cloud * 256) {
        const double2 p = pts[i];
        const int cxi = min(G.nx - 1, max(0, (int)((p.x - G.lox) / G.h)));
        const int cyi = min(G.ny - 1, max(0, (int)((p.y - G.loy) / G.h)));
        double bd[kBigKnn];
        int bi[kBigKnn];
        int cnt = 0;
        auto offer = [&](double d, int j) {
            if (cnt == K && !(d < bd[K - 1] || (d == bd[K - 1] && j < bi[K - 1]))) return;
            int q = cnt < K ? cnt : K - 1;
            while (q > 0 && (bd[q - 1] > d || (bd[q - 1] == d && bi[q - 1] > j))) { bd[q] = bd[q - 1]; bi[q] = bi[q - 1]; --q; }
            bd[q] = d; bi[q] = j;
            if (cnt < K) ++cnt;
        };
        bool done = false;
        for (int r = 0; r <= kBigMaxRing; ++r) {
            const int x0 = cxi - r, x1 = cxi + r, y0 = cyi - r, y1 = cyi + r;
            for (int y = max(y0, 0); y <= min(y1, G.ny - 1); ++y) {
                const bool edge_row = (y == y0) || (y == y1);
                const int xstep = edge_row ? 1 : max(2 * r, 1);
                for (int x = x0; x <= x1; x += xstep) {
                    if (x < 0 || x >= G.nx) continue;
                    const unsigned b = big_cell_hash(x, y) & G.mask;
                    const int eb = b ? G.start[b - 1] : 0, ee = G.start[b];
                    for (int e = eb; e < ee; ++e) {
                        const int2 cc = G.cell[e];
                        if (cc.x != x || cc.y != y) continue;
                        const int j = G.items[e];
                        const double dx = p.x - pts[j].x, dy = p.y - pts[j].y;
                        offer(dx * dx + dy * dy, j);
                    }
                }
            }
            if ((x0 <= 0) && (y0 <= 0) && (x1 >= G.nx - 1) && (y1 >= G.ny - 1)) { done = true; break; }
            if (cnt == K) {
                double bound = INFINITY;
                if (x0 > 0) bound = fmin(bound, p.x - (G.lox + x0 * G.h));
                if (x1 < G.nx - 1) bound = fmin(bound, (G.lox + (x1 + 1) * G.h) - p.x);
                if (y0 > 0) bound = fmin(bound, p.y - (G.loy + y0 * G.h));
                if (y1 < G.ny - 1) bound = fmin(bound, (G.loy + (y1 + 1) * G.h) - p.y);
                bound = bound * (1.0 - 1e-9) - 1e-12 * G.h;
                if (bound > 0.0 && bd[K - 1] < bound * bound) { done = true; break; }
            }
        }
        if (!done) {
            cnt = 0;
            for (int j = 0; j < m; ++j) {
                const double dx = p.x - pts[j].x, dy = p.y - pts[j].y;
                offer(dx * dx + dy * dy, j);
            }
        }
        double mx = 0.0, my = 0.0;
        for (int q = 0; q < cnt; ++q) { mx += pts[bi[q]].x; my += pts[bi[q]].y; }
        mx /= (double)cnt; my /= (double)cnt;
        double sxx = 0.0, sxy = 0.0, syy = 0.0;
        for (int q = 0; q < cnt; ++q) {
            const double dx = pts[bi[q]].x - mx, dy = pts[bi[q]].y - my;
            sxx += dx * dx; sxy += dx * dy; syy += dy * dy;
        }
        double nv[2];
        sym2_min_eigvec(sxx, sxy, syy, nv);
        nrm[2 * i] = nv[0];
        nrm[2 * i + 1] = nv[1];
    }
}

// ---- launchers ---------------------------------------------------------------------------------
int launch_big_voxel(const CloudSet& cs, int dim, double voxel, unsigned long long* key_buf, unsigned* idx_buf,
                     long long total_points, cudaStream_t stream) {
    if (dim == 2) big_voxel_kernel<2><<<cs.n_clouds, kBT, 0, stream>>>(cs, voxel, key_buf, idx_buf, total_points);
    else          big_voxel_kernel<3><<<cs.n_clouds, kBT, 0, stream>>>(cs, voxel, key_buf, idx_buf, total_points);
    ICPB_LAUNCH_CHECK();
    return ICPB200_OK;
}

int launch_big_grid(const CloudSet& cs, double cell_size, const long long* d_grid_off, const int* d_buckets,
                    int* start_buf, int* items_buf, int2* cell_buf, BigGrid* d_grids, cudaStream_t stream) {
    big_grid_kernel<<<cs.n_clouds, kBT, 0, stream>>>(cs, cell_size, d_grid_off, d_buckets, start_buf, items_buf,
                                                      cell_buf, d_grids);
    ICPB_LAUNCH_CHECK();
    return ICPB200_OK;
}

int launch_big_normals(const CloudSet& cs, const BigGrid* d_grids, int normal_k, long long max_points,
                       cudaStream_t stream) {
    const int blocks_per_cloud = (int)std::max<long long>(1, std::min<long long>(64, (max_points + 255) / 256));
    big_normals_kernel<<<cs.n_clouds * blocks_per_cloud, 256, 0, stream>>>(cs, d_grids, normal_k, blocks_per_cloud);
    ICPB_LAUNCH_CHECK();
    return ICPB200_OK;
}

}  // namespace icpb
