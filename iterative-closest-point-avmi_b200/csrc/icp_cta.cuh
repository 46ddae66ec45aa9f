// Device building blocks shared by the per-pair ICP kernel and the standalone
// voxel-downsample kernel: deterministic block reductions, the CTA-local
// voxel-grid mean (utilities/icp.py:117-129) and its bitonic sort.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace icpb {

constexpr int kNT = 256;               // threads per CTA
constexpr int kNW = kNT / 32;          // warps per CTA
constexpr int kRedMax = 12;            // widest block reduction (values)

// Shared-memory header: lives at the start of dynamic shared memory in every
// kernel that uses these helpers.
struct CtaShared {
    double red[2][kNW][kRedMax];       // double-buffered warp partials
    int scan_tot[kNW + 1];
    int bcast_i[4];
    unsigned int pair;
    // registration state (ICP kernel only)
    double r_tot[9], t_tot[3], r[9], t[3];
    double center[3], lo_t[3], hi_t[3];
    double grid_h;
    int grid_nx, grid_ny;
    int n_s, n_t;
    int amb_n;
    // split-sweep partials (K3): [warp][lane]
    float part_b1[kNW][32], part_b2[kNW][32], part_b3[kNW][32];
    int part_bt[kNW][32], part_bt2[kNW][32];
};

struct SumOp { __device__ static double f(double a, double b) { return a + b; } };
struct MinOp { __device__ static double f(double a, double b) { return fmin(a, b); } };

// All-threads block reduction of NV doubles.  Fixed butterfly + fixed warp
// order => bitwise deterministic, and every thread returns the same value.
// `phase` alternates the scratch buffer so one barrier per reduction suffices.
template <int NV, class Op>
__device__ __forceinline__ void block_reduce(double (&v)[NV], CtaShared& sh, int& phase) {
    static_assert(NV <= kRedMax, "reduction too wide");
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] = Op::f(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    double(*buf)[kRedMax] = sh.red[phase & 1];
    if (l == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) buf[w][k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = buf[0][k];
#pragma unroll
        for (int ww = 1; ww < kNW; ++ww) s = Op::f(s, buf[ww][k]);
        v[k] = s;
    }
    ++phase;
}

// Exclusive prefix sum of one int per thread; also returns the block total.
__device__ __forceinline__ int block_excl_scan(int v, CtaShared& sh, int& total) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (l >= o) inc += t;
    }
    __syncthreads();                       // scan_tot may still be read by a previous call
    if (l == 31) sh.scan_tot[w] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int ww = 0; ww < kNW; ++ww) {
        const int c = sh.scan_tot[ww];
        if (ww < w) base += c;
        tot += c;
    }
    total = tot;
    return base + inc - v;
}

// Ascending bitonic sort of (key, idx) pairs, compared lexicographically, so
// equal keys keep ascending input index (a stable order).  n_pad is a power of
// two; padding entries carry key = ~0, idx = ~0.
__device__ __forceinline__ void bitonic_sort_pairs(unsigned long long* keys, unsigned int* idx, int n_pad) {
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pad; i += kNT) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long ka = keys[i], kb = keys[p];
                    const unsigned int ia = idx[i], ib = idx[p];
                    const bool a_gt_b = (ka > kb) || (ka == kb && ia > ib);
                    const bool asc = (i & k) == 0;
                    if (a_gt_b == asc) {
                        keys[i] = kb; keys[p] = ka;
                        idx[i] = ib; idx[p] = ia;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// Voxel-grid mean of one cloud, executed by the whole CTA.
//   raw : n rows of DIM float64 (global)        out : rows of DIM float64 (global)
//   keys/idx : shared scratch for n_pad entries (n_pad = power of two >= n, >= kNT)
// Restates utilities/icp.py:117-129: cell = floor((p - min) / voxel) per axis
// (fp64 subtract + divide, exactly as written), rows ordered lexicographically
// by cell index, each mean an input-order sum divided by the member count.
// Returns the number of occupied voxels, or -1 if the voxel index range does
// not fit 62 bits.  lo/hi receive the raw bounding box.
template <int DIM>
__device__ int cta_voxel_means(const double* __restrict__ raw, int n, double voxel,
                               double* __restrict__ out, unsigned long long* keys,
                               unsigned int* idx, int n_pad, CtaShared& sh, int& phase,
                               double* lo, double* hi) {
    // bounding box: min(p) and min(-p) in one reduction
    double b[2 * DIM];
#pragma unroll
    for (int a = 0; a < 2 * DIM; ++a) b[a] = INFINITY;
    for (int i = threadIdx.x; i < n; i += kNT) {
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            const double v = raw[(size_t)i * DIM + a];
            b[a] = fmin(b[a], v);
            b[DIM + a] = fmin(b[DIM + a], -v);
        }
    }
    block_reduce<2 * DIM, MinOp>(b, sh, phase);
    unsigned long long ext[DIM];
    double span = 1.0;
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
        lo[a] = b[a];
        hi[a] = -b[DIM + a];
        // floor((p - lo) / voxel) is monotone in p, so the largest index is at p = hi
        const double top = floor((hi[a] - lo[a]) / voxel);
        ext[a] = (unsigned long long)top + 1ull;
        span *= (top + 1.0);
    }
    if (!(span < 4.0e18)) return -1;

    for (int i = threadIdx.x; i < n_pad; i += kNT) {
        unsigned long long key = ~0ull;
        unsigned int id = ~0u;
        if (i < n) {
            key = 0ull;
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
                const double c = floor((raw[(size_t)i * DIM + a] - lo[a]) / voxel);
                key = key * ext[a] + (unsigned long long)c;
            }
            id = (unsigned int)i;
        }
        keys[i] = key;
        idx[i] = id;
    }
    __syncthreads();
    bitonic_sort_pairs(keys, idx, n_pad);

    // run heads -> output rows
    const int chunk = n_pad / kNT;
    const int beg = threadIdx.x * chunk;
    int heads = 0;
    for (int i = beg; i < beg + chunk; ++i)
        heads += (i < n) && (i == 0 || keys[i] != keys[i - 1]);
    int total;
    int pos = block_excl_scan(heads, sh, total);
    for (int i = beg; i < beg + chunk; ++i) {
        if ((i < n) && (i == 0 || keys[i] != keys[i - 1])) {
            const unsigned long long key = keys[i];
            double s[DIM];
#pragma unroll
            for (int a = 0; a < DIM; ++a) s[a] = 0.0;
            int cnt = 0;
            for (int m = i; m < n && keys[m] == key; ++m) {      // ascending input index
                const size_t src = (size_t)idx[m] * DIM;
#pragma unroll
                for (int a = 0; a < DIM; ++a) s[a] += raw[src + a];
                ++cnt;
            }
#pragma unroll
            for (int a = 0; a < DIM; ++a) out[(size_t)pos * DIM + a] = s[a] / (double)cnt;
            ++pos;
        }
    }
    __syncthreads();
    return total;
}

}  // namespace icpb
