// ICP registration kernels for sm_100a.
//
//   K1 voxel_clouds_kernel  one CTA per referenced cloud: voxel-grid mean
//                           (/root/reference/utilities/icp.py:150-151, 117-129)
//   K2 normals_kernel       one CTA per point-to-line target cloud: exact
//                           (k+1)-NN + 2x2 PCA normals (icp.py:165-167, 51-76)
//   K3 icp_pairs_kernel     persistent, one CTA per (source, target) pair from
//                           an atomic queue; the whole loop of icp.py:175-220
//                           (nearest neighbour -> gate -> solve -> apply ->
//                           error -> convergence test) runs out of shared
//                           memory without leaving the SM.
//
// Nearest neighbour in K3 (brute mode).  A warp takes 32 source points per
// "chunk" (one per lane, up to 8 chunks in registers) and sweeps the whole
// target, 32 points per tile, with the direct-difference fp32 distance
// (2 FADD + FMUL + FFMA per pair, the form SURVEY.md H2 found to reproduce
// the fp64 argmin).  The sweep only tracks per-tile minima; the winning tile
// is re-evaluated in fp64 and the runner-up tile minimum bounds every other
// target.  If that bound cannot exclude a closer point (margin = fp32
// rounding of the recentred coordinates) the point is re-decided by a full
// fp64 scan.  The correspondence is therefore the exact fp64 nearest
// neighbour (lowest index on exact ties), as scipy's KDTree query returns.
//
// Exact carry-over between iterations.  Each decision also yields D2, a lower
// bound on the distance from the point to every target other than its match.
// In later iterations a point that has moved by at most `moved` since then
// keeps its match without a sweep whenever dist(p, match) < D2 - moved
// (triangle inequality: every other target is still farther).  Points that
// fail the test are compacted and swept again.  ICP steps shrink
// geometrically, so most of the 150-iteration limit-cycle pairs' work
// disappears while every correspondence stays the exact nearest neighbour.
//
// All reductions are fp64, fixed-order (warp butterfly, then warps in order):
// no atomics in any sum, so results are bitwise reproducible run to run.
#include <stdlib.h>

#include <cooperative_groups.h>

#include "icp_b200.h"
#include "common.cuh"
#include "icp_cta.cuh"
#include "icp_kernel.h"
#include "linalg_small.cuh"

namespace icpb {

constexpr int kSMax = 8;               // chunks (of 32 source points) per warp per sweep round
constexpr unsigned kContClasses = 4;   // cost classes of the handed-over pairs (IcpArgs::cont_bucket)
constexpr int kKnnMax = 64;            // normal_k + 1 upper bound
constexpr int kGridCells = 4096;       // shared-memory uniform grid for the normals kNN
constexpr float kFar = 3.0e18f;        // padding target coordinate (distance^2 ~ 1.8e37, finite)

__host__ __device__ inline int pad_index(int j) { return j + (j >> 5); }   // 33-stride tiles
__host__ __device__ inline size_t align16(size_t b) { return (b + 15) & ~size_t(15); }

// ---- shared-memory sizes (must match the carve-up in the kernels) -----------
size_t icp_pair_smem_bytes(int dim, int cap_s, int cap_t, int nt) {
    size_t b = align16(nt == 512 ? sizeof(CtaSharedT<16>) : sizeof(CtaShared));
    b += sizeof(double) * dim * (size_t)(cap_t + cap_t / 32);          // tgt64, tile-padded SoA
    b = align16(b);
    b += sizeof(float) * (dim == 2 ? 2 : 4) * (size_t)cap_t;           // tgt32
    b += sizeof(double) * dim * (size_t)cap_s;                         // cur64 SoA
    b += sizeof(int) * (size_t)cap_s;                                  // match
    b += sizeof(float) * (1 + dim) * (size_t)cap_s;                    // d2lb, decision position (fp32, recentred)
    b += sizeof(unsigned short) * 3 * (size_t)cap_s;                   // todo, todo ordered by x, ambiguous
    b = align16(b);
    if (nt == 512 && dim == 2) b += sizeof(double) * 2 * (size_t)cap_t + 16;   // target normals, brought in by a TMA bulk copy
    return align16(b);
}
size_t icp_voxel_smem_bytes(int sort_pad) { return align16(sizeof(CtaShared)) + (size_t)sort_pad * 12; }
size_t icp_normals_smem_bytes(int cap_t) {
    const size_t grid_path = sizeof(int) * (kGridCells + 1) + 4 + sizeof(unsigned short) * 4 * (size_t)cap_t + 16;
    const size_t sweep_path = (sizeof(float2) + 2 * sizeof(float) + sizeof(unsigned short)) * (size_t)cap_t + 16;
    return align16(sizeof(CtaShared)) + align16(sizeof(double) * 2 * (size_t)(cap_t + cap_t / 32)) +
           (grid_path > sweep_path ? grid_path : sweep_path);
}

// ---- K0: which clouds are referenced / are point-to-line targets -------------
__global__ void mark_used_kernel(int n_pairs, const int* __restrict__ src_idx, const int* __restrict__ tgt_idx,
                                 unsigned char* used_s, unsigned char* used_t, unsigned char* is_tgt, int n_s, int n_t) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const int cs = src_idx ? src_idx[p] : p, ct = tgt_idx ? tgt_idx[p] : p;
    // (the device-resident entry point cannot look at its index arrays on the host: a pair that names a cloud outside the
    // set is skipped here and reported by the pair kernel with status ICPB200_BAD_VOXELS)
    if (cs < 0 || cs >= n_s || ct < 0 || ct >= n_t) return;
    used_s[cs] = 1;
    used_t[ct] = 1;
    if (is_tgt) is_tgt[ct] = 1;
}

// ---- K1: voxel-grid means, one CTA per cloud -----------------------------------
template <int DIM>
__global__ void __launch_bounds__(kNT) voxel_clouds_kernel(const CloudSet cs, double voxel, int sort_pad, int first) {
    extern __shared__ __align__(16) unsigned char smem[];
    CtaShared& sh = *reinterpret_cast<CtaShared*>(smem);
    const int c = first + blockIdx.x;
    if (cs.used && !cs.used[c]) return;
    const long long beg = cs.off[c];
    const long long n = cs.off[c + 1] - beg;
    if (n <= 0 || n > sort_pad) {
        if (threadIdx.x == 0) cs.ds_n[c] = -1;
        return;
    }
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem + align16(sizeof(CtaShared)));
    unsigned int* sidx = reinterpret_cast<unsigned int*>(smem + align16(sizeof(CtaShared)) + (size_t)sort_pad * 8);
    int phase = 0;
    double lo[DIM], hi[DIM];
    const int m = cta_voxel_means<DIM>(cs.raw + beg * DIM, (int)n, voxel, cs.ds + beg * DIM, keys, sidx, sort_pad,
                                       sh, phase, lo, hi);
    if (threadIdx.x == 0) {
        cs.ds_n[c] = m;
        for (int k = 0; k < 3; ++k) {
            cs.box[(size_t)c * 6 + k] = k < DIM ? lo[k] : 0.0;
            cs.box[(size_t)c * 6 + 3 + k] = k < DIM ? hi[k] : 0.0;
        }
    }
}

// Best-K list of (distance^2, index), ascending, ties by lower index.
// REG = true keeps kKnnReg slots in registers (bubble insertion, fully
// unrolled, no memory traffic) and serves K <= kKnnReg; REG = false is the
// general local-memory version for larger K.
constexpr int kKnnReg = 16;

template <bool REG>
struct BestK {
    static constexpr int N = REG ? kKnnReg : kKnnMax;
    double d[N];
    int j[N];
    int cnt;          // filled slots (local-memory version only)
    __device__ __forceinline__ void init() {
        if (REG) {
#pragma unroll
            for (int m = 0; m < N; ++m) { d[m] = INFINITY; j[m] = 0x7fffffff; }
        }
        cnt = 0;
    }
    __device__ __forceinline__ static bool less(double da, int ja, double db, int jb) {
        return da < db || (da == db && ja < jb);
    }
    __device__ __forceinline__ void offer(double dn, int jn, int K) {
        if (REG) {
            if (!less(dn, jn, d[N - 1], j[N - 1])) return;
#pragma unroll
            for (int m = 0; m < N; ++m) {                   // bubble the new entry down to its place
                const bool sw = less(dn, jn, d[m], j[m]);
                const double td = d[m]; const int tj = j[m];
                d[m] = sw ? dn : td;  j[m] = sw ? jn : tj;
                dn = sw ? td : dn;    jn = sw ? tj : jn;
            }
        } else {
            if (cnt == K && !less(dn, jn, d[K - 1], j[K - 1])) return;
            int m = cnt < K ? cnt : K - 1;
            while (m > 0 && less(dn, jn, d[m - 1], j[m - 1])) { d[m] = d[m - 1]; j[m] = j[m - 1]; --m; }
            d[m] = dn; j[m] = jn;
            if (cnt < K) ++cnt;
        }
    }
    // K-th smallest distance^2 so far (+inf while fewer than K candidates were seen)
    __device__ __forceinline__ double kth(int K) const {
        if (REG) {
            double v = INFINITY;
#pragma unroll
            for (int m = 0; m < N; ++m) v = (m == K - 1) ? d[m] : v;
            return v;
        }
        return cnt == K ? d[K - 1] : INFINITY;
    }
};

__device__ __forceinline__ unsigned cell_hash(int cx, int cy) {
    return (((unsigned)cx * 73856093u) ^ ((unsigned)cy * 19349663u)) & (kGridCells - 1);
}

constexpr int kKnnMaxRing = 12;        // beyond this ring the query falls back to a full scan

template <bool REG>
__device__ void knn_normals_pass(const double* tx, const double* ty, int n_t, int K, double* __restrict__ normals_out,
                                 const int* cell_start, const unsigned short* items, const ushort2* item_cell,
                                 const CtaShared& sh, const unsigned short* list, int n_list) {
    const double h = sh.grid_h, lox = sh.lo_t[0], loy = sh.lo_t[1];
    const int gx = sh.grid_nx, gy = sh.grid_ny;
    const int count = list ? n_list : n_t;
    for (int q = threadIdx.x; q < count; q += kNT) {
        const int i = list ? (int)list[q] : q;
        const double px = tx[pad_index(i)], py = ty[pad_index(i)];
        const int cxi = min(gx - 1, max(0, (int)((px - lox) / h)));
        const int cyi = min(gy - 1, max(0, (int)((py - loy) / h)));
        BestK<REG> best;
        best.init();
        bool done = false;
        for (int r = 0; r <= kKnnMaxRing; ++r) {
            const int x0 = cxi - r, x1 = cxi + r, y0 = cyi - r, y1 = cyi + r;
            for (int y = max(y0, 0); y <= min(y1, gy - 1); ++y) {
                const bool edge_row = (y == y0) || (y == y1);
                const int xstep = edge_row ? 1 : max(2 * r, 1);
                for (int x = x0; x <= x1; x += xstep) {
                    if (x < 0 || x >= gx) continue;
                    const unsigned b = cell_hash(x, y);
                    const int beg = b ? cell_start[b - 1] : 0, end = cell_start[b];
                    for (int e = beg; e < end; ++e) {
                        const ushort2 cc = item_cell[e];
                        if (cc.x != x || cc.y != y) continue;          // another cell hashed to this bucket
                        const int j = items[e];
                        const double dx = px - tx[pad_index(j)], dy = py - ty[pad_index(j)];
                        best.offer(dx * dx + dy * dy, j, K);
                    }
                }
            }
            // everything not yet visited lies outside the (2r+1)^2 block of cells
            if ((x0 <= 0) && (y0 <= 0) && (x1 >= gx - 1) && (y1 >= gy - 1)) { done = true; break; }
            const double kth = best.kth(K);
            if (kth < INFINITY) {
                double bound = INFINITY;
                if (x0 > 0) bound = fmin(bound, px - (lox + x0 * h));
                if (x1 < gx - 1) bound = fmin(bound, (lox + (x1 + 1) * h) - px);
                if (y0 > 0) bound = fmin(bound, py - (loy + y0 * h));
                if (y1 < gy - 1) bound = fmin(bound, (loy + (y1 + 1) * h) - py);
                bound = bound * (1.0 - 1e-9) - 1e-12 * h;
                if (bound > 0.0 && kth < bound * bound) { done = true; break; }
            }
        }
        if (!done) {                                   // isolated point: exact scan of the whole cloud
            best.init();
            for (int j = 0; j < n_t; ++j) {
                const double dx = px - tx[pad_index(j)], dy = py - ty[pad_index(j)];
                best.offer(dx * dx + dy * dy, j, K);
            }
        }
        // PCA of the K nearest (np.cov is two-pass: subtract the mean first)
        double mx = 0.0, my = 0.0;
#pragma unroll
        for (int m = 0; m < BestK<REG>::N; ++m)
            if (m < K) { mx += tx[pad_index(best.j[m])]; my += ty[pad_index(best.j[m])]; }
        mx /= (double)K; my /= (double)K;
        double sxx = 0.0, sxy = 0.0, syy = 0.0;
#pragma unroll
        for (int m = 0; m < BestK<REG>::N; ++m)
            if (m < K) {
                const double dx = tx[pad_index(best.j[m])] - mx, dy = ty[pad_index(best.j[m])] - my;
                sxx += dx * dx; sxy += dx * dy; syy += dy * dy;
            }
        double nrm[2];
        sym2_min_eigvec(sxx, sxy, syy, nrm);
        normals_out[2 * i] = nrm[0];
        normals_out[2 * i + 1] = nrm[1];
    }
}

// Fast selection for K <= 16 and clouds of at most 4096 points: candidates are
// ranked by a 32-bit key = fp32 distance^2 with its low 12 mantissa bits replaced by
// the point index, kept in a 16-slot min/max bubble network (2 integer ops per
// slot).  The keys only pre-select: the 16 survivors are re-evaluated in fp64 and
// the first K are accepted as the exact (k+1)-NN set when they are separated from
// everything else (the other survivors exactly, the rejected candidates through
// the key's truncation bound).  Points that fail the test are redone by the exact
// fp64 pass.  The normal itself is always computed from fp64 coordinates.
__device__ void knn_normals_fast(const double* tx, const double* ty, int n_t, int K, double* __restrict__ normals_out,
                                 const int* cell_start, const unsigned short* items, const ushort2* item_cell,
                                 CtaShared& sh, unsigned short* redo) {
    const double h = sh.grid_h, lox = sh.lo_t[0], loy = sh.lo_t[1];
    const int gx = sh.grid_nx, gy = sh.grid_ny;
    for (int i = threadIdx.x; i < n_t; i += kNT) {
        const double px = tx[pad_index(i)], py = ty[pad_index(i)];
        const int cxi = min(gx - 1, max(0, (int)((px - lox) / h)));
        const int cyi = min(gy - 1, max(0, (int)((py - loy) / h)));
        unsigned key[kKnnReg];
#pragma unroll
        for (int m = 0; m < kKnnReg; ++m) key[m] = 0xffffffffu;
        bool done = false;
        for (int r = 0; r <= kKnnMaxRing && !done; ++r) {
            const int x0 = cxi - r, x1 = cxi + r, y0 = cyi - r, y1 = cyi + r;
            for (int y = max(y0, 0); y <= min(y1, gy - 1); ++y) {
                const bool edge_row = (y == y0) || (y == y1);
                const int xstep = edge_row ? 1 : max(2 * r, 1);
                for (int x = x0; x <= x1; x += xstep) {
                    if (x < 0 || x >= gx) continue;
                    const unsigned b = cell_hash(x, y);
                    const int beg = b ? cell_start[b - 1] : 0, end = cell_start[b];
                    for (int e = beg; e < end; ++e) {
                        const ushort2 cc = item_cell[e];
                        if (cc.x != x || cc.y != y) continue;
                        const int j = items[e];
                        const float dx = (float)(px - tx[pad_index(j)]), dy = (float)(py - ty[pad_index(j)]);
                        unsigned kk = (__float_as_uint(fmaf(dy, dy, dx * dx)) & 0xfffff000u) | (unsigned)j;
                        if (kk >= key[kKnnReg - 1]) continue;
#pragma unroll
                        for (int m = 0; m < kKnnReg; ++m) {
                            const unsigned lo_ = min(key[m], kk);
                            kk = max(key[m], kk);
                            key[m] = lo_;
                        }
                    }
                }
            }
            if ((x0 <= 0) && (y0 <= 0) && (x1 >= gx - 1) && (y1 >= gy - 1)) { done = true; break; }
            unsigned kth = 0xffffffffu;
#pragma unroll
            for (int m = 0; m < kKnnReg; ++m) kth = (m == K - 1) ? key[m] : kth;
            if (kth != 0xffffffffu) {
                double bound = INFINITY;
                if (x0 > 0) bound = fmin(bound, px - (lox + x0 * h));
                if (x1 < gx - 1) bound = fmin(bound, (lox + (x1 + 1) * h) - px);
                if (y0 > 0) bound = fmin(bound, py - (loy + y0 * h));
                if (y1 < gy - 1) bound = fmin(bound, (loy + (y1 + 1) * h) - py);
                bound = bound * (1.0 - 1e-9) - 1e-12 * h;
                const double kth_ub = (double)__uint_as_float(kth | 0xfffu) * (1.0 + 1e-6);   // >= true distance^2
                if (bound > 0.0 && kth_ub < bound * bound) done = true;
            }
        }
        // exact re-evaluation of the survivors
        double e_in = -1.0, e_out = INFINITY;          // largest of the first K, smallest of the rest
        double mx = 0.0, my = 0.0;
#pragma unroll
        for (int m = 0; m < kKnnReg; ++m) {
            if (key[m] == 0xffffffffu) continue;
            const int j = (int)(key[m] & 0xfffu);
            const double qx = tx[pad_index(j)], qy = ty[pad_index(j)];
            const double d = (px - qx) * (px - qx) + (py - qy) * (py - qy);
            if (m < K) { e_in = fmax(e_in, d); mx += qx; my += qy; }
            else e_out = fmin(e_out, d);
        }
        const double lb_rejected = key[kKnnReg - 1] == 0xffffffffu
                                       ? INFINITY
                                       : (double)__uint_as_float(key[kKnnReg - 1] & 0xfffff000u) * (1.0 - 1e-6);
        unsigned kthk = 0xffffffffu;
#pragma unroll
        for (int m = 0; m < kKnnReg; ++m) kthk = (m == K - 1) ? key[m] : kthk;
        if (!done || kthk == 0xffffffffu || !(e_in < e_out) || !(e_in < lb_rejected)) {
            redo[atomicAdd(&sh.bcast_i[1], 1)] = (unsigned short)i;     // exact pass decides
            continue;
        }
        mx /= (double)K; my /= (double)K;
        double sxx = 0.0, sxy = 0.0, syy = 0.0;
#pragma unroll
        for (int m = 0; m < kKnnReg; ++m)
            if (m < K) {
                const int j = (int)(key[m] & 0xfffu);
                const double dx = tx[pad_index(j)] - mx, dy = ty[pad_index(j)] - my;
                sxx += dx * dx; sxy += dx * dy; syy += dy * dy;
            }
        double nrm[2];
        sym2_min_eigvec(sxx, sxy, syy, nrm);
        normals_out[2 * i] = nrm[0];
        normals_out[2 * i + 1] = nrm[1];
    }
}

// ---- K2 fast path: lock-step slab sweep ---------------------------------------------
// voxel_downsample emits its rows ordered by voxel column (icp.py:121-127), so a
// cloud is sorted by x up to one voxel.  A warp takes 32 consecutive queries (one
// per lane) and walks the cloud outwards from them, down and up in turns; every
// lane evaluates the SAME candidate at the same step, so the candidate is one
// broadcast shared-memory load and the lanes never diverge.  A direction stops
// once the prefix maximum (suffix minimum) of x over the unvisited part puts it
// farther from every lane's query than that lane's current K-th distance; the
// rule needs no sortedness to be correct, only to be fast.
//
// Candidates are ranked by a 32-bit key: fp32 distance^2 on box-centred
// coordinates with the low mantissa bits replaced by the point index, kept in a
// 16-slot min/max network that runs only when some lane improves its list.  The
// keys only pre-select: the survivors are re-evaluated in fp64, membership of the
// K nearest is repaired by exact swaps, and the set is accepted only when it is
// provably separated from everything the keys rejected or the sweep skipped
// (truncation, fp32 rounding and recentring slack included).  Anything else goes
// to the exact fp64 pass below.  The normal always comes from fp64 coordinates.
// Sorting networks on registers (every index is a compile-time constant after unrolling).
template <int N>
__device__ __forceinline__ void bitonic_sort_regs(unsigned (&v)[N]) {
#pragma unroll
    for (int k = 2; k <= N; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1)
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned lo_ = min(v[i], v[p]), hi_ = max(v[i], v[p]);
                    const bool asc = (i & k) == 0;
                    v[i] = asc ? lo_ : hi_;
                    v[p] = asc ? hi_ : lo_;
                }
            }
}
// v is bitonic on entry, ascending on exit
template <int N>
__device__ __forceinline__ void bitonic_merge_regs(unsigned (&v)[N]) {
#pragma unroll
    for (int j = N >> 1; j > 0; j >>= 1)
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int p = i ^ j;
            if (p > i) {
                const unsigned lo_ = min(v[i], v[p]), hi_ = max(v[i], v[p]);
                v[i] = lo_;
                v[p] = hi_;
            }
        }
}

__device__ __forceinline__ void pca_normal_masked(const double* tx, const double* ty, const unsigned (&key)[kKnnReg],
                                                  unsigned idx_mask, unsigned in_mask, int K, double* out2) {
    double mx = 0.0, my = 0.0;
#pragma unroll
    for (int m = 0; m < kKnnReg; ++m)
        if ((in_mask >> m) & 1u) {
            const int j = (int)(key[m] & idx_mask);
            mx += tx[pad_index(j)]; my += ty[pad_index(j)];
        }
    mx /= (double)K; my /= (double)K;
    double sxx = 0.0, sxy = 0.0, syy = 0.0;
#pragma unroll
    for (int m = 0; m < kKnnReg; ++m)
        if ((in_mask >> m) & 1u) {
            const int j = (int)(key[m] & idx_mask);
            const double dx = tx[pad_index(j)] - mx, dy = ty[pad_index(j)] - my;
            sxx += dx * dx; sxy += dx * dy; syy += dy * dy;
        }
    sym2_min_eigvec(sxx, sxy, syy, out2);
}

template <int KT>
__device__ void cta_normals_sweep(const double* tx, const double* ty, int n, int K_rt, double* __restrict__ normals_out,
                                  float2* pf, float* pmax, float* smin, unsigned short* redo, CtaShared& sh,
                                  int part, int split) {
    __shared__ float wtot[2][kNW];
    __shared__ int next_block;                              // query blocks are handed out dynamically
    const int K = KT > 0 ? KT : K_rt;                       // compile-time for the usual neighbourhood sizes
    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double cx = 0.5 * (sh.lo_t[0] + sh.hi_t[0]), cy = 0.5 * (sh.lo_t[1] + sh.hi_t[1]);
    const double half = 0.5 * fmax(sh.hi_t[0] - sh.lo_t[0], sh.hi_t[1] - sh.lo_t[1]);
    for (int j = tid; j < n; j += kNT)
        pf[j] = make_float2((float)(tx[pad_index(j)] - cx), (float)(ty[pad_index(j)] - cy));
    if (tid == 0) { sh.bcast_i[1] = 0; next_block = kNW; }
    __syncthreads();
    {   // prefix maximum / suffix minimum of the fp32 x coordinate
        const int per = (n + kNT - 1) / kNT;
        const int beg = tid * per, end = min(beg + per, n);
        float mx = -INFINITY, mn = INFINITY;               // mx over [beg, end), mn over the mirrored range
        for (int j = beg; j < end; ++j) {
            mx = fmaxf(mx, pf[j].x);
            mn = fminf(mn, pf[n - 1 - j].x);
        }
        float imx = mx, imn = mn;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float a = __shfl_up_sync(full, imx, o), b = __shfl_up_sync(full, imn, o);
            if (lane >= o) { imx = fmaxf(imx, a); imn = fminf(imn, b); }
        }
        if (lane == 31) { wtot[0][warp] = imx; wtot[1][warp] = imn; }
        float emx = __shfl_up_sync(full, imx, 1), emn = __shfl_up_sync(full, imn, 1);
        if (lane == 0) { emx = -INFINITY; emn = INFINITY; }
        __syncthreads();
        for (int w = 0; w < warp; ++w) { emx = fmaxf(emx, wtot[0][w]); emn = fminf(emn, wtot[1][w]); }
        for (int j = beg; j < end; ++j) {
            emx = fmaxf(emx, pf[j].x);          pmax[j] = emx;
            emn = fminf(emn, pf[n - 1 - j].x);  smin[n - 1 - j] = emn;
        }
    }
    __syncthreads();

    const unsigned idx_mask = n <= 1024 ? 0x3ffu : 0xfffu;
    const double eps_abs = 2.5e-7 * half;                   // |fp32-world distance - true distance|
    // (a cloud may be shared by `split` CTAs when the launch is small: this one takes blocks part, part + split, ...)
    for (int m = warp; (part + m * split) * 32 < n; m = __shfl_sync(full, lane == 0 ? atomicAdd(&next_block, 1) : 0, 0)) {
        const int base = (part + m * split) * 32;
        const int i = min(base + lane, n - 1);
        const bool valid = base + lane < n;
        const float2 q = pf[i];
        unsigned key[kKnnReg];
#pragma unroll
        for (int m = 0; m < kKnnReg; ++m) key[m] = 0xffffffffu;
        // Candidates are taken sixteen at a time (eight below, eight above), sorted in registers and merged into the
        // list: min(list[i], batch[15 - i]) keeps the sixteen smallest of the union as a bitonic sequence, one merge
        // network sorts it -- 240 min/max per batch instead of 32 per candidate for an insertion network.
        auto make_key = [&](int j) {
            const float2 t = pf[j];
            const float dx = q.x - t.x, dy = q.y - t.y;
            return (__float_as_uint(fmaf(dy, dy, dx * dx)) & ~idx_mask) | (unsigned)j;
        };
        auto kth_ub = [&]() {                               // >= fp32 distance^2 of the K-th candidate (NaN if none)
            unsigned kth = 0xffffffffu;
#pragma unroll
            for (int m = 0; m < kKnnReg; ++m) kth = (m == K - 1) ? key[m] : kth;
            return __uint_as_float(kth | idx_mask) * 1.02f;
        };
        int dn = min(base + 31, n - 1), up = base + 32;
        bool ddone = false, udone = up >= n;
        float gdn = INFINITY, gup = INFINITY;               // |dx| lower bound to the unvisited part, per lane
        while (!ddone || !udone) {
            unsigned batch[kKnnReg];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int jd = dn - u, ju = up + u;
                batch[u] = (!ddone && jd >= 0) ? make_key(jd) : 0xffffffffu;
                batch[8 + u] = (!udone && ju < n) ? make_key(ju) : 0xffffffffu;
            }
            unsigned bmin = batch[0];
#pragma unroll
            for (int u = 1; u < kKnnReg; ++u) bmin = min(bmin, batch[u]);
            if (__any_sync(full, bmin < key[kKnnReg - 1])) {
                bitonic_sort_regs<kKnnReg>(batch);
#pragma unroll
                for (int m = 0; m < kKnnReg; ++m) key[m] = min(key[m], batch[kKnnReg - 1 - m]);
                bitonic_merge_regs<kKnnReg>(key);
            }
            if (!ddone) {
                dn -= 8;
                if (dn < 0) ddone = true;
                else {
                    const float g = q.x - pmax[dn];
                    if (__all_sync(full, g > 0.f && g * g > kth_ub())) { ddone = true; gdn = g; }
                }
            }
            if (!udone) {
                up += 8;
                if (up >= n) udone = true;
                else {
                    const float g = smin[up] - q.x;
                    if (__all_sync(full, g > 0.f && g * g > kth_ub())) { udone = true; gup = g; }
                }
            }
        }
        // exact membership of the K nearest among the survivors (lexicographic (d, j) order)
        const double px = tx[pad_index(i)], py = ty[pad_index(i)];
        unsigned in_mask = (1u << K) - 1u;
        double e_in = 0.0;
        bool ok;
        {
            unsigned kthk = 0xffffffffu;
#pragma unroll
            for (int m = 0; m < kKnnReg; ++m) kthk = (m == K - 1) ? key[m] : kthk;
            ok = kthk != 0xffffffffu;
        }
        for (int pass = 0; pass < 6; ++pass) {
            double din = -1.0, dout = INFINITY;
            int jin = -1, jout = 0x7fffffff, min_ = 0, mout = 0;
#pragma unroll
            for (int m = 0; m < kKnnReg; ++m) {
                if (key[m] == 0xffffffffu) continue;
                const int j = (int)(key[m] & idx_mask);
                const double ex = px - tx[pad_index(j)], ey = py - ty[pad_index(j)];
                const double d = ex * ex + ey * ey;
                if ((in_mask >> m) & 1u) {
                    if (d > din || (d == din && j > jin)) { din = d; jin = j; min_ = m; }
                } else {
                    if (d < dout || (d == dout && j < jout)) { dout = d; jout = j; mout = m; }
                }
            }
            e_in = din;
            const bool sorted = din < dout || (din == dout && jin < jout);
            if (!sorted && ok) in_mask ^= (1u << min_) | (1u << mout);
            if (pass == 5 && !sorted) ok = false;
            if (!__any_sync(full, !sorted && ok)) break;
        }
        // everything outside the survivor list: rejected by key (>= truncated 16th key) or skipped by the sweep
        float lbf = fminf(gdn, gup);
        if (key[kKnnReg - 1] != 0xffffffffu)
            lbf = fminf(lbf, sqrtf(__uint_as_float(key[kKnnReg - 1] & ~idx_mask)) * (1.0f - 1e-6f));
        const double lb = (double)lbf * (1.0 - 1e-6) - eps_abs;
        ok = ok && lb > 0.0 && e_in < lb * lb;
        if (valid) {
            if (ok) {
                double nrm[2];
                pca_normal_masked(tx, ty, key, idx_mask, in_mask, K, nrm);
                normals_out[2 * i] = nrm[0];
                normals_out[2 * i + 1] = nrm[1];
            } else {
                redo[atomicAdd(&sh.bcast_i[1], 1)] = (unsigned short)i;
            }
        }
    }
    __syncthreads();
    // exact fp64 pass for the undecided points: same outward walk, one thread per point
    const int n_redo = sh.bcast_i[1];
    const double slack = 4.0e-7 * half;                     // fp32 recentring of both x coordinates, and then some
    for (int r = tid; r < n_redo; r += kNT) {
        const int i = redo[r];
        const double px = tx[pad_index(i)], py = ty[pad_index(i)];
        const double qx = (double)pf[i].x;
        BestK<true> best;
        best.init();
        for (int j = i; j >= 0; --j) {
            const double dx = px - tx[pad_index(j)], dy = py - ty[pad_index(j)];
            best.offer(dx * dx + dy * dy, j, K);
            if (j > 0) {
                const double g = qx - (double)pmax[j - 1] - slack;
                if (g > 0.0 && g * g > best.kth(K)) break;
            }
        }
        for (int j = i + 1; j < n; ++j) {
            const double g = (double)smin[j] - qx - slack;
            if (g > 0.0 && g * g > best.kth(K)) break;
            const double dx = px - tx[pad_index(j)], dy = py - ty[pad_index(j)];
            best.offer(dx * dx + dy * dy, j, K);
        }
        double mx = 0.0, my = 0.0;
#pragma unroll
        for (int m = 0; m < kKnnReg; ++m)
            if (m < K) { mx += tx[pad_index(best.j[m])]; my += ty[pad_index(best.j[m])]; }
        mx /= (double)K; my /= (double)K;
        double sxx = 0.0, sxy = 0.0, syy = 0.0;
#pragma unroll
        for (int m = 0; m < kKnnReg; ++m)
            if (m < K) {
                const double dx = tx[pad_index(best.j[m])] - mx, dy = ty[pad_index(best.j[m])] - my;
                sxx += dx * dx; sxy += dx * dy; syy += dy * dy;
            }
        double nrm[2];
        sym2_min_eigvec(sxx, sxy, syy, nrm);
        normals_out[2 * i] = nrm[0];
        normals_out[2 * i + 1] = nrm[1];
    }
}

// ---- K2: normals: exact (k+1)-NN on a shared-memory uniform grid + 2x2 PCA -------
// Restates utilities/icp.py:51-76.  All distances fp64; the neighbour set is
// the exact (k+1)-NN (ties broken by lower index).  np.cov's 1/(K-1) scale is
// dropped: it does not change the eigenvector.
__device__ void cta_normals_2d(const double* tx, const double* ty, int n_t, int normal_k, double voxel,
                               double* __restrict__ normals_out, int* cell_start,
                               unsigned short* items, ushort2* item_cell, unsigned short* redo, CtaShared& sh) {
    const int K = min(normal_k, n_t - 1) + 1;
    if (threadIdx.x == 0) {
        // cell edge: K voxel spacings.  Measured on 1080-beam scans (13 neighbours, 4 cm voxels): the
        // 13th neighbour is ~0.57 m away (median); 0.5 m cells visit ~18 cells and ~27 candidates per
        // point, 0.17 m cells ~94 cells for ~17 candidates -- the empty-cell visits cost more than the
        // candidates they save.  Cell coordinates must fit 16 bits.
        const double w = sh.hi_t[0] - sh.lo_t[0], hgt = sh.hi_t[1] - sh.lo_t[1];
        double h = fmax((double)K * voxel, fmax(w, hgt) / 30000.0);
        if (!(h > 0.0)) h = 1.0;
        sh.grid_h = h;
        sh.grid_nx = (int)(w / h) + 1;
        sh.grid_ny = (int)(hgt / h) + 1;
    }
    for (int c = threadIdx.x; c <= kGridCells; c += kNT) cell_start[c] = 0;
    __syncthreads();
    const double h = sh.grid_h, lox = sh.lo_t[0], loy = sh.lo_t[1];
    const int gx = sh.grid_nx, gy = sh.grid_ny;
    auto cell_of = [&](double x, double y, int& cxi, int& cyi) {
        cxi = min(gx - 1, max(0, (int)((x - lox) / h)));
        cyi = min(gy - 1, max(0, (int)((y - loy) / h)));
    };
    for (int i = threadIdx.x; i < n_t; i += kNT) {
        int cxi, cyi;
        cell_of(tx[pad_index(i)], ty[pad_index(i)], cxi, cyi);
        atomicAdd(&cell_start[cell_hash(cxi, cyi)], 1);
    }
    __syncthreads();
    {   // in-place exclusive scan over the buckets (kGridCells / kNT per thread)
        constexpr int per = kGridCells / kNT;
        const int beg = threadIdx.x * per;
        int local = 0;
        for (int c = beg; c < beg + per; ++c) local += cell_start[c];
        int total;
        int run = block_excl_scan(local, sh, total);
        for (int c = beg; c < beg + per; ++c) {
            const int v = cell_start[c];
            cell_start[c] = run;
            run += v;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_t; i += kNT) {
        int cxi, cyi;
        cell_of(tx[pad_index(i)], ty[pad_index(i)], cxi, cyi);
        const int pos = atomicAdd(&cell_start[cell_hash(cxi, cyi)], 1);
        items[pos] = (unsigned short)i;
        item_cell[pos] = make_ushort2((unsigned short)cxi, (unsigned short)cyi);
    }
    __syncthreads();
    // now bucket b holds items[(b ? cell_start[b-1] : 0) .. cell_start[b])
    if (K <= kKnnReg && n_t <= 4096) {
        if (threadIdx.x == 0) sh.bcast_i[1] = 0;
        __syncthreads();
        knn_normals_fast(tx, ty, n_t, K, normals_out, cell_start, items, item_cell, sh, redo);
        __syncthreads();
        const int n_redo = sh.bcast_i[1];
        if (n_redo > 0) knn_normals_pass<true>(tx, ty, n_t, K, normals_out, cell_start, items, item_cell, sh, redo, n_redo);
    } else if (K <= kKnnReg) {
        knn_normals_pass<true>(tx, ty, n_t, K, normals_out, cell_start, items, item_cell, sh, nullptr, 0);
    } else {
        knn_normals_pass<false>(tx, ty, n_t, K, normals_out, cell_start, items, item_cell, sh, nullptr, 0);
    }
}

__global__ void __launch_bounds__(kNT) normals_kernel(const CloudSet cs, int normal_k, int cap_t, double voxel, int first) {
    extern __shared__ __align__(16) unsigned char smem[];
    CtaShared& sh = *reinterpret_cast<CtaShared*>(smem);
    const int c = first + blockIdx.x;
    if (!cs.is_tgt[c]) return;
    const int n = cs.ds_n[c];
    if (n <= 0) return;
    const int tstride = cap_t + cap_t / 32;
    double* tx = reinterpret_cast<double*>(smem + align16(sizeof(CtaShared)));
    double* ty = tx + tstride;
    unsigned char* rest = smem + align16(sizeof(CtaShared)) + align16(sizeof(double) * 2 * (size_t)tstride);
    int* cell_start = reinterpret_cast<int*>(rest);
    ushort2* item_cell = reinterpret_cast<ushort2*>(rest + sizeof(int) * (kGridCells + 1));
    unsigned short* items = reinterpret_cast<unsigned short*>(item_cell + cap_t);
    unsigned short* redo = items + cap_t;
    const long long beg = cs.off[c];
    const double* ds = cs.ds + beg * 2;
    for (int j = threadIdx.x; j < n; j += kNT) {
        tx[pad_index(j)] = ds[2 * j];
        ty[pad_index(j)] = ds[2 * j + 1];
    }
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; ++k) { sh.lo_t[k] = cs.box[(size_t)c * 6 + k]; sh.hi_t[k] = cs.box[(size_t)c * 6 + 3 + k]; }
    }
    __syncthreads();
    cta_normals_2d(tx, ty, n, normal_k, voxel, cs.nrm + beg * 2, cell_start, items, item_cell, redo, sh);
}

// K2 fast kernel: normal_k + 1 <= 16 and every cloud <= 4096 points (host-checked).
__global__ void __launch_bounds__(kNT, 4) normals_sweep_kernel(const CloudSet cs, int normal_k, int cap_t, int first, int split) {
    extern __shared__ __align__(16) unsigned char smem[];
    CtaShared& sh = *reinterpret_cast<CtaShared*>(smem);
    const int c = first + blockIdx.x / split, part = blockIdx.x % split;
    if (!cs.is_tgt[c]) return;
    const int n = cs.ds_n[c];
    if (n <= 0) return;
    const int tstride = cap_t + cap_t / 32;
    double* tx = reinterpret_cast<double*>(smem + align16(sizeof(CtaShared)));
    double* ty = tx + tstride;
    unsigned char* rest = smem + align16(sizeof(CtaShared)) + align16(sizeof(double) * 2 * (size_t)tstride);
    float2* pf = reinterpret_cast<float2*>(rest);
    float* pmax = reinterpret_cast<float*>(pf + cap_t);
    float* smin = pmax + cap_t;
    unsigned short* redo = reinterpret_cast<unsigned short*>(smin + cap_t);
    const long long beg = cs.off[c];
    const double* ds = cs.ds + beg * 2;
    for (int j = threadIdx.x; j < n; j += kNT) {
        tx[pad_index(j)] = ds[2 * j];
        ty[pad_index(j)] = ds[2 * j + 1];
    }
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; ++k) { sh.lo_t[k] = cs.box[(size_t)c * 6 + k]; sh.hi_t[k] = cs.box[(size_t)c * 6 + 3 + k]; }
    }
    __syncthreads();
    const int K = min(normal_k, n - 1) + 1;
    if (K == 13) cta_normals_sweep<13>(tx, ty, n, K, cs.nrm + beg * 2, pf, pmax, smin, redo, sh, part, split);          // normal_k = 12 (config.yaml)
    else if (K == 11) cta_normals_sweep<11>(tx, ty, n, K, cs.nrm + beg * 2, pf, pmax, smin, redo, sh, part, split);     // normal_k = 10 (ICP default)
    else cta_normals_sweep<0>(tx, ty, n, K, cs.nrm + beg * 2, pf, pmax, smin, redo, sh, part, split);
}

// ---- K3: fp32 sweep --------------------------------------------------------------
// Tracks, per source point, the smallest per-tile minimum (b1, tile bt) and
// the smallest minimum over all OTHER tiles (b2).
template <int S>
__device__ __forceinline__ void sweep2d(const float4* __restrict__ t32, int tile_begin, int tile_end,
                                        const float (&sx)[S], const float (&sy)[S],
                                        float (&b1)[S], float (&b2)[S], int (&bt)[S]) {
#pragma unroll
    for (int s = 0; s < S; ++s) { b1[s] = INFINITY; b2[s] = INFINITY; bt[s] = tile_begin; }
    for (int tile = tile_begin; tile < tile_end; ++tile) {
        float tm[S];
#pragma unroll
        for (int s = 0; s < S; ++s) tm[s] = INFINITY;
        const float4* p = t32 + tile * 16;              // two targets per float4
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float4 t = p[q];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                float dx = sx[s] - t.x, dy = sy[s] - t.y;
                float d = fmaf(dy, dy, dx * dx);
                tm[s] = fminf(tm[s], d);
                dx = sx[s] - t.z; dy = sy[s] - t.w;
                d = fmaf(dy, dy, dx * dx);
                tm[s] = fminf(tm[s], d);
            }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const bool lt = tm[s] < b1[s];
            b2[s] = lt ? b1[s] : fminf(b2[s], tm[s]);
            bt[s] = lt ? tile : bt[s];
            b1[s] = fminf(b1[s], tm[s]);
        }
    }
}

template <int S>
__device__ __forceinline__ void sweep3d(const float4* __restrict__ t32, int tile_begin, int tile_end,
                                        const float (&sx)[S], const float (&sy)[S],
                                        const float (&sz)[S], float (&b1)[S], float (&b2)[S],
                                        int (&bt)[S]) {
#pragma unroll
    for (int s = 0; s < S; ++s) { b1[s] = INFINITY; b2[s] = INFINITY; bt[s] = tile_begin; }
    for (int tile = tile_begin; tile < tile_end; ++tile) {
        float tm[S];
#pragma unroll
        for (int s = 0; s < S; ++s) tm[s] = INFINITY;
        const float4* p = t32 + tile * 32;              // one target per float4 (x, y, z, 0)
#pragma unroll 8
        for (int q = 0; q < 32; ++q) {
            const float4 t = p[q];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const float dx = sx[s] - t.x, dy = sy[s] - t.y, dz = sz[s] - t.z;
                const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                tm[s] = fminf(tm[s], d);
            }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const bool lt = tm[s] < b1[s];
            b2[s] = lt ? b1[s] : fminf(b2[s], tm[s]);
            bt[s] = lt ? tile : bt[s];
            b1[s] = fminf(b1[s], tm[s]);
        }
    }
}

// ---- TMA bulk copy (cp.async.bulk) global -> shared with an mbarrier: sm_90+/sm_100 data movement, one thread issues ------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(src_gmem), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(b), "r"(parity) : "memory");
}

// Relative slack of an fp32 swept distance against the exact one.  dx = sx - tx, dx * dx and the fma each round once:
// d^2 is off by at most 3 * 2^-24 = 1.8e-7 relative, the distance by 0.9e-7; the coordinates' own rounding is covered
// separately by the absolute term 1.8e-7 * (|s| + |t|).  3e-7 leaves a factor three (it was 1e-6: far-away points,
// whose candidates differ by less than that, went to the fp64 fallback for nothing).
constexpr double kRel32 = 3.0e-7;

template <int DIM>
struct Loop {
    // views into shared memory, valid during the iteration loop
    const double* tx; const double* ty; const double* tz;     // tile-padded target fp64
    double* cx; double* cy; double* cz;                        // current source fp64 (SoA)
    float4* t32;
    int* match;
    float* d2lb;             // lower bound on the distance to every non-match target at decision time
    float* p0;               // position at that decision (fp32, recentred; SoA with stride cap): the bound holds
                             // as long as the NET displacement since then leaves room, however long the path
    int cap;
    unsigned short* todo;    // points that need a sweep this iteration (in the order the classify pass found them)
    unsigned short* todo2;   // the same points ordered by their current x (bins of the target's x range), for the slab sweep
    const unsigned short* list;   // what the sweeps read: todo or todo2
    unsigned short* amb;     // points whose fp32 ranking could not be separated: decided in fp64
    const double2* nrm_s;    // 512-thread variant: the target's normals in shared memory (nullptr: read them through L2)
    int* amb_n;              // number of entries of amb[]                  } this CTA's own, or -- in a helper CTA of a
    unsigned* slab_evals;    // candidates the slab sweeps looked at (statistics) } cluster -- the owner's, through DSMEM
    int n_s, n_t, n_tiles;
    double c0, c1, c2;       // recentring offset
    float ta;                // sum over axes of max |target - centre|
    // grid mode: the target stays in global memory (rows of DIM doubles) behind a hash grid
    const double* tg;
    BigGrid grid;
};

// target point j: shared memory (tile-padded SoA) in brute mode, global memory in grid mode
template <int DIM, bool GRID>
__device__ __forceinline__ void tgt_at(const Loop<DIM>& L, int j, double& x, double& y, double& z) {
    if (GRID) {
        if (DIM == 2) { const double2 v = __ldg(reinterpret_cast<const double2*>(L.tg) + j); x = v.x; y = v.y; z = 0.0; }
        else { x = __ldg(L.tg + 3 * (size_t)j); y = __ldg(L.tg + 3 * (size_t)j + 1); z = __ldg(L.tg + 3 * (size_t)j + 2); }
    } else {
        const int jp = pad_index(j);
        x = L.tx[jp]; y = L.ty[jp]; z = DIM == 3 ? L.tz[jp] : 0.0;
    }
}

// exact squared distance in fp64 between a point and target j
template <int DIM, bool GRID = false>
__device__ __forceinline__ double dist2_64(const Loop<DIM>& L, double px, double py, double pz, int j) {
    double qx, qy, qz;
    tgt_at<DIM, GRID>(L, j, qx, qy, qz);
    const double dx = px - qx, dy = py - qy;
    double d = dx * dx + dy * dy;
    if (DIM == 3) { const double dz = pz - qz; d += dz * dz; }
    return d;
}

__device__ __forceinline__ float f32_down(double v) { return __double2float_rd(v); }

// remember where point i stood when its correspondence was decided
template <int DIM>
__device__ __forceinline__ void stamp_p0(const Loop<DIM>& L, int i) {
    L.p0[i] = (float)(L.cx[i] - L.c0);
    L.p0[L.cap + i] = (float)(L.cy[i] - L.c1);
    if (DIM == 3) L.p0[2 * L.cap + i] = (float)(L.cz[i] - L.c2);
}

// Correspondence word.  Brute mode (targets <= 4096 points) packs two candidates: the match in
// the low half and, in the high half, alt + 1 -- the one other target that may overtake it
// (0 = none).  d2lb then bounds the distance to every target OTHER THAN THESE TWO, so a point
// that flips between two near-equidistant targets (the reference's limit cycles) is re-decided
// from two exact distances instead of a sweep.  Grid mode stores the plain index.
template <bool GRID> __device__ __forceinline__ int m_idx(int w) { return GRID ? w : (w & 0xffff); }
template <bool GRID> __device__ __forceinline__ int m_alt(int w) { return GRID ? -1 : ((w >> 16) - 1); }
__device__ __forceinline__ int m_pack(int j, int alt) { return j | ((alt + 1) << 16); }

// Decide the correspondence of source point i from its sweep result: re-evaluate
// the winning tile in fp64, bound everything else by the runner-up tile minimum.
template <int DIM, class SH>
__device__ __forceinline__ void decide(const Loop<DIM>& L, SH& sh, int i, float sxv, float syv, float szv,
                                       float b2v, int btv) {
    const double px = L.cx[i], py = L.cy[i], pz = DIM == 3 ? L.cz[i] : 0.0;
    const int j0 = btv * 32;
    const int j1 = min(j0 + 32, L.n_t);
    double best = INFINITY, second = INFINITY, third = INFINITY;
    int bj = j0, j2 = -1;
    for (int j = j0; j < j1; ++j) {
        const double d = dist2_64<DIM>(L, px, py, pz, j);
        if (d < best) { third = second; second = best; j2 = bj; best = d; bj = j; }
        else if (d < second) { third = second; second = d; j2 = j; }
        else if (d < third) third = d;
    }
    if (!(second < INFINITY)) j2 = -1;
    // Can a target outside the winning tile be closer?  fp32 coordinate
    // rounding moves a distance by at most eps32 * (|s| + |t|) summed over
    // axes; 3x safety on the 2^-24 unit roundoff, and the sweep's own four
    // roundings folded into the relative factor.
    const float mag = fabsf(sxv) + fabsf(syv) + (DIM == 3 ? fabsf(szv) : 0.f) + L.ta;
    const double slack = 1.8e-7 * (double)mag;
    const double other = sqrt((double)b2v) * (1.0 - kRel32) - slack;
    const double d1 = sqrt(best);
    if (!(other > d1)) {
        L.match[i] = m_pack(bj, -1);                   // the best known candidate: the fp64 fallback starts from it
        L.amb[atomicAdd(L.amb_n, 1)] = (unsigned short)i;
    } else {
        L.match[i] = m_pack(bj, j2);
        L.d2lb[i] = f32_down(fmin(other, sqrt(third) * (1.0 - 1e-12)));       // everything but bj and j2
        stamp_p0<DIM>(L, i);
    }
}

// One sweep round for this warp: chunks first_chunk + s * (NT / 32), s < S, of the todo list.
template <int DIM, int S, int NT, class SH>
__device__ __forceinline__ void nn_round(const Loop<DIM>& L, SH& sh, int first_chunk, int n_todo) {
    const int lane = threadIdx.x & 31;
    float sx[S], sy[S], sz[S];
    float b1[S], b2[S];
    int bt[S], pt[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int q = (first_chunk + s * (NT / 32)) * 32 + lane;
        pt[s] = q < n_todo ? (int)L.list[q] : -1;
        const int i = max(pt[s], 0);
        sx[s] = pt[s] >= 0 ? (float)(L.cx[i] - L.c0) : 0.f;
        sy[s] = pt[s] >= 0 ? (float)(L.cy[i] - L.c1) : 0.f;
        sz[s] = (DIM == 3 && pt[s] >= 0) ? (float)(L.cz[i] - L.c2) : 0.f;
    }
    if (DIM == 2) sweep2d<S>(L.t32, 0, L.n_tiles, sx, sy, b1, b2, bt);
    else          sweep3d<S>(L.t32, 0, L.n_tiles, sx, sy, sz, b1, b2, bt);
#pragma unroll
    for (int s = 0; s < S; ++s)
        if (pt[s] >= 0) decide<DIM>(L, sh, pt[s], sx[s], sy[s], sz[s], b2[s], bt[s]);
}

// Fewer chunks than warps (the nearly-converged regime): F warps share one chunk,
// each sweeping 1/F of the target.  With one point per lane the sweep can afford
// to track the argmin itself and the runner-up distance (3 more instructions per
// pair), so the decision needs a single fp64 evaluation instead of a 32-point
// tile scan: accept j1 when the fp32 runner-up, shrunk by the rounding slack, is
// still farther than the exact distance to j1; otherwise the full fp64 scan decides.
template <int DIM, int NT, class SH>
__device__ __forceinline__ void nn_split(const Loop<DIM>& L, SH& sh, int n_todo, int n_chunks, int F) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = w / F, part = w % F;
    const bool live = chunk < n_chunks;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    int pt = -1;
    if (live) {
        const int q = chunk * 32 + lane;
        pt = q < n_todo ? (int)L.list[q] : -1;
        const int i = max(pt, 0);
        sx = pt >= 0 ? (float)(L.cx[i] - L.c0) : 0.f;
        sy = pt >= 0 ? (float)(L.cy[i] - L.c1) : 0.f;
        sz = (DIM == 3 && pt >= 0) ? (float)(L.cz[i] - L.c2) : 0.f;
        const int j0 = (int)((long long)part * L.n_tiles / F) * 32, j1 = (int)((long long)(part + 1) * L.n_tiles / F) * 32;
        float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
        int bj = j0, bj2 = j0;
        auto offer = [&](float d, int j) {
            const bool lt1 = d < b1, lt2 = d < b2;
            b3 = lt2 ? b2 : fminf(b3, d);
            b2 = lt1 ? b1 : (lt2 ? d : b2);
            bj2 = lt1 ? bj : (lt2 ? j : bj2);
            b1 = lt1 ? d : b1;
            bj = lt1 ? j : bj;
        };
        if (DIM == 2) {
            const float2* t = reinterpret_cast<const float2*>(L.t32);
#pragma unroll 8
            for (int j = j0; j < j1; ++j) {
                const float2 q2 = t[j];
                const float dx = sx - q2.x, dy = sy - q2.y;
                offer(fmaf(dy, dy, dx * dx), j);
            }
        } else {
#pragma unroll 8
            for (int j = j0; j < j1; ++j) {
                const float4 q4 = L.t32[j];
                const float dx = sx - q4.x, dy = sy - q4.y, dz = sz - q4.z;
                offer(fmaf(dz, dz, fmaf(dy, dy, dx * dx)), j);
            }
        }
        sh.part_b1[w][lane] = b1; sh.part_b2[w][lane] = b2; sh.part_b3[w][lane] = b3;
        sh.part_bt[w][lane] = bj; sh.part_bt2[w][lane] = bj2;
    }
    __syncthreads();
    if (live && part == 0 && pt >= 0) {
        // merge the F partial top-3 lists (fp32 ranking), then decide between the two front
        // runners with exact distances; the third bounds every other target
        float m1 = INFINITY, m2 = INFINITY, m3 = INFINITY;
        int mj = 0, mj2 = 0;
        auto offer = [&](float d, int j) {
            const bool lt1 = d < m1, lt2 = d < m2;
            m3 = lt2 ? m2 : fminf(m3, d);
            m2 = lt1 ? m1 : (lt2 ? d : m2);
            mj2 = lt1 ? mj : (lt2 ? j : mj2);
            m1 = lt1 ? d : m1;
            mj = lt1 ? j : mj;
        };
        for (int f = 0; f < F; ++f) {
            offer(sh.part_b1[w + f][lane], sh.part_bt[w + f][lane]);
            offer(sh.part_b2[w + f][lane], sh.part_bt2[w + f][lane]);
            m3 = fminf(m3, sh.part_b3[w + f][lane]);
        }
        mj = min(mj, L.n_t - 1);                       // padding targets can only win if n_t == 0
        const bool two = mj2 < L.n_t && mj2 != mj && m2 < 1.0e30f;
        const double px = L.cx[pt], py = L.cy[pt], pz = DIM == 3 ? L.cz[pt] : 0.0;
        double e1 = dist2_64<DIM>(L, px, py, pz, mj);
        int j1 = mj, j2 = -1;
        if (two) {
            const double e2 = dist2_64<DIM>(L, px, py, pz, mj2);
            j2 = mj2;
            if (e2 < e1 || (e2 == e1 && mj2 < mj)) { j1 = mj2; j2 = mj; e1 = e2; }
        }
        const double d1 = sqrt(e1);
        const float mag = fabsf(sx) + fabsf(sy) + (DIM == 3 ? fabsf(sz) : 0.f) + L.ta;
        const double other = sqrt((double)(two ? m3 : m2)) * (1.0 - kRel32) - 1.8e-7 * (double)mag;
        if (!(other > d1)) {
            L.match[pt] = m_pack(j1, -1);
            L.amb[atomicAdd(L.amb_n, 1)] = (unsigned short)pt;
        } else {
            L.match[pt] = m_pack(j1, j2);
            L.d2lb[pt] = f32_down(other);
            stamp_p0<DIM>(L, pt);
        }
    }
}

// The same split for lists of 3 chunks up to half the CTA's warps (pairs that keep re-deciding a few hundred points, the
// helpers' stretches of a shared sweep): tracking three candidates costs 13 instructions per evaluation, 8 of them on the
// half-rate pipe, and made a list of 8 chunks as expensive as a bulk sweep of 20.  Here each warp sweeps its share of
// the target's TILES with the tile sweep's bookkeeping (one min per evaluation; the best tile and the best of the other
// tiles per point); the partial results meet in shared memory and the first warp of the chunk decides as the bulk sweep
// does -- the winning tile re-evaluated in fp64, everything else bounded by the runner-up tile.
template <int DIM, int NT, class SH>
__device__ __forceinline__ void nn_split_tiles(const Loop<DIM>& L, SH& sh, int n_todo, int n_chunks, int F) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = w / F, part = w % F;
    const bool live = chunk < n_chunks;
    float sx[1] = {0.f}, sy[1] = {0.f}, sz[1] = {0.f};
    int pt = -1;
    if (live) {
        const int q = chunk * 32 + lane;
        pt = q < n_todo ? (int)L.list[q] : -1;
        const int i = max(pt, 0);
        sx[0] = pt >= 0 ? (float)(L.cx[i] - L.c0) : 0.f;
        sy[0] = pt >= 0 ? (float)(L.cy[i] - L.c1) : 0.f;
        sz[0] = (DIM == 3 && pt >= 0) ? (float)(L.cz[i] - L.c2) : 0.f;
        const int t0 = (int)((long long)part * L.n_tiles / F), t1 = (int)((long long)(part + 1) * L.n_tiles / F);
        float b1[1], b2[1];
        int bt[1];
        if (DIM == 2) sweep2d<1>(L.t32, t0, t1, sx, sy, b1, b2, bt);
        else          sweep3d<1>(L.t32, t0, t1, sx, sy, sz, b1, b2, bt);
        sh.part_b1[w][lane] = b1[0]; sh.part_b2[w][lane] = b2[0]; sh.part_bt[w][lane] = bt[0];
    }
    __syncthreads();
    if (live && part == 0 && pt >= 0) {
        float m1 = INFINITY, m2 = INFINITY;
        int mt = 0;
        for (int f = 0; f < F; ++f) {
            const float v1 = sh.part_b1[w + f][lane], v2 = sh.part_b2[w + f][lane];
            const bool lt = v1 < m1;
            m2 = fminf(fminf(m2, v2), lt ? m1 : v1);           // the best of all tiles but the winner
            mt = lt ? sh.part_bt[w + f][lane] : mt;
            m1 = fminf(m1, v1);
        }
        decide<DIM>(L, sh, pt, sx[0], sy[0], sz[0], m2, mt);
    }
}

// Bulk sweeps (many points to decide): K1 emits every cloud ordered by voxel column, so both the
// target and the todo list (ascending source index) are sorted by x up to one voxel.  A warp takes
// 32 consecutive todo points and walks the target outwards from their position, down and up in
// turns; all lanes evaluate the SAME candidate (one broadcast load), tracking their three best in
// fp32.  A direction stops once the x-gap to the unvisited part (whose x is bounded through the
// voxel-column order) exceeds every lane's third-best distance.  The decision is nn_split's: the two
// front runners are compared in fp64, everything else -- visited or not -- is bounded by
// min(third best, gaps) minus the fp32 slack; otherwise the full fp64 scan decides.  On C2 a chunk
// visits ~60 of the 800 targets instead of all of them, and the 32-target fp64 re-evaluation of
// the tile sweep is gone.
// DS (lists of at most half as many chunks as the CTA has warps): two warps per chunk, one walks down, the other up; the
// up-walker leaves its three best and its gap in shared memory and the down-walker merges them in front of the decision.
template <int DIM, int NT, bool DS, class SH>
__device__ __forceinline__ void nn_slab(const Loop<DIM>& L, SH& sh, int n_todo, float vox) {
    const unsigned full = 0xffffffffu;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks = (n_todo + 31) >> 5;
    const float* tf = reinterpret_cast<const float*>(L.t32);
    constexpr int kStride = DIM == 2 ? 2 : 4;                 // floats per target in t32
    for (int chunk = DS ? (w >> 1) : w; chunk < (DS ? (NT / 64) : n_chunks); chunk += (NT / 32)) {
        const bool live = !DS || chunk < n_chunks;            // (DS: one round, every warp reaches the barrier below)
        const bool walk_dn = !DS || (w & 1) == 0, walk_up = !DS || (w & 1) == 1;
        if (DS && !live) { __syncthreads(); continue; }
        const int q = chunk * 32 + lane;
        const int pt = q < n_todo ? (int)L.list[q] : -1;
        const int i = pt >= 0 ? pt : (int)L.list[chunk * 32];   // idle lanes shadow the chunk's first point
        const float sx = (float)(L.cx[i] - L.c0), sy = (float)(L.cy[i] - L.c1);
        const float sz = DIM == 3 ? (float)(L.cz[i] - L.c2) : 0.f;
        float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
        int bj = 0, bj2 = -1;
        auto dist = [&](int j) {
            const float dx = sx - tf[j * kStride], dy = sy - tf[j * kStride + 1];
            float d = fmaf(dy, dy, dx * dx);
            if (DIM == 3) { const float dz = sz - tf[j * kStride + 2]; d = fmaf(dz, dz, d); }
            return d;
        };
        auto offer_d = [&](float d, int j) {
            const bool lt1 = d < b1, lt2 = d < b2;
            b3 = lt2 ? b2 : fminf(b3, d);
            b2 = lt1 ? b1 : (lt2 ? d : b2);
            bj2 = lt1 ? bj : (lt2 ? j : bj2);
            b1 = lt1 ? d : b1;
            bj = lt1 ? j : bj;
        };
        auto offer = [&](int j) { offer_d(dist(j), j); };
        // eight candidates at a time: their distances are independent (the three-best update is a dependent chain of
        // compares and selects), and once the lists have warmed up a block rarely holds anything closer than some
        // lane's third best -- then it costs eight evaluations and a vote instead of eight chained updates
        auto offer8 = [&](int j0, int step) {
            float d[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) d[u] = dist(j0 + u * step);
            const float m = fminf(fminf(fminf(d[0], d[1]), fminf(d[2], d[3])), fminf(fminf(d[4], d[5]), fminf(d[6], d[7])));
            if (__any_sync(full, m < b3)) {
#pragma unroll
                for (int u = 0; u < 8; ++u) offer_d(d[u], j0 + u * step);
            }
        };
        // start where lane 0's point would be inserted (any start is correct, this one is short)
        const float x0 = __shfl_sync(full, sx, 0);
        int lo = 0, hi = L.n_t;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (tf[mid * kStride] < x0) lo = mid + 1; else hi = mid;
        }
        int dn = lo - 1, up = lo;
        bool ddone = dn < 0 || !walk_dn, udone = up >= L.n_t || !walk_up;
        float gdn = INFINITY, gup = INFINITY;                 // x-gap to the unvisited part, per lane
        while (!ddone || !udone) {
            if (!ddone) {
                const int stop = max(dn - 7, 0);
                if (dn >= 7) {
                    offer8(dn, -1);
                } else {
                    for (int j = dn; j >= stop; --j) offer(j);
                }
                dn = stop - 1;
                if (dn < 0) ddone = true;
                else {
                    const float g = sx - (tf[dn * kStride] + vox);      // every j' <= dn has x < x[dn] + voxel
                    if (__all_sync(full, g > 0.f && g * g > b3 * 1.0001f)) { ddone = true; gdn = g; }
                }
            }
            if (!udone) {
                const int stop = min(up + 7, L.n_t - 1);
                if (up + 7 < L.n_t) {
                    offer8(up, 1);
                } else {
                    for (int j = up; j <= stop; ++j) offer(j);
                }
                up = stop + 1;
                if (up >= L.n_t) udone = true;
                else {
                    const float g = (tf[up * kStride] - vox) - sx;      // every j' >= up has x > x[up] - voxel
                    if (__all_sync(full, g > 0.f && g * g > b3 * 1.0001f)) { udone = true; gup = g; }
                }
            }
        }
        if (lane == 0) atomicAdd(L.slab_evals, (unsigned)((walk_dn ? lo - 1 - dn : 0) + (walk_up ? up - lo : 0)) * 32u);      // statistics only
        if (DS) {
            if (walk_up) {
                sh.part_b1[w][lane] = b1; sh.part_b2[w][lane] = b2; sh.part_b3[w][lane] = b3;
                sh.part_bt[w][lane] = bj; sh.part_bt2[w][lane] = bj2; sh.part_gap[w][lane] = gup;
            }
            __syncthreads();
            if (walk_up) continue;
            offer_d(sh.part_b1[w + 1][lane], sh.part_bt[w + 1][lane]);
            offer_d(sh.part_b2[w + 1][lane], sh.part_bt2[w + 1][lane]);
            b3 = fminf(b3, sh.part_b3[w + 1][lane]);
            gup = sh.part_gap[w + 1][lane];
        }
        if (pt < 0) continue;
        const bool two = bj2 >= 0 && bj2 != bj && b2 < 1.0e30f;
        const double px = L.cx[pt], py = L.cy[pt], pz = DIM == 3 ? L.cz[pt] : 0.0;
        double e1 = dist2_64<DIM>(L, px, py, pz, bj);
        int j1 = bj, j2 = -1;
        if (two) {
            const double e2 = dist2_64<DIM>(L, px, py, pz, bj2);
            j2 = bj2;
            if (e2 < e1 || (e2 == e1 && bj2 < bj)) { j1 = bj2; j2 = bj; e1 = e2; }
        }
        const double d1 = sqrt(e1);
        const float mag = fabsf(sx) + fabsf(sy) + (DIM == 3 ? fabsf(sz) : 0.f) + L.ta + vox;
        const double rest = fmin(sqrt((double)(two ? b3 : b2)) * (1.0 - kRel32), (double)fminf(gdn, gup) * (1.0 - kRel32));
        const double other = rest - 1.8e-7 * (double)mag;
        if (!(other > d1)) {
            L.match[pt] = m_pack(j1, -1);
            L.amb[atomicAdd(L.amb_n, 1)] = (unsigned short)pt;
        } else {
            L.match[pt] = m_pack(j1, j2);
            L.d2lb[pt] = f32_down(other);
            stamp_p0<DIM>(L, pt);
        }
    }
}

// lists of fewer chunks go to the split sweep (every warp takes a share of the target: lowest latency for a few points)
constexpr int kSlabMinChunks = 3;

template <int DIM, int NT, class SH>
__device__ __forceinline__ void nn_dispatch(const Loop<DIM>& L, SH& sh, int n_todo, float vox) {
    const int w = threadIdx.x >> 5;
    const int n_chunks = (n_todo + 31) >> 5;
    const int F = n_chunks > 0 ? (NT / 32) / n_chunks : 1;
    if (vox > 0.f && n_chunks >= kSlabMinChunks) {
        if (F >= 2) nn_slab<DIM, NT, true>(L, sh, n_todo, vox);
        else nn_slab<DIM, NT, false>(L, sh, n_todo, vox);
        return;
    }
    if (F >= 2) {
        if (n_chunks >= kSlabMinChunks) nn_split_tiles<DIM, NT>(L, sh, n_todo, n_chunks, F);
        else nn_split<DIM, NT>(L, sh, n_todo, n_chunks, F);
        return;
    }
    for (int base = 0; base < n_chunks; base += kSMax * (NT / 32)) {
        const int first = base + w;
        const int mine = first < n_chunks ? min(kSMax, (n_chunks - first + (NT / 32) - 1) / (NT / 32)) : 0;   // warp-uniform
        switch (mine) {
            case 0: break;
            case 1: nn_round<DIM, 1, NT>(L, sh, first, n_todo); break;
            case 2: nn_round<DIM, 2, NT>(L, sh, first, n_todo); break;
            case 3: nn_round<DIM, 3, NT>(L, sh, first, n_todo); break;
            case 4: nn_round<DIM, 4, NT>(L, sh, first, n_todo); break;
            case 5: nn_round<DIM, 5, NT>(L, sh, first, n_todo); break;
            case 6: nn_round<DIM, 6, NT>(L, sh, first, n_todo); break;
            case 7: nn_round<DIM, 7, NT>(L, sh, first, n_todo); break;
            default: nn_round<DIM, 8, NT>(L, sh, first, n_todo); break;
        }
    }
}

// Points whose runner-up bound was inconclusive: one warp per point scans the
// whole target in fp64 (lowest index wins exact ties) and keeps the two best.
template <int DIM, int NT, class SH>
__device__ __forceinline__ void resolve_ambiguous(const Loop<DIM>& L, const SH& sh) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int n_amb = sh.amb_n;
    for (int a = w; a < n_amb; a += (NT / 32)) {
        const int i = L.amb[a];
        const double px = L.cx[i], py = L.cy[i], pz = DIM == 3 ? L.cz[i] : 0.0;
        double best = INFINITY, second = INFINITY, third = INFINITY;
        int bj = 0x7fffffff, j2 = 0x7fffffff;
        auto offer = [&](double d, int j) {                     // (d, j) lexicographic: lowest index wins exact ties
            if (d < best || (d == best && j < bj)) { third = second; second = best; j2 = bj; best = d; bj = j; }
            else if (d < second || (d == second && j < j2)) { third = second; second = d; j2 = j; }
            else if (d < third) third = d;
        };
        for (int j = l; j < L.n_t; j += 32) offer(dist2_64<DIM>(L, px, py, pz, j), j);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, best, o);
            const double os = __shfl_xor_sync(0xffffffffu, second, o);
            const double ot = __shfl_xor_sync(0xffffffffu, third, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            const int oj2 = __shfl_xor_sync(0xffffffffu, j2, o);
            if (oj != 0x7fffffff) offer(od, oj);
            if (oj2 != 0x7fffffff) offer(os, oj2);
            third = fmin(third, ot);
        }
        if (l == 0) {
            L.match[i] = m_pack(bj, j2 == 0x7fffffff ? -1 : j2);
            L.d2lb[i] = f32_down(sqrt(third) * (1.0 - 1e-12));
            stamp_p0<DIM>(L, i);
        }
    }
}

// Many points at once whose fp32 ranking could not be separated (a source that sits metres or kilometres away from the
// target -- a registration that is diverging, SURVEY H1's cousins in the loop-closure batch -- has whole walls of
// near-equidistant candidates): one warp per point scanning the whole target in fp64 serialises them.  Instead a LANE
// takes a point and the warp walks the tiles in lock step: the fp32 tile minimum (one broadcast load per target, the
// sweep's arithmetic) filters, and only tiles that may hold a target as close as the best one known -- U, the exact
// distance to the fp32 argmin the failed decision left in match[] -- are evaluated in fp64.  A target in a skipped tile
// has fp32 d^2 > Tf = ((U + slack)(1 + kRel32))^2, so its exact distance exceeds U: it is neither the match nor the
// rival, and sqrt(Tf)(1 - kRel32) - slack bounds it from below for the carry-over test.
template <int DIM, int NT, class SH>
__device__ __forceinline__ void resolve_ambiguous_lockstep(const Loop<DIM>& L, const SH& sh) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_amb = sh.amb_n, n_chunks = (n_amb + 31) >> 5;
    const float* tf = reinterpret_cast<const float*>(L.t32);
    constexpr int kStride = DIM == 2 ? 2 : 4;
    for (int chunk = w; chunk < n_chunks; chunk += NT / 32) {
        const int q = chunk * 32 + lane;
        const bool live = q < n_amb;
        const int i = L.amb[live ? q : chunk * 32];               // idle lanes shadow the chunk's first point
        const double px = L.cx[i], py = L.cy[i], pz = DIM == 3 ? L.cz[i] : 0.0;
        const float sx = (float)(px - L.c0), sy = (float)(py - L.c1), sz = DIM == 3 ? (float)(pz - L.c2) : 0.f;
        const int j0 = m_idx<false>(L.match[i]);
        double best = dist2_64<DIM>(L, px, py, pz, j0), second = INFINITY, third = INFINITY;
        int bj = j0, j2 = 0x7fffffff;
        const double slack = 1.8e-7 * (double)(fabsf(sx) + fabsf(sy) + (DIM == 3 ? fabsf(sz) : 0.f) + L.ta);
        const double reach = (sqrt(best) + slack) * (1.0 + kRel32);
        const float Tf = __double2float_ru(reach * reach);
        auto offer = [&](double d, int j) {                     // (d, j) lexicographic: lowest index wins exact ties
            if (d < best || (d == best && j < bj)) { third = second; second = best; j2 = bj; best = d; bj = j; }
            else if (d < second || (d == second && j < j2)) { third = second; second = d; j2 = j; }
            else if (d < third) third = d;
        };
        for (int tile = 0; tile < L.n_tiles; ++tile) {
            float tm = INFINITY;
            const float* p = tf + (size_t)tile * 32 * kStride;
#pragma unroll 8
            for (int u = 0; u < 32; ++u) {
                const float dx = sx - p[u * kStride], dy = sy - p[u * kStride + 1];
                float d = fmaf(dy, dy, dx * dx);
                if (DIM == 3) { const float dz = sz - p[u * kStride + 2]; d = fmaf(dz, dz, d); }
                tm = fminf(tm, d);
            }
            if (tm <= Tf) {
                const int je = min(tile * 32 + 32, L.n_t);
                for (int j = tile * 32; j < je; ++j)
                    if (j != j0) offer(dist2_64<DIM>(L, px, py, pz, j), j);
            }
        }
        if (live) {
            const double skipped = sqrt((double)Tf) * (1.0 - kRel32) - slack;          // every target in a tile that was skipped
            L.match[i] = m_pack(bj, j2 == 0x7fffffff ? -1 : j2);
            L.d2lb[i] = f32_down(fmin(sqrt(third) * (1.0 - 1e-12), skipped));
            stamp_p0<DIM>(L, i);
        }
    }
}

// Far field.  When the source has left the target's neighbourhood altogether -- a diverged registration keeps iterating to
// max_iterations in the reference, icp.py:177-223, with every point hundreds or millions of metres from every target --
// fp32 cannot rank the candidates at all, and nothing is pruned.  But then almost no target can be anybody's nearest
// neighbour: with q0 = source point 0, D = q0 - c (c the target's box centre), u = q - q0, v = t - c,
//        |q - t|^2 = |D|^2 + 2 D.u + |u - v|^2 - 2 D.v ,
// so for two targets a, b and ANY source point q:  d^2(q,a) - d^2(q,b) = (w_a - w_b) + (|u - v_a|^2 - |u - v_b|^2) with
// w = -2 D.v and the bracket inside [-(S + T)^2, (S + T)^2] (S = max |u|, T >= max |v|).  A target whose w exceeds the
// minimum by more than (S + T)^2 is therefore never a nearest neighbour: only the "front" of the target facing the
// source survives -- a handful of points -- and every source point is decided among them by exact fp64 distances
// (lowest index on exact ties).  Returns false (nothing written) when the source is not that far or the front is large.
constexpr int kFrontMax = 128;

template <int DIM, int NT, class SH>
__device__ __forceinline__ bool far_field_matches(const Loop<DIM>& L, SH& sh, int& phase) {
    const int tid = threadIdx.x;
    const double q0x = L.cx[0], q0y = L.cy[0], q0z = DIM == 3 ? L.cz[0] : 0.0;
    const double Dx = q0x - L.c0, Dy = q0y - L.c1, Dz = DIM == 3 ? q0z - L.c2 : 0.0;
    const double Dn2 = Dx * Dx + Dy * Dy + Dz * Dz;
    const double T = (double)L.ta;                              // sum over axes of max |v|: >= max |v|
    if (!(Dn2 > 4096.0 * T * T) || !(Dn2 < 1e300)) return false;                  // uniform: every thread reads the same words
    double m[1] = {0.0};
    for (int i = tid; i < L.n_s; i += NT) {
        const double ux = L.cx[i] - q0x, uy = L.cy[i] - q0y, uz = DIM == 3 ? L.cz[i] - q0z : 0.0;
        m[0] = fmin(m[0], -(ux * ux + uy * uy + uz * uz));
    }
    block_reduce<1, MinOp>(m, sh, phase);
    const double S = sqrt(-m[0]);
    const double Dn = sqrt(Dn2);
    if (!(Dn > 64.0 * (S + T))) return false;
    // (S + T)^2, plus what the fp64 roundings of w and of the distances themselves (|d^2| <= (Dn + S + T)^2) can move
    const double E2 = (S + T) * (S + T) * (1.0 + 1e-9) + 2e-15 * (Dn + S + T) * (Dn + S + T);
    auto w_of = [&](int j) {
        double x, y, z;
        tgt_at<DIM, false>(L, j, x, y, z);
        return -2.0 * (Dx * (x - L.c0) + Dy * (y - L.c1) + (DIM == 3 ? Dz * (z - L.c2) : 0.0));
    };
    double wm[1] = {INFINITY};
    for (int j = tid; j < L.n_t; j += NT) wm[0] = fmin(wm[0], w_of(j));
    if (tid == 0) sh.front_n = 0;
    block_reduce<1, MinOp>(wm, sh, phase);                      // its barriers also publish front_n = 0
    for (int j = tid; j < L.n_t; j += NT)
        if (w_of(j) <= wm[0] + E2) {
            const int k = atomicAdd(&sh.front_n, 1);
            if (k < kFrontMax) L.amb[k] = (unsigned short)j;   // the ambiguity list is idle between iterations
        }
    __syncthreads();
    const int nf = sh.front_n;
    if (nf > kFrontMax || nf > L.cap) return false;
    for (int i = tid; i < L.n_s; i += NT) {
        const double px = L.cx[i], py = L.cy[i], pz = DIM == 3 ? L.cz[i] : 0.0;
        double best = INFINITY;
        int bj = 0x7fffffff;
        for (int k = 0; k < nf; ++k) {
            const int j = L.amb[k];
            const double d = dist2_64<DIM>(L, px, py, pz, j);
            if (d < best || (d == best && j < bj)) { best = d; bj = j; }
        }
        L.match[i] = m_pack(bj, -1);
        L.d2lb[i] = -1.f;                                       // decided afresh in every iteration spent out here
    }
    __syncthreads();
    return true;
}

// The points to decide, ordered by their CURRENT x for the slab sweep.  The classify pass lists them by source index,
// i.e. by x in the source's own frame; once the registration has rotated the source -- loop-closure candidates start
// up to half a radian apart -- 32 consecutive ones spread over metres of x and a warp's walk covers most of the target.
// A counting sort into 64 bins of the target's x range (two passes over the list, one 64-entry scan) restores the
// locality; the order inside a bin is whatever the atomics gave, which is irrelevant to an exact search.
template <int DIM, int NT, class SH>
__device__ __forceinline__ void order_todo_by_x(Loop<DIM>& L, SH& sh, int n_todo, double xlo, double xhi) {
    const int tid = threadIdx.x;
    if (tid < 65) sh.bins[tid] = 0;
    __syncthreads();
    const double scale = xhi > xlo ? 64.0 / (xhi - xlo) : 0.0;
    auto bin_of = [&](int i) {
        const double b = (L.cx[i] - xlo) * scale;
        return b > 0.0 ? (b < 63.0 ? (int)b : 63) : 0;          // NaN -> 0
    };
    for (int q = tid; q < n_todo; q += NT) atomicAdd(&sh.bins[bin_of(L.todo[q])], 1);
    __syncthreads();
    if (tid < 32) {                                            // exclusive scan of 64 counts by one warp
        const int a0 = sh.bins[2 * tid], a1 = sh.bins[2 * tid + 1];
        int inc = a0 + a1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= o) inc += t;
        }
        sh.bins[2 * tid] = inc - a0 - a1;
        sh.bins[2 * tid + 1] = inc - a1;
    }
    __syncthreads();
    for (int q = tid; q < n_todo; q += NT) {
        const int i = L.todo[q];
        L.todo2[atomicAdd(&sh.bins[bin_of(i)], 1)] = (unsigned short)i;
    }
    __syncthreads();
    L.list = L.todo2;
}

// ---- grid-mode nearest neighbour (big targets): exact fp64 ring search ------------------
// One thread per source point that needs a decision.  Cells are visited ring by
// ring around the query's cell; after ring r every unvisited target lies outside
// the (2r+1)^2 block, i.e. at least `bound` away, so the search stops as soon as
// the best distance found is below that bound.  The runner-up distance (needed by
// the carry-over test) is min(second best seen, bound).  Queries that would need
// more than kGridMaxRing rings fall back to a scan of the whole target.
constexpr int kGridMaxRing = 24;

template <int DIM, int NT>
__device__ __forceinline__ void grid_nn(const Loop<DIM>& L, int n_todo, unsigned long long (&gst)[3]) {
    const BigGrid& G = L.grid;
    for (int q = threadIdx.x; q < n_todo; q += NT) {
        const int i = L.todo[q];
        const double px = L.cx[i], py = L.cy[i];
        // the query may lie outside the target's bounding box: clamp its cell, keep exact bounds
        const double fx = (px - G.lox) / G.h, fy = (py - G.loy) / G.h;
        const int cxi = min(G.nx - 1, max(0, (int)floor(fx)));
        const int cyi = min(G.ny - 1, max(0, (int)floor(fy)));
        double best = INFINITY, second = INFINITY;
        int bj = 0x7fffffff;
        bool done = false;
        double bound = 0.0;
        for (int r = 0; r <= kGridMaxRing; ++r) {
            const int x0 = cxi - r, x1 = cxi + r, y0 = cyi - r, y1 = cyi + r;
            for (int y = max(y0, 0); y <= min(y1, G.ny - 1); ++y) {
                const bool edge_row = (y == y0) || (y == y1);
                const int xstep = edge_row ? 1 : max(2 * r, 1);
                for (int x = x0; x <= x1; x += xstep) {
                    if (x < 0 || x >= G.nx) continue;
                    const unsigned b = big_cell_hash(x, y) & G.mask;
                    const int beg = b ? G.start[b - 1] : 0, end = G.start[b];
                    ++gst[2];
                    for (int e = beg; e < end; ++e) {
                        const int2 cc = G.cell[e];
                        if (cc.x != x || cc.y != y) continue;          // another cell hashed to this bucket
                        const int j = G.items[e];
                        ++gst[1];
                        const double d = dist2_64<DIM, true>(L, px, py, 0.0, j);
                        if (d < best || (d == best && j < bj)) { second = best; best = d; bj = j; }
                        else if (d < second) second = d;
                    }
                }
            }
            if ((x0 <= 0) && (y0 <= 0) && (x1 >= G.nx - 1) && (y1 >= G.ny - 1)) { done = true; bound = INFINITY; break; }
            bound = INFINITY;
            if (x0 > 0) bound = fmin(bound, px - (G.lox + x0 * G.h));
            if (x1 < G.nx - 1) bound = fmin(bound, (G.lox + (x1 + 1) * G.h) - px);
            if (y0 > 0) bound = fmin(bound, py - (G.loy + y0 * G.h));
            if (y1 < G.ny - 1) bound = fmin(bound, (G.loy + (y1 + 1) * G.h) - py);
            bound = bound * (1.0 - 1e-9) - 1e-12 * G.h;
            if (bound > 0.0 && best < bound * bound) { done = true; break; }
        }
        if (!done) {                                   // far from everything: exact scan of the whole target
            best = INFINITY; second = INFINITY; bj = 0x7fffffff; bound = INFINITY;
            for (int j = 0; j < L.n_t; ++j) {
                const double d = dist2_64<DIM, true>(L, px, py, 0.0, j);
                if (d < best) { second = best; best = d; bj = j; }
                else if (d < second) second = d;
            }
        }
        L.match[i] = bj;
        L.d2lb[i] = f32_down(fmin(sqrt(second), bound) * (1.0 - 1e-12));
        stamp_p0<DIM>(L, i);
        ++gst[0];
    }
}

// ---- K3: the kernel ----------------------------------------------------------------
// MINB = CTAs per SM the register allocation must allow: 2 for the 2-D bulk launch (3 for 3-D and grid mode), 1 for the
// hand-over launch (one CTA per SM anyway); at 1 and 2 everything stays in registers instead of spilling.
// CL = CTAs of a thread-block cluster (1: no cluster).  The CL = 4 variant is the hand-over launch of the 2-D brute-force
// path.  Its CTAs take pairs from the queue independently, like the plain variant; the launch ends with a handful of
// pairs that re-decide hundreds of points in every one of 150 iterations -- 3-6 ms chains on one SM each, while the SMs
// around them have run out of work (a rank of an 8-GPU job holds 5-20 of them and 1.5-2 ms of work per SM).  A CTA that
// finds the queue empty therefore looks for a cluster mate that is still running a pair, stages the same target and
// takes a share of every bulk nearest-neighbour sweep of that pair: it reads the points to decide from the owner's shared
// memory and writes the decisions back (distributed shared memory).  Owner and helpers meet through three words in the
// owner's shared memory: coop_word (helpers register with a compare-and-swap while the pair is open), coop_pub (one
// 64-bit publication per shared sweep: epoch, helper mask, list length) and coop_arrived (helpers done).  The decisions
// are the same exact nearest neighbours whoever computes them, so the results do not depend on who helped.
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
constexpr unsigned kCoopOpen = 0x80000000u;
constexpr int kCoopMinShare = 96;           // sweeps of fewer points are not worth a publication
constexpr int kCoopCluster = 8;             // CTAs per cluster of the hand-over launch
// The hand-over of a 2-D brute-force batch is launched twice, side by side: as clusters (a.coop_ctas CTAs) and as plain
// CTAs (one per SM).  Which of the two does the work is decided here, from what the bulk launch handed over, the same in
// every CTA of both launches:
//   - cluster mode when long chains, not throughput, will bound the launch: a few pairs that re-decided 96+ points per
//     iteration when they were parked among at most a.coop_factor (10) handed-over pairs per cluster CTA (a rank's share
//     of a multi-GPU batch).
//     The plain launch then only fills the SMs the cluster shape leaves unused.  The cluster kernel runs every phase ~12 %
//     slower than the plain one (register pressure), which a chain-bound launch does not feel;
//   - plain mode otherwise (C2 and C5 on one GPU: throughput-bound): the cluster CTAs leave at once.
// In cluster mode class 0 of the handed-over pairs (256+ points per iteration) are the likely chains: when they are few,
// each gets its cluster's rank 0 and `h` mates that help from the first iteration on instead of taking pairs of their own.
// *skip = queue positions dealt statically to the cluster CTAs (the dynamic queue continues behind them).
__device__ __forceinline__ bool coop_plan(const IcpArgs& a, unsigned& n_chain, unsigned& h, unsigned& skip) {
    const unsigned n_cl_ctas = (unsigned)a.coop_ctas, n_clusters = n_cl_ctas / kCoopCluster;
    n_chain = 0; h = 0; skip = 0;
    if (n_cl_ctas == 0u) return false;
    const unsigned total = a.cont_count[0], heavy = a.cont_count[1] + a.cont_count[2];
    if (heavy < 4u || total > (unsigned)a.coop_factor * n_cl_ctas) return false;
    n_chain = min(a.cont_count[1], n_clusters);
    if (n_chain > 0) h = min((unsigned)kCoopCluster, n_cl_ctas / (3u * n_chain));
    h = h > 0 ? h - 1 : 0;
    skip = n_cl_ctas - n_chain * h;
    return true;
}
// a wait that is never satisfied is a protocol bug: fail the launch instead of hanging the device
#define ICPB_COOP_SPIN(cond)                                            \
    for (unsigned spin_ = 0; !(cond); ++spin_) {                        \
        __nanosleep(64);                                                \
        if (spin_ > (1u << 26)) __trap();                               \
    }

// A CTA of the cluster variant helping a mate (see icp_pairs_kernel): the target staged in L is this CTA's own copy, the
// source side (points, decisions, lists) is the owner's, reached through distributed shared memory.  Not inlined: kept
// out of the kernel's register allocation (inlined, the 512-thread variant spilled five times as much in its own loop).
template <int DIM, int NT, class SH>
__device__ __forceinline__ void coop_help(const Loop<DIM>& L, SH& sh, unsigned cl_rank, float slab_vox, unsigned long long* stats) {
    namespace cg = cooperative_groups;
    const int tid = threadIdx.x;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned orank = (unsigned)sh.help_rank;
    SH* osh = cluster.map_shared_rank(&sh, orank);
    Loop<DIM> H = L;
    H.cx = cluster.map_shared_rank(L.cx, orank); H.cy = cluster.map_shared_rank(L.cy, orank);
    H.cz = cluster.map_shared_rank(L.cz, orank);
    H.match = cluster.map_shared_rank(L.match, orank); H.d2lb = cluster.map_shared_rank(L.d2lb, orank);
    H.p0 = cluster.map_shared_rank(L.p0, orank); H.amb = cluster.map_shared_rank(L.amb, orank);
    H.amb_n = &osh->amb_n; H.slab_evals = &osh->slab_evals;
    const unsigned short* o_todo = cluster.map_shared_rank(L.todo, orank);
    const unsigned short* o_todo2 = cluster.map_shared_rank(L.todo2, orank);
    const unsigned my_bit = 1u << cl_rank;
    unsigned seen = 0;
    if (tid == 0) {
        // register -- only while the pair this target was staged for is still the one running there
        seen = (unsigned)(*(volatile unsigned long long*)&osh->coop_pub >> 32);
        const unsigned want = sh.help_word & ~0xffu;
        int ok = 0;
        for (;;) {
            const unsigned cur = *(volatile unsigned*)&osh->coop_word;
            if ((cur & ~0xffu) != want) break;
            if (atomicCAS(&osh->coop_word, cur, cur | my_bit) == cur) { ok = 1; break; }
        }
        sh.help_ok = ok;
        if (ok && stats) atomicAdd(&stats[15], 1ull);      // statistics: helpers that joined a pair
    }
    __syncthreads();
    while (sh.help_ok) {
        if (tid == 0) {
            unsigned long long pub;
            ICPB_COOP_SPIN((unsigned)((pub = *(volatile unsigned long long*)&osh->coop_pub) >> 32) != seen);
            seen = (unsigned)(pub >> 32);
            fence_cluster();
            sh.help_pub = pub;
        }
        __syncthreads();
        const unsigned long long pub = sh.help_pub;
        const unsigned mask = (unsigned)(pub >> 20) & 0xffu;
        const int flags = (int)(pub >> 16) & 0xf, n = (int)(pub & 0xffffu);
        __syncthreads();                               // (sh.help_pub is rewritten in the next round)
        if (!(mask & my_bit)) continue;                // published before this CTA registered
        if (!(flags & 1)) {
            const int parts = __popc(mask) + 1, k = __popc(mask & (my_bit - 1u)) + 1;      // the owner is part 0
            const int per = ((n + parts * 32 - 1) / (parts * 32)) * 32;
            const int lo = k * per, cnt = min(per, n - lo);
            H.list = ((flags & 4) ? o_todo2 : o_todo) + lo;
            if (cnt > 0) nn_dispatch<DIM, NT>(H, sh, cnt, (flags & 2) ? slab_vox : 0.f);
            __syncthreads();
        }
        if (tid == 0) { fence_cluster(); atomicAdd(&osh->coop_arrived, 1u); }
        if (flags & 1) break;                          // the pair is finished
    }
}

template <int DIM, bool GRID, int MINB, int NT, int CL = 1>
__global__ void __launch_bounds__(NT, MINB) icp_pairs_kernel(const IcpArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    using SH = CtaSharedT<NT / 32>;
    SH& sh = *reinterpret_cast<SH*>(smem);
    const int tid = threadIdx.x;
    namespace cg = cooperative_groups;
    // (the cluster variant keeps its few words of state in shared memory: the kernel sits at the register limit)
    const int tstride = a.cap_t + a.cap_t / 32;
    Loop<DIM> L;
    {
        unsigned char* q = smem + align16(sizeof(SH));
        double* tgt64 = reinterpret_cast<double*>(q);
        L.tx = tgt64; L.ty = tgt64 + tstride; L.tz = tgt64 + 2 * (size_t)tstride;
        q += sizeof(double) * DIM * (size_t)tstride;
        q = smem + align16((size_t)(q - smem));
        L.t32 = reinterpret_cast<float4*>(q);          q += sizeof(float) * (DIM == 2 ? 2 : 4) * (size_t)a.cap_t;
        L.cx = reinterpret_cast<double*>(q);           q += sizeof(double) * (size_t)a.cap_s;
        L.cy = reinterpret_cast<double*>(q);           q += sizeof(double) * (size_t)a.cap_s;
        L.cz = reinterpret_cast<double*>(q);           if (DIM == 3) q += sizeof(double) * (size_t)a.cap_s;
        L.match = reinterpret_cast<int*>(q);           q += sizeof(int) * (size_t)a.cap_s;
        L.d2lb = reinterpret_cast<float*>(q);          q += sizeof(float) * (size_t)a.cap_s;
        L.p0 = reinterpret_cast<float*>(q);            q += sizeof(float) * DIM * (size_t)a.cap_s;
        L.cap = a.cap_s;
        L.todo = reinterpret_cast<unsigned short*>(q); q += sizeof(unsigned short) * (size_t)a.cap_s;
        L.todo2 = reinterpret_cast<unsigned short*>(q); q += sizeof(unsigned short) * (size_t)a.cap_s;
        L.amb = reinterpret_cast<unsigned short*>(q);       q += sizeof(unsigned short) * (size_t)a.cap_s;
        L.list = L.todo;
        L.amb_n = &sh.amb_n;
        L.slab_evals = &sh.slab_evals;
        L.nrm_s = nullptr;
        if (NT == 512 && DIM == 2 && !GRID) L.nrm_s = reinterpret_cast<const double2*>(smem + align16((size_t)(q - smem)));
    }
    double* tx = const_cast<double*>(L.tx);
    double* ty = const_cast<double*>(L.ty);
    double* tz = const_cast<double*>(L.tz);
    int phase = 0;
    unsigned long long st_evals = 0, st_amb = 0, st_iters = 0, st_swept = 0, st_kept = 0, st_far = 0;   // thread 0 only
    // thread 0's cycles per phase over iterations >= 8 (the nearly-converged regime): profiling aid
    long long ph[6] = {0, 0, 0, 0, 0, 0}, ph_t = 0;
    unsigned long long ph_iters = 0;
    unsigned long long gst[3] = {0, 0, 0};         // grid mode, per thread: queries, candidates evaluated, cells visited

    __shared__ __align__(8) unsigned long long nrm_bar;      // completion of the normals' bulk copy (512-thread variant)
    unsigned nrm_parity = 0;
    if (tid == 0) sh.ph_seen[0] = sh.ph_seen[1] = sh.ph_seen[2] = 0;      // (the loop below opens with a barrier)
    if (CL > 1) {
        if (tid == 0) { sh.coop_word = 0u; sh.coop_idle = 0u; sh.coop_arrived = 0u; sh.coop_pub = 0ull; sh.coop_parts = 1; sh.coop_load = 0;
                        sh.help_rank = -1; sh.help_dry = 0; sh.coop_first = 1; sh.coop_seq = 0u; sh.coop_epoch = 0u; }
        cg::this_cluster().sync();      // no mate reads these words before they are set
    }
    if (NT == 512 && DIM == 2 && !GRID) {
        if (tid == 0) mbar_init(&nrm_bar, 1);
        __syncthreads();
    }
    for (;;) {
        __syncthreads();
        // phase 1 takes pairs in order; phase 2 (a.resume) takes the pairs phase 1 handed over
        if (tid == 0) {
            const unsigned cl_rank = CL > 1 ? cg::this_cluster().block_rank() : 0u;
            if (CL > 1) sh.help_rank = -1;
            if (CL > 1 && sh.help_dry) {
                sh.pair = 0xffffffffu;
            } else if (a.resume) {
                // the handed-over pairs of this launch's classes, most expensive class first (longest processing time first:
                // the last CTAs to finish should be running short pairs, not a 150-iteration chain of full sweeps)
                unsigned n_chain, h, skip;
                const bool cluster_mode = coop_plan(a, n_chain, h, skip);
                bool joined = false;
                unsigned q;
                // the launch that is not in charge leaves: every cluster CTA in plain mode, the plain CTAs beyond the SMs
                // the cluster launch leaves free in cluster mode (all of them when it leaves none)
                const bool leave = CL > 1 ? !cluster_mode
                                          : (cluster_mode && (int)blockIdx.x >= (int)gridDim.x - a.coop_ctas);
                if (leave) {
                    q = 0xffffffffu;
                } else if (CL > 1 && sh.coop_first) {
                    // cluster variant, first pair: the head of the queue is dealt statically, rank 0 of cluster c takes
                    // position c -- one of the most expensive pairs per cluster, so that every long chain has mates that
                    // run dry and come to help -- then the other ranks in turn, minus the CTAs reserved for the chains
                    const unsigned n_clusters = gridDim.x / CL, c = blockIdx.x / CL, r = cl_rank;
                    if (r >= 1 && r <= h && c < n_chain) {
                        // reserved: help rank 0 with its chain from its first iteration on (if it is one: long lists)
                        cg::cluster_group cluster = cg::this_cluster();
                        SH* m = cluster.map_shared_rank(&sh, 0);
                        unsigned w = 0;
                        int load = 0;
                        ICPB_COOP_SPIN((((w = *(volatile unsigned*)&m->coop_word) & kCoopOpen) &&
                                        (load = *(volatile int*)&m->coop_load) > 0) || *(volatile unsigned*)&m->coop_idle);
                        if ((w & kCoopOpen) && load >= kCoopMinShare) {
                            fence_cluster();
                            sh.help_rank = 0; sh.help_word = w;
                            sh.pair = *(volatile unsigned*)&m->coop_pair;
                            sh.bcast_i[2] = *(volatile int*)&m->coop_slot;
                            joined = true;
                        }
                        q = joined ? 0u : skip + atomicAdd(a.queue, 1u);
                    } else {
                        q = c;
                        if (r >= 1) {
                            q = n_clusters + c - (r <= h ? n_chain : 0u);
                            for (unsigned rr = 1; rr < r; ++rr) q += n_clusters - (rr <= h ? n_chain : 0u);
                        }
                    }
                } else {
                    q = skip + atomicAdd(a.queue, 1u);
                }
                if (CL > 1) sh.coop_first = 0;
                if (!joined) {
                    unsigned rest = q, b = (unsigned)a.class_lo;
                    while (b < (unsigned)a.class_hi && rest >= a.cont_count[1 + b]) { rest -= a.cont_count[1 + b]; ++b; }
                    const int slot = b < (unsigned)a.class_hi ? a.cont_bucket[(size_t)b * a.cont_cap + rest] : -1;
                    sh.pair = slot >= 0 ? (unsigned)a.cont_list[slot] : 0xffffffffu;
                    sh.bcast_i[2] = slot;               // where the pair's state was parked
                    if (CL > 1 && slot < 0) sh.help_dry = 1;
                }
            } else {
                const unsigned q = atomicAdd(a.queue, 1u);
                sh.pair = q < (unsigned)a.n_pairs ? (a.pair_order ? (unsigned)a.pair_order[a.pair_first + q] : (unsigned)a.pair_first + q)
                                                  : 0xffffffffu;
                sh.bcast_i[2] = (int)q;
            }
        }
        if (CL > 1 && tid == 0) {
            const unsigned cl_rank = cg::this_cluster().block_rank();
            if (sh.help_rank < 0 && sh.pair == 0xffffffffu) {
                // nothing left in the queue: help a cluster mate that is still running a pair (the one with the fewest
                // helpers), wait while a mate is between pairs, leave when every mate is as idle as this CTA
                *(volatile unsigned*)&sh.coop_idle = 1u;
                cg::cluster_group cluster = cg::this_cluster();
                for (unsigned spin = 0;; ++spin) {
                    bool busy = false;
                    int best = -1, best_n = 0;
                    unsigned best_w = 0;
                    for (unsigned r = 0; r < (unsigned)CL; ++r) {
                        if (r == cl_rank) continue;
                        SH* m = cluster.map_shared_rank(&sh, r);
                        const unsigned w = *(volatile unsigned*)&m->coop_word;
                        if (w & kCoopOpen) {
                            // worth helping: long lists; between several, the most points per CTA already on it
                            const int load = *(volatile int*)&m->coop_load;
                            const int n = load >= kCoopMinShare ? load / (__popc(w & 0xffu) + 1) : 0;
                            if (n > best_n) { best = (int)r; best_n = n; best_w = w; }
                            if (n == 0) busy = true;           // a light pair for now: it may turn heavy, or end
                        } else if (*(volatile unsigned*)&m->coop_idle == 0u) {
                            busy = true;
                        }
                    }
                    if (best >= 0) {
                        fence_cluster();
                        SH* m = cluster.map_shared_rank(&sh, (unsigned)best);
                        sh.help_rank = best; sh.help_word = best_w;
                        sh.pair = *(volatile unsigned*)&m->coop_pair;
                        sh.bcast_i[2] = *(volatile int*)&m->coop_slot;
                        break;
                    }
                    if (!busy) break;
                    __nanosleep(256);
                    if (spin > (1u << 26)) __trap();
                }
            }
        }
        __syncthreads();
        if (sh.pair == 0xffffffffu) break;
        const int p = (int)sh.pair;
        const int slot_in = sh.bcast_i[2];
        const long long pair_t0 = (tid == 0 && a.pair_prof) ? clock64() : 0;
        const unsigned long long pp_swept0 = st_swept, pp_amb0 = st_amb, pp_it0 = st_iters;
        unsigned long long st_recent = 0;              // thread 0: points swept in the last iterations before a hand-over

        const int cs = a.src_idx ? a.src_idx[p] : p;
        const int ct = a.tgt_idx ? a.tgt_idx[p] : p;
        const bool named = cs >= 0 && cs < a.s.n_clouds && ct >= 0 && ct < a.t.n_clouds;
        const int n_s = named ? a.s.ds_n[cs] : 0, n_t = named ? a.t.ds_n[ct] : 0;
        if (n_s <= 0 || n_t <= 0 || n_s > a.cap_s || (!GRID && n_t > a.cap_t)) {
            if (tid == 0) {
                for (int k = 0; k < DIM * DIM; ++k) a.R_out[(size_t)p * DIM * DIM + k] = (k % (DIM + 1) == 0) ? 1.0 : 0.0;
                for (int k = 0; k < DIM; ++k) a.t_out[(size_t)p * DIM + k] = 0.0;
                a.err_out[p] = INFINITY;
                if (a.prev_out) a.prev_out[p] = INFINITY;
                a.iters_out[p] = 0;
                a.status_out[p] = ICPB200_BAD_VOXELS;
            }
            continue;
        }
        const double* src_ds = a.s.ds + a.s.off[cs] * DIM;
        const double* tgt_ds = a.t.ds + a.t.off[ct] * DIM;
        const double* normals = a.t.nrm ? a.t.nrm + a.t.off[ct] * 2 : nullptr;
        const double* box = a.t.box + (size_t)ct * 6;
        const double tgt_xlo = box[0], tgt_xhi = box[3];
        // 512-thread variant: the target's normals come into shared memory by one TMA bulk copy issued now and awaited in
        // front of the first accumulation -- it lands under the staging of the clouds and the first nearest-neighbour sweep
        const bool nrm_tma = NT == 512 && DIM == 2 && !GRID && normals != nullptr && a.method == ICPB200_POINT_TO_LINE && a.tma_normals;
        bool nrm_pending = false;
        if (nrm_tma) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the last pair's reads of this buffer are done
                tma_bulk_g2s(const_cast<double2*>(L.nrm_s), normals, (unsigned)n_t * 16u, &nrm_bar);
            }
            nrm_pending = true;
        }
        L.n_s = n_s; L.n_t = n_t; L.n_tiles = (n_t + 31) / 32;
        L.c0 = 0.5 * (box[0] + box[3]); L.c1 = 0.5 * (box[1] + box[4]); L.c2 = 0.5 * (box[2] + box[5]);
        L.tg = tgt_ds;
        if (GRID) L.grid = a.grids[ct];

        if (tid == 0) {
            // icp.py:153-160
            const bool init = a.R_init != nullptr && a.t_init != nullptr;
            for (int k = 0; k < DIM * DIM; ++k)
                sh.r_tot[k] = init ? a.R_init[(size_t)p * DIM * DIM + k] : ((k % (DIM + 1) == 0) ? 1.0 : 0.0);
            for (int k = 0; k < DIM; ++k) sh.t_tot[k] = init ? a.t_init[(size_t)p * DIM + k] : 0.0;
            sh.amb_n = 0;
            sh.slab_evals = 0u;
            sh.bcast_i[0] = 0;                     // todo counter
            sh.slab_off = 0;
            if (DIM == 3) for (int k = 0; k < 9; ++k) sh.kab_v[k] = (k % 4 == 0) ? 1.0 : 0.0;
        }
        // ---- stage the target: fp64 (tile-padded SoA) and recentred fp32
        if (GRID) {
            __syncthreads();                        // publishes sh.r_tot
            L.ta = 0.f;
        } else {
            float* f = reinterpret_cast<float*>(L.t32);
            constexpr int per = DIM == 2 ? 2 : 4;
            double ext[DIM];
#pragma unroll
            for (int k = 0; k < DIM; ++k) ext[k] = 0.0;
            for (int j = tid; j < L.n_tiles * 32; j += NT) {
                if (j < n_t) {
                    const int jp = pad_index(j);
                    const double x = tgt_ds[(size_t)j * DIM], y = tgt_ds[(size_t)j * DIM + 1];
                    tx[jp] = x; ty[jp] = y;
                    const double v0 = x - L.c0, v1 = y - L.c1;
                    f[j * per] = (float)v0; f[j * per + 1] = (float)v1;
                    ext[0] = fmax(ext[0], fabs(v0)); ext[1] = fmax(ext[1], fabs(v1));
                    if (DIM == 3) {
                        const double z = tgt_ds[(size_t)j * DIM + 2];
                        tz[jp] = z;
                        const double v2 = z - L.c2;
                        f[j * per + 2] = (float)v2; f[j * per + 3] = 0.f;
                        ext[DIM - 1] = fmax(ext[DIM - 1], fabs(v2));
                    }
                } else {
                    f[j * per] = kFar; f[j * per + 1] = kFar;
                    if (DIM == 3) { f[j * per + 2] = kFar; f[j * per + 3] = 0.f; }
                }
            }
            double neg[DIM];
#pragma unroll
            for (int k = 0; k < DIM; ++k) neg[k] = -ext[k];
            block_reduce<DIM, MinOp>(neg, sh, phase);       // includes the barrier that publishes sh.r_tot
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) s += -neg[k];
            L.ta = (float)(s * 1.0000002);
        }
        {   // icp.py:153-160: transformed = source @ R_init.T + t_init (or a copy)
            double r[9], t[3];
            for (int k = 0; k < DIM * DIM; ++k) r[k] = sh.r_tot[k];
            for (int k = 0; k < DIM; ++k) t[k] = sh.t_tot[k];
            for (int i = tid; i < n_s; i += NT) {
                const double x = src_ds[(size_t)i * DIM], y = src_ds[(size_t)i * DIM + 1];
                if (DIM == 2) {
                    L.cx[i] = x * r[0] + y * r[1] + t[0];
                    L.cy[i] = x * r[2] + y * r[3] + t[1];
                } else {
                    const double z = src_ds[(size_t)i * DIM + 2];
                    L.cx[i] = x * r[0] + y * r[1] + z * r[2] + t[0];
                    L.cy[i] = x * r[3] + y * r[4] + z * r[5] + t[1];
                    L.cz[i] = x * r[6] + y * r[7] + z * r[8] + t[2];
                }
                L.d2lb[i] = -1.f;                  // no decision yet: forces a sweep
                L.match[i] = 0;
            }
        }
        double prev = INFINITY, err = INFINITY;
        int iters = 0;
        if (a.resume) {                            // continue where phase 1 stopped
            const size_t so = (size_t)slot_in * a.cap_s;
            const double* cc = a.cont_cur + so * DIM;
            for (int i = tid; i < n_s; i += NT) {
                L.cx[i] = cc[i]; L.cy[i] = cc[(size_t)a.cap_s + i];
                if (DIM == 3) L.cz[i] = cc[2 * (size_t)a.cap_s + i];
                L.match[i] = a.cont_match[so + i];
                L.d2lb[i] = a.cont_d2lb[so + i];
                for (int k = 0; k < DIM; ++k) L.p0[k * L.cap + i] = a.cont_moved[so * DIM + (size_t)k * a.cap_s + i];
            }
            const double* sc = a.cont_scalar + (size_t)slot_in * 16;
            prev = sc[12]; err = sc[13]; iters = (int)sc[14];
            __syncthreads();
            if (tid == 0) {
                for (int k = 0; k < DIM * DIM; ++k) sh.r_tot[k] = sc[k];
                for (int k = 0; k < DIM; ++k) sh.t_tot[k] = sc[9 + k];
            }
        }
        __syncthreads();

        // ---- iterate (icp.py:175-220)
        const bool p2l = (a.method == ICPB200_POINT_TO_LINE) && DIM == 2 && normals != nullptr;
        const bool gated = a.max_corr >= 0.0;
        const double gate2 = a.max_corr * a.max_corr;                 // icp.py:169
        const int min_inl = max(3, n_s / 10);                         // icp.py:186
        int status = ICPB200_MAX_ITER;
        bool handed_over = false;
        // clouds come out of K1 ordered by voxel column: x is sorted up to one voxel (plus fp32 rounding)
        const float slab_vox = a.brute_slab ? __double2float_ru(a.voxel * 1.000001) + 1e-6f * L.ta : 0.f;
        if (CL > 1 && sh.help_rank < 0) {
            // this pair takes helpers from now on
            if (tid == 0) {
                sh.coop_pair = (unsigned)p; sh.coop_slot = slot_in; sh.coop_arrived = 0u; sh.coop_load = 0;
                sh.coop_seq = (sh.coop_seq + 1u) & 0x7fffffu;
                fence_cluster();
                *(volatile unsigned*)&sh.coop_word = kCoopOpen | (sh.coop_seq << 8);
            }
        }
        if (CL > 1 && sh.help_rank >= 0) {
            // ---- helping a cluster mate: a share of every bulk sweep of its pair, nothing else (coop_help)
            coop_help<DIM, NT, SH>(L, sh, cg::this_cluster().block_rank(), slab_vox, a.stats);
        } else
        for (int it = iters; it < a.max_iter; ++it) {
            if (!a.resume && a.phase_cap > 0 && it == a.phase_cap) {
                // not converged within the bulk budget: park the state, a dedicated launch finishes it
                if (tid == 0) sh.bcast_i[3] = (int)atomicAdd(a.cont_count, 1u);
                __syncthreads();
                const int slot = sh.bcast_i[3];
                const size_t so = (size_t)slot * a.cap_s;
                double* cc = a.cont_cur + so * DIM;
                for (int i = tid; i < n_s; i += NT) {
                    cc[i] = L.cx[i]; cc[(size_t)a.cap_s + i] = L.cy[i];
                    if (DIM == 3) cc[2 * (size_t)a.cap_s + i] = L.cz[i];
                    a.cont_match[so + i] = L.match[i];
                    a.cont_d2lb[so + i] = L.d2lb[i];
                    for (int k = 0; k < DIM; ++k) a.cont_moved[so * DIM + (size_t)k * a.cap_s + i] = L.p0[k * L.cap + i];
                }
                if (tid == 0) {
                    double* sc = a.cont_scalar + (size_t)slot * 16;
                    for (int k = 0; k < DIM * DIM; ++k) sc[k] = sh.r_tot[k];
                    for (int k = 0; k < DIM; ++k) sc[9 + k] = sh.t_tot[k];
                    sc[12] = prev; sc[13] = err; sc[14] = (double)iters;
                    a.cont_list[slot] = p;
                    // cost class from the points swept in the last four iterations: a pair that is still re-deciding most of
                    // its points after eight iterations will go on doing so (early iterations sweep a lot in every pair)
                    const unsigned long long per_it = st_recent / 4ull;
                    // (absolute counts: the sweeps' cost goes with the points to decide, not with their share of the cloud)
                    const int cls = per_it >= 256ull ? 0 : per_it >= 96ull ? 1 : per_it >= 24ull ? 2 : 3;      // (class 0: coop_plan)
                    a.cont_bucket[(size_t)cls * a.cont_cap + atomicAdd(&a.cont_count[1 + cls], 1u)] = slot;
                }
                handed_over = true;
                break;
            }
            const bool prof = tid == 0 && it >= 8;
            if (prof) ph_t = clock64();
            // ---- correspondences (icp.py:179).  A source that has left the target's neighbourhood altogether is decided
            // against the target's front set; otherwise carry over where the movement bound allows and sweep the rest
            bool far = false;
            if (!GRID) far = far_field_matches<DIM, NT>(L, sh, phase);
            if (far && tid == 0) ++st_far;
            for (int i = tid; i < (far ? 0 : n_s); i += NT) {
                bool keep = false;
                const float lb = L.d2lb[i];
                if (lb >= 0.f) {
                    // dist(p, match) < D2 - moved, compared on squares (no fp64 sqrt), with rounding slack;
                    // moved = net displacement since the decision, rounded up (the stamped position is
                    // fp32: 2^-24 relative per axis, covered by the 1.2e-7 * magnitude term)
                    const int mw = L.match[i], j = m_idx<GRID>(mw), alt = m_alt<GRID>(mw);
                    const double px = L.cx[i], py = L.cy[i], pz = DIM == 3 ? L.cz[i] : 0.0;
                    const float x0 = L.p0[i], y0 = L.p0[L.cap + i], z0 = DIM == 3 ? L.p0[2 * L.cap + i] : 0.f;
                    const float ux = (float)(px - L.c0) - x0, uy = (float)(py - L.c1) - y0;
                    const float uz = DIM == 3 ? (float)(pz - L.c2) - z0 : 0.f;
                    const float moved = __fmaf_ru(__fsqrt_ru(__fmaf_ru(uz, uz, __fmaf_ru(uy, uy, __fmul_ru(ux, ux)))), 1.000001f,
                                                  2.4e-7f * (fabsf(x0) + fabsf(y0) + fabsf(z0) + fabsf(ux) + fabsf(uy) + fabsf(uz)));
                    const double room = (double)lb - (double)moved - 1e-12;
                    double d2 = dist2_64<DIM, GRID>(L, px, py, pz, j);
                    if (alt >= 0) {                    // the one rival: exact comparison, lowest index on ties
                        const double da = dist2_64<DIM, GRID>(L, px, py, pz, alt);
                        if (da < d2 || (da == d2 && alt < j)) { d2 = da; L.match[i] = m_pack(alt, j); }
                    }
                    keep = room > 0.0 && d2 * (1.0 + 4e-9) < room * room;
                }
                if (!keep) L.todo[atomicAdd(&sh.bcast_i[0], 1)] = (unsigned short)i;
            }
            __syncthreads();
            if (prof) { const long long c = clock64(); ph[0] += c - ph_t; ph_t = c; }
            const int n_todo = sh.bcast_i[0];
            L.list = L.todo;
            bool use_slab = false;
            if (n_todo > 0) {
                if (GRID) {
                    grid_nn<DIM, NT>(L, n_todo, gst);
                } else {
                    // The slab sweep prunes by x alone: when the points sit metres from the target (a pair that does not
                    // align) its walks cover most of the target, one candidate block after the other, and the register-
                    // blocked tile sweep is the faster way to look at everything.  The sweep measures itself (slab_off).
                    use_slab = slab_vox > 0.f && sh.slab_off == 0;
                    if (use_slab && ((n_todo + 31) >> 5) >= kSlabMinChunks)              // the slab sweep will run
                        order_todo_by_x<DIM, NT>(L, sh, n_todo, tgt_xlo, tgt_xhi);
                }
            }
            if (!GRID) {
                int my_n = n_todo;
                if (CL > 1) {
                    // helpers registered with this pair take a stretch each of a sweep of more than a few chunks (a stretch
                    // of 1-2 chunks is swept by all 16 warps of its CTA together, nn_split)
                    if (tid == 0) {
                        *(volatile int*)&sh.coop_load = n_todo;
                        const unsigned mask = *(volatile unsigned*)&sh.coop_word & 0xffu;
                        const bool share = mask != 0u && n_todo >= kCoopMinShare;
                        sh.coop_parts = share ? __popc(mask) + 1 : 1;
                        if (share) {
                            const int flags = (use_slab ? 2 : 0) | (L.list == L.todo2 ? 4 : 0);
                            fence_cluster();                   // the list and the points (written before the last barrier)
                            *(volatile unsigned long long*)&sh.coop_pub = ((unsigned long long)++sh.coop_epoch << 32) |
                                ((unsigned long long)mask << 20) | ((unsigned long long)flags << 16) | (unsigned long long)n_todo;
                        }
                    }
                    __syncthreads();
                    const int parts = sh.coop_parts;
                    if (parts > 1) my_n = min(((n_todo + parts * 32 - 1) / (parts * 32)) * 32, n_todo);
                }
                if (n_todo > 0) nn_dispatch<DIM, NT>(L, sh, my_n, use_slab ? slab_vox : 0.f);
                __syncthreads();
                if (CL > 1 && my_n != n_todo) {
                    if (tid == 0) {
                        const unsigned want = (unsigned)sh.coop_parts - 1u;
                        ICPB_COOP_SPIN(*(volatile unsigned*)&sh.coop_arrived == want);
                        *(volatile unsigned*)&sh.coop_arrived = 0u;
                        fence_cluster();                       // the helpers' decisions are in this CTA's shared memory
                    }
                    __syncthreads();
                }
                if (n_todo > 0) {
                    // a few undecided points: one warp each scans the target; many: a lane each, tiles filtered in fp32
                    if (sh.amb_n > NT / 32) resolve_ambiguous_lockstep<DIM, NT>(L, sh);
                    else if (sh.amb_n > 0) resolve_ambiguous<DIM, NT>(L, sh);
                }
            }
            if (tid == 0) {
                const int n_chunks = (n_todo + 31) >> 5;
                if (!GRID) {
                    const bool bulk = n_chunks >= kSlabMinChunks;                    // nn_dispatch's choice: not the split sweep
                    const bool slab = use_slab && bulk;
                    const unsigned long long all = (unsigned long long)n_chunks * 32ull * (unsigned long long)L.n_tiles * 32ull;
                    st_evals += slab ? (unsigned long long)sh.slab_evals : all;
                    if (slab) { if ((unsigned long long)sh.slab_evals * 5ull > all) sh.slab_off = 4; }   // walked > 20 % of the target: the tile sweep is 3-4x cheaper per candidate
                    else if (bulk && sh.slab_off > 0) --sh.slab_off;
                    sh.slab_evals = 0u;
                }
                st_swept += n_todo; st_kept += n_s - n_todo; st_iters += 1;
                if (!a.resume && a.phase_cap > 0 && it >= a.phase_cap - 4) st_recent += n_todo;
            }
            __syncthreads();
            if (prof) { const long long c = clock64(); ph[1] += c - ph_t; ph_t = c; }
            if (tid == 0) { st_amb += sh.amb_n; sh.amb_n = 0; sh.bcast_i[0] = 0; }
            if (a.trace_match && p == 0 && it < a.trace_iters)
                for (int i = tid; i < n_s; i += NT) a.trace_match[(size_t)it * a.trace_stride + i] = m_idx<GRID>(L.match[i]);

            double rr[9], tt[3];
            if (p2l) {
                if (nrm_pending) { mbar_wait(&nrm_bar, nrm_parity); nrm_parity ^= 1u; nrm_pending = false; }
                // icp.py:88-104 on the inliers
                double acc[10];
#pragma unroll
                for (int k = 0; k < 10; ++k) acc[k] = 0.0;
                for (int i0 = tid; i0 < n_s; i0 += 4 * NT) {
                    int jj[4];
                    double2 nn[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {                     // the normals live in global memory (L2)
                        const int i = i0 + u * NT;
                        jj[u] = i < n_s ? m_idx<GRID>(L.match[i]) : 0;
                        nn[u] = nrm_tma ? L.nrm_s[jj[u]] : __ldg(reinterpret_cast<const double2*>(normals) + jj[u]);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * NT;
                        if (i >= n_s) continue;
                        const double px = L.cx[i], py = L.cy[i];
                        double qx_, qy_, qz_;
                        tgt_at<DIM, GRID>(L, jj[u], qx_, qy_, qz_);
                        const double dx = px - qx_, dy = py - qy_;
                        if (gated) {
                            const double nd = sqrt(dx * dx + dy * dy);    // KDTree distance, squared again (icp.py:184)
                            if (!(nd * nd < gate2)) continue;             // icp.py:185
                        }
                        const double nx = nn[u].x, ny = nn[u].y;
                        const double c = ny * px - nx * py;               // icp.py:97
                        const double b = -(nx * dx + ny * dy);            // icp.py:101
                        acc[0] += c * c;  acc[1] += c * nx;  acc[2] += c * ny;
                        acc[3] += nx * nx; acc[4] += nx * ny; acc[5] += ny * ny;
                        acc[6] += c * b;  acc[7] += nx * b;  acc[8] += ny * b;
                        acc[9] += 1.0;
                    }
                }
                block_reduce<10, SumOp>(acc, sh, phase);
                if (gated && acc[9] < (double)min_inl) { status = ICPB200_FEW_INLIERS; break; }
                if (prof) { const long long c = clock64(); ph[2] += c - ph_t; ph_t = c; }
                if (tid == 0) {
                    const double ata[9] = {acc[0], acc[1], acc[2], acc[1], acc[3], acc[4], acc[2], acc[4], acc[5]};
                    const double atb[3] = {acc[6], acc[7], acc[8]};
                    double x[3];
                    if (solve3_lu_rcp(ata, atb, x) == 0) {
                        double ct_, st_;
                        sincos(x[0], &st_, &ct_);                             // icp.py:110-114
                        sh.r[0] = ct_; sh.r[1] = -st_; sh.r[2] = st_; sh.r[3] = ct_;
                        sh.t[0] = x[1]; sh.t[1] = x[2];
                    } else {                                                  // icp.py:107-108
                        sh.r[0] = 1.0; sh.r[1] = 0.0; sh.r[2] = 0.0; sh.r[3] = 1.0;
                        sh.t[0] = 0.0; sh.t[1] = 0.0;
                    }
                }
            } else {
                // icp.py:197-207: centroids, cross-covariance, SVD
                double m[2 * DIM + 1];
#pragma unroll
                for (int k = 0; k < 2 * DIM + 1; ++k) m[k] = 0.0;
                for (int i = tid; i < n_s; i += NT) {
                    const double px = L.cx[i], py = L.cy[i], pz = DIM == 3 ? L.cz[i] : 0.0;
                    double qx, qy, qz;
                    tgt_at<DIM, GRID>(L, m_idx<GRID>(L.match[i]), qx, qy, qz);
                    if (gated) {
                        const double dx = px - qx, dy = py - qy, dz = pz - qz;
                        const double nd = sqrt(dx * dx + dy * dy + dz * dz);
                        if (!(nd * nd < gate2)) continue;
                    }
                    m[0] += px; m[1] += py; m[DIM] += qx; m[DIM + 1] += qy;
                    if (DIM == 3) { m[2] += pz; m[DIM + 2] += qz; }
                    m[2 * DIM] += 1.0;
                }
                block_reduce<2 * DIM + 1, SumOp>(m, sh, phase);
                const double n_inl = m[2 * DIM];
                if (gated && n_inl < (double)min_inl) { status = ICPB200_FEW_INLIERS; break; }
                double mu_s[DIM], mu_t[DIM];
#pragma unroll
                for (int k = 0; k < DIM; ++k) { mu_s[k] = m[k] / n_inl; mu_t[k] = m[DIM + k] / n_inl; }
                double w[DIM * DIM];
#pragma unroll
                for (int k = 0; k < DIM * DIM; ++k) w[k] = 0.0;
                for (int i = tid; i < n_s; i += NT) {
                    double ps[3] = {L.cx[i], L.cy[i], DIM == 3 ? L.cz[i] : 0.0};
                    double qs[3];
                    tgt_at<DIM, GRID>(L, m_idx<GRID>(L.match[i]), qs[0], qs[1], qs[2]);
                    if (gated) {
                        const double dx = ps[0] - qs[0], dy = ps[1] - qs[1], dz = ps[2] - qs[2];
                        const double nd = sqrt(dx * dx + dy * dy + dz * dz);
                        if (!(nd * nd < gate2)) continue;
                    }
#pragma unroll
                    for (int u = 0; u < DIM; ++u)
#pragma unroll
                        for (int v = 0; v < DIM; ++v) w[u * DIM + v] += (ps[u] - mu_s[u]) * (qs[v] - mu_t[v]);
                }
                block_reduce<DIM * DIM, SumOp>(w, sh, phase);
                if (tid == 0) {
                    if (DIM == 2) kabsch2(w, sh.r); else kabsch3(w, sh.r, sh.kab_v);
                    for (int u = 0; u < DIM; ++u) {                          // t = mu_t - r mu_s
                        double s = mu_t[u];
                        for (int v = 0; v < DIM; ++v) s -= sh.r[u * DIM + v] * mu_s[v];
                        sh.t[u] = s;
                    }
                }
            }
            if (prof) { const long long c = clock64(); ph[3] += c - ph_t; ph_t = c; }
            __syncthreads();
            if (tid == NT - 1) {
                // icp.py:210-211: r_total = r r_total ; t_total = r t_total + t -- off the critical path: nothing in the
                // next iteration needs the totals, so the last thread (it has the fewest points to move) updates them
                // while the others already apply the step
                double nr[9], nt[3];
                for (int u = 0; u < DIM; ++u) {
                    for (int v = 0; v < DIM; ++v) {
                        double s = 0.0;
                        for (int k = 0; k < DIM; ++k) s += sh.r[u * DIM + k] * sh.r_tot[k * DIM + v];
                        nr[u * DIM + v] = s;
                    }
                    double s = 0.0;
                    for (int k = 0; k < DIM; ++k) s += sh.r[u * DIM + k] * sh.t_tot[k];
                    nt[u] = s + sh.t[u];
                }
                for (int k = 0; k < DIM * DIM; ++k) sh.r_tot[k] = nr[k];
                for (int k = 0; k < DIM; ++k) sh.t_tot[k] = nt[k];
            }
            for (int k = 0; k < DIM * DIM; ++k) rr[k] = sh.r[k];
            for (int k = 0; k < DIM; ++k) tt[k] = sh.t[k];
            // icp.py:212 apply to ALL points; icp.py:215 error vs the OLD matches
            double e[1] = {0.0};
            for (int i = tid; i < n_s; i += NT) {
                double qx, qy, qz;
                tgt_at<DIM, GRID>(L, m_idx<GRID>(L.match[i]), qx, qy, qz);
                const double x = L.cx[i], y = L.cy[i];
                double nx_, ny_, nz_ = 0.0, mv;
                if (DIM == 2) {
                    nx_ = x * rr[0] + y * rr[1] + tt[0];
                    ny_ = x * rr[2] + y * rr[3] + tt[1];
                    mv = (nx_ - x) * (nx_ - x) + (ny_ - y) * (ny_ - y);
                } else {
                    const double z = L.cz[i];
                    nx_ = x * rr[0] + y * rr[1] + z * rr[2] + tt[0];
                    ny_ = x * rr[3] + y * rr[4] + z * rr[5] + tt[1];
                    nz_ = x * rr[6] + y * rr[7] + z * rr[8] + tt[2];
                    L.cz[i] = nz_;
                    mv = (nx_ - x) * (nx_ - x) + (ny_ - y) * (ny_ - y) + (nz_ - z) * (nz_ - z);
                }
                L.cx[i] = nx_; L.cy[i] = ny_;
                const double dx = qx - nx_, dy = qy - ny_;
                double d = dx * dx + dy * dy;
                if (DIM == 3) { const double dz = qz - nz_; d += dz * dz; }
                e[0] += d;
            }
            block_reduce<1, SumOp>(e, sh, phase);
            if (prof) { const long long c = clock64(); ph[4] += c - ph_t; ph_t = c; ++ph_iters; }
            err = e[0] / (double)n_s;
            ++iters;
            const double delta = fabs(prev - err);                   // icp.py:216
            if (delta < a.err_thr) { status = ICPB200_CONVERGED; break; }   // icp.py:217-219
            prev = err;
        }
        if (CL > 1 && tid == 0 && sh.help_rank < 0) {
            // close the pair: nobody registers any more, and whoever did is told to leave and has left before this CTA
            // touches its shared memory for the next pair
            const unsigned mask = atomicAnd(&sh.coop_word, ~kCoopOpen) & 0xffu;
            if (mask) {
                fence_cluster();
                *(volatile unsigned long long*)&sh.coop_pub = ((unsigned long long)++sh.coop_epoch << 32) | ((unsigned long long)mask << 20) | (1ull << 16);
                const unsigned want = (unsigned)__popc(mask);
                ICPB_COOP_SPIN(*(volatile unsigned*)&sh.coop_arrived == want);
                *(volatile unsigned*)&sh.coop_arrived = 0u;
            }
        }
        if (nrm_pending) { mbar_wait(&nrm_bar, nrm_parity); nrm_parity ^= 1u; }      // never used (no iteration ran): still drain it
        __syncthreads();
        if (tid == 0 && a.pair_prof && !(CL > 1 && sh.help_rank >= 0)) {
            st_amb += sh.amb_n;                    // (a break may leave the last iteration's count uncollected; profiling only)
            unsigned long long* pp = a.pair_prof + 8 * (size_t)p;
            atomicAdd(&pp[0], (unsigned long long)(clock64() - pair_t0));
            atomicAdd(&pp[1], st_swept - pp_swept0);
            atomicAdd(&pp[2], st_amb - pp_amb0);
            atomicAdd(&pp[3], st_iters - pp_it0);
            // cycles by phase (iterations >= 8): classify, nearest neighbours, the rest; what the previous pairs of this
            // CTA added is remembered in shared memory
            const long long ph_rest = ph[2] + ph[3] + ph[4];
            atomicAdd(&pp[4], (unsigned long long)(ph[0] - sh.ph_seen[0]));
            atomicAdd(&pp[5], (unsigned long long)(ph[1] - sh.ph_seen[1]));
            atomicAdd(&pp[6], (unsigned long long)(ph_rest - sh.ph_seen[2]));
            sh.ph_seen[0] = ph[0]; sh.ph_seen[1] = ph[1]; sh.ph_seen[2] = ph_rest;
            st_amb -= sh.amb_n;
        }
        if (tid == 0 && !handed_over && !(CL > 1 && sh.help_rank >= 0)) {
            for (int k = 0; k < DIM * DIM; ++k) a.R_out[(size_t)p * DIM * DIM + k] = sh.r_tot[k];
            for (int k = 0; k < DIM; ++k) a.t_out[(size_t)p * DIM + k] = sh.t_tot[k];
            a.err_out[p] = err;
            if (a.prev_out) a.prev_out[p] = prev;
            a.iters_out[p] = iters;
            a.status_out[p] = status;
        }
    }
    if (CL > 1) cg::this_cluster().sync();      // no CTA of a cluster may leave while a peer can still touch its shared memory
    if (GRID && a.stats) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            unsigned long long v = gst[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((tid & 31) == 0 && v) atomicAdd(&a.stats[16 + k], v);
        }
    }
    if (tid == 0 && a.stats) {
        atomicAdd(&a.stats[0], st_evals);
        atomicAdd(&a.stats[1], st_amb);
        atomicAdd(&a.stats[2], st_iters);
        atomicAdd(&a.stats[3], st_swept);
        atomicAdd(&a.stats[4], st_kept);
        atomicAdd(&a.stats[14], st_far);
        for (int k = 0; k < 5; ++k) atomicAdd(&a.stats[8 + k], (unsigned long long)ph[k]);
        atomicAdd(&a.stats[13], ph_iters);
    }
}

// ---- K8: rotation-search scoring (features.py:165-242, slam.py:111-183) ---------------------------
// score(angle) = mean over the source of the squared distance to the nearest target after
// p' = R(angle) p + shift.  One CTA per (problem, angle): the target sits in shared memory (fp64
// tile-padded SoA + fp32 recentred, exactly as K3 stages it), the source is rotated on the fly in
// fp64, the same fp32 tile sweep finds the winning tile and the same fp64 re-evaluation + bound
// (or a full fp64 scan when the bound cannot exclude another tile) makes the distance exact.
// With want_nn the per-point nearest index and distance are written instead (one angle).
struct RotProblem {
    const double* tx; const double* ty;            // shared-memory target
    const float4* t32;
    int n_t, n_tiles;
    double c0, c1;
    float ta;
};

__device__ __forceinline__ void rot_exact(const RotProblem& P, double px, double py, float sxv, float syv, float b2v, int btv,
                                          double& d2_out, int& j_out) {
    const int j0 = btv * 32, j1 = min(j0 + 32, P.n_t);
    double best = INFINITY;
    int bj = j0;
    for (int j = j0; j < j1; ++j) {
        const int jp = pad_index(j);
        const double dx = px - P.tx[jp], dy = py - P.ty[jp];
        const double d = dx * dx + dy * dy;
        if (d < best) { best = d; bj = j; }
    }
    const float mag = fabsf(sxv) + fabsf(syv) + P.ta;
    const double other = sqrt((double)b2v) * (1.0 - 1.0e-6) - 1.8e-7 * (double)mag;
    if (!(other > sqrt(best))) {                   // another tile may hold a closer point: exact scan
        best = INFINITY; bj = 0;
        for (int j = 0; j < P.n_t; ++j) {
            const int jp = pad_index(j);
            const double dx = px - P.tx[jp], dy = py - P.ty[jp];
            const double d = dx * dx + dy * dy;
            if (d < best) { best = d; bj = j; }
        }
    }
    d2_out = best; j_out = bj;
}

template <int S>
__device__ __forceinline__ double rot_round(const RotProblem& P, const double* __restrict__ src, int n_s, int first_chunk,
                                            double ca, double sa, double ox, double oy, double* nn_d, int* nn_idx,
                                            double* part_d2, int* part_j, int j_base, bool first_slice, bool last_slice) {
    const int lane = threadIdx.x & 31;
    float sx[S], sy[S], b1[S], b2[S];
    double px[S], py[S];
    int bt[S], pt[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int i = (first_chunk + s * kNW) * 32 + lane;
        pt[s] = i < n_s ? i : -1;
        const double x = pt[s] >= 0 ? src[2 * (size_t)i] : 0.0, y = pt[s] >= 0 ? src[2 * (size_t)i + 1] : 0.0;
        px[s] = x * ca + y * (-sa) + ox;           // src @ R.T + shift (features.py:208-209)
        py[s] = x * sa + y * ca + oy;
        sx[s] = (float)(px[s] - P.c0);
        sy[s] = (float)(py[s] - P.c1);
    }
    sweep2d<S>(P.t32, 0, P.n_tiles, sx, sy, b1, b2, bt);
    double acc = 0.0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        if (pt[s] < 0) continue;
        double d2;
        int j;
        rot_exact(P, px[s], py[s], sx[s], sy[s], b2[s], bt[s], d2, j);
        j += j_base;
        if (part_d2) {                             // the target comes in slices: keep the running minimum (lowest index on ties)
            if (!first_slice) {
                const double pd = part_d2[pt[s]];
                const int pj = part_j[pt[s]];
                if (pd < d2 || (pd == d2 && pj < j)) { d2 = pd; j = pj; }
            }
            if (!last_slice) { part_d2[pt[s]] = d2; part_j[pt[s]] = j; continue; }
        }
        const double d = sqrt(d2);                 // KDTree returns the distance; the callers square it again
        acc += d * d;
        if (nn_d) { nn_d[pt[s]] = d; nn_idx[pt[s]] = j; }
    }
    return acc;
}

__global__ void __launch_bounds__(kNT) rot_scores_kernel(const RotArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    CtaShared& sh = *reinterpret_cast<CtaShared*>(smem);
    const int p = blockIdx.y, k0 = blockIdx.x * a.angles_per_cta, tid = threadIdx.x, w = tid >> 5;
    const long long ab = a.ang_off[p];
    const int n_ang = (int)(a.ang_off[p + 1] - ab);
    if (k0 >= n_ang) return;
    const double* src = a.src + 2 * a.src_off[p];
    const int n_s = (int)(a.src_off[p + 1] - a.src_off[p]);
    const int n_t_all = (int)(a.tgt_off[p + 1] - a.tgt_off[p]);
    // targets beyond shared memory come in slices of slice_len points, one launch per slice: the running minima live in
    // a.part_d2 / a.part_j, the last slice turns them into scores
    const int j_base = a.slice_len > 0 ? a.slice * a.slice_len : 0;
    const int n_t = a.slice_len > 0 ? max(0, min(a.slice_len, n_t_all - j_base)) : n_t_all;
    const bool first_slice = a.slice_len <= 0 || a.slice == 0;
    const bool last_slice = a.slice_len <= 0 || a.slice == a.n_slices - 1;
    const double* tgt = a.tgt + 2 * (a.tgt_off[p] + j_base);
    const int k1 = min(k0 + a.angles_per_cta, n_ang);
    if (n_s <= 0 || n_t_all <= 0) {
        for (int k = k0 + tid; k < k1; k += kNT) a.scores[ab + k] = INFINITY;
        return;
    }
    if (n_t <= 0 && !last_slice) return;           // this problem's target ended in an earlier slice
    const int tstride = a.cap_t + a.cap_t / 32;
    double* tx = reinterpret_cast<double*>(smem + align16(sizeof(CtaShared)));
    double* ty = tx + tstride;
    float* f = reinterpret_cast<float*>(smem + align16(sizeof(CtaShared)) + align16(sizeof(double) * 2 * (size_t)tstride));
    int phase = 0;
    // bounding box -> recentring offset
    double b[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
    for (int j = tid; j < n_t; j += kNT) {
        const double x = tgt[2 * (size_t)j], y = tgt[2 * (size_t)j + 1];
        b[0] = fmin(b[0], x); b[1] = fmin(b[1], y); b[2] = fmin(b[2], -x); b[3] = fmin(b[3], -y);
    }
    block_reduce<4, MinOp>(b, sh, phase);
    RotProblem P;
    P.tx = tx; P.ty = ty; P.t32 = reinterpret_cast<const float4*>(f);
    P.n_t = n_t; P.n_tiles = (n_t + 31) / 32;
    P.c0 = 0.5 * (b[0] - b[2]); P.c1 = 0.5 * (b[1] - b[3]);
    P.ta = (float)((0.5 * (-b[2] - b[0]) + 0.5 * (-b[3] - b[1])) * 1.0000002);
    for (int j = tid; j < P.n_tiles * 32; j += kNT) {
        if (j < n_t) {
            const double x = tgt[2 * (size_t)j], y = tgt[2 * (size_t)j + 1];
            tx[pad_index(j)] = x; ty[pad_index(j)] = y;
            f[2 * j] = (float)(x - P.c0); f[2 * j + 1] = (float)(y - P.c1);
        } else {
            f[2 * j] = kFar; f[2 * j + 1] = kFar;
        }
    }
    __syncthreads();
    const int n_chunks = (n_s + 31) >> 5;
    // the staged target serves every angle of this CTA's group
    for (int k = k0; k < k1; ++k) {
    const double ang = a.angles[ab + k];
    double sa, ca;
    sincos(ang, &sa, &ca);                         // np.cos / np.sin of the same double
    const double ox = a.shift[2 * p], oy = a.shift[2 * p + 1];
    double* nn_d = a.nn_dist ? a.nn_dist + a.src_off[p] : nullptr;
    int* nn_i = a.nn_idx ? a.nn_idx + a.src_off[p] : nullptr;
    double acc[1] = {0.0};
    double* pd2 = a.part_d2 ? a.part_d2 + (size_t)(ab + k) * a.part_stride : nullptr;
    int* pj = a.part_j ? a.part_j + (size_t)(ab + k) * a.part_stride : nullptr;
    constexpr int kRotS = 4;                       // chunks per warp per round (register budget: two CTAs per SM)
    for (int base = 0; base < n_chunks; base += kRotS * kNW) {
        const int first = base + w;
        const int mine = first < n_chunks ? min(kRotS, (n_chunks - first + kNW - 1) / kNW) : 0;   // warp-uniform
        switch (mine) {
            case 0: break;
            case 1: acc[0] += rot_round<1>(P, src, n_s, first, ca, sa, ox, oy, nn_d, nn_i, pd2, pj, j_base, first_slice, last_slice); break;
            case 2: acc[0] += rot_round<2>(P, src, n_s, first, ca, sa, ox, oy, nn_d, nn_i, pd2, pj, j_base, first_slice, last_slice); break;
            case 3: acc[0] += rot_round<3>(P, src, n_s, first, ca, sa, ox, oy, nn_d, nn_i, pd2, pj, j_base, first_slice, last_slice); break;
            default: acc[0] += rot_round<4>(P, src, n_s, first, ca, sa, ox, oy, nn_d, nn_i, pd2, pj, j_base, first_slice, last_slice); break;
        }
    }
    block_reduce<1, SumOp>(acc, sh, phase);
    if (tid == 0 && last_slice) a.scores[ab + k] = acc[0] / (double)n_s;         // np.mean(dists ** 2)
    }
}

size_t rot_smem_bytes(int cap_t) {
    return align16(sizeof(CtaShared)) + align16(sizeof(double) * 2 * (size_t)(cap_t + cap_t / 32)) + sizeof(float) * 2 * (size_t)cap_t + 16;
}

int launch_rot_scores(const RotArgs& a_in, int n_problems, int max_angles, int sm_count, cudaStream_t stream) {
    RotArgs a = a_in;
    // one angle per CTA while that still leaves SMs idle, groups of up to 8 once the launch fills the machine
    a.angles_per_cta = 1;
    while (a.angles_per_cta < 8 && (long long)n_problems * ((max_angles + 2 * a.angles_per_cta - 1) / (2 * a.angles_per_cta)) >= 4LL * sm_count)
        a.angles_per_cta *= 2;
    max_angles = (max_angles + a.angles_per_cta - 1) / a.angles_per_cta;
    const size_t smem = rot_smem_bytes(a.cap_t);
    ICPB_CUDA(cudaFuncSetAttribute(rot_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rot_scores_kernel<<<dim3((unsigned)max_angles, (unsigned)n_problems), kNT, smem, stream>>>(a);
    ICPB_LAUNCH_CHECK();
    return ICPB200_OK;
}

// ---- host-side launchers -------------------------------------------------------------
int launch_mark_used(const IcpArgs& a, bool p2l, cudaStream_t stream) {
    mark_used_kernel<<<(a.n_pairs + 255) / 256, 256, 0, stream>>>(a.n_pairs, a.src_idx, a.tgt_idx, a.s.used, a.t.used,
                                                                  p2l ? a.t.is_tgt : nullptr, a.s.n_clouds, a.t.n_clouds);
    ICPB_LAUNCH_CHECK();
    return ICPB200_OK;
}

int launch_voxel_clouds(const CloudSet& cs, int dim, double voxel, int sort_pad, cudaStream_t stream, int first, int count) {
    if (count < 0) count = cs.n_clouds - first;
    if (count == 0) return ICPB200_OK;
    const size_t smem = icp_voxel_smem_bytes(sort_pad);
    if (dim == 2) {
        ICPB_CUDA(cudaFuncSetAttribute(voxel_clouds_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        voxel_clouds_kernel<2><<<count, kNT, smem, stream>>>(cs, voxel, sort_pad, first);
    } else {
        ICPB_CUDA(cudaFuncSetAttribute(voxel_clouds_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        voxel_clouds_kernel<3><<<count, kNT, smem, stream>>>(cs, voxel, sort_pad, first);
    }
    ICPB_LAUNCH_CHECK();
    return ICPB200_OK;
}

int launch_normals(const CloudSet& cs, int cap_t, int normal_k, double voxel, cudaStream_t stream, int first, int count) {
    if (count < 0) count = cs.n_clouds - first;
    if (count == 0) return ICPB200_OK;
    const size_t smem = icp_normals_smem_bytes(cap_t);
    if (normal_k + 1 <= kKnnReg && cap_t <= 4096) {
        // few clouds (the online loop registers one pair per call): several CTAs share a cloud's query blocks
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int split = std::max(1, std::min(8, sms / std::max(count, 1)));
        // (2 or 3 CTAs per SM with more registers and fewer spills measured slower: 368 / 351 us against 343 us on C2)
        ICPB_CUDA(cudaFuncSetAttribute(normals_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        normals_sweep_kernel<<<count * split, kNT, smem, stream>>>(cs, normal_k, cap_t, first, split);
        ICPB_LAUNCH_CHECK();
        return ICPB200_OK;
    }
    ICPB_CUDA(cudaFuncSetAttribute(normals_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    normals_kernel<<<count, kNT, smem, stream>>>(cs, normal_k, cap_t, voxel, first);
    ICPB_LAUNCH_CHECK();
    return ICPB200_OK;
}

template <int DIM, bool GRID, int MINB, int NT>
static int launch_pairs_t(const IcpArgs& a, int n_ctas, size_t smem, cudaStream_t stream) {
    ICPB_CUDA(cudaFuncSetAttribute(icp_pairs_kernel<DIM, GRID, MINB, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    icp_pairs_kernel<DIM, GRID, MINB, NT><<<n_ctas, NT, smem, stream>>>(a);
    ICPB_LAUNCH_CHECK();
    return ICPB200_OK;
}

// CTAs per SM the bulk launches are compiled for (profiles/kernel_ab.py, grid_batch_ab.py): 2-D brute 2 (128 registers,
// no spills: 1.22 against 1.26 ms on C2), grid mode 2 (4.56 against 4.75 ms for 592 scan->submap pairs), 3-D 3 (2 measured
// the same on the teapot batch)
static constexpr int bulk_minb(int dim, bool grid) { return grid ? 2 : (dim == 2 ? 2 : 3); }
// Launches of at most one CTA per SM (the hand-over launch, single calls, small batches) are latency-bound -- a
// registration is a chain of dependent iterations -- and take the 512-thread variant: two source points per thread instead
// of four, sixteen warps to hide the shared-memory and fp64 latencies, every register the SM has (128 per thread).
constexpr int kRoomyNT = 512;

int launch_icp_pairs(const IcpArgs& a, int dim, bool grid, int n_ctas, size_t smem_min, cudaStream_t stream) {
    static const int sms = [] {
        int dev = 0, n = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n;
    }();
    static const int roomy_nt = [] {                  // A/B switch (profiles/README.md): ICPB200_ROOMY_NT=256
        const char* e = getenv("ICPB200_ROOMY_NT");
        return e && atoi(e) == 256 ? 256 : kRoomyNT;
    }();
    const bool roomy = a.resume || n_ctas <= sms;
    const int nt = roomy ? roomy_nt : kNT;
    const size_t smem = std::max(smem_min, icp_pair_smem_bytes(dim, a.cap_s, a.cap_t, nt));
    if (grid) {                                                                // grid mode is 2-D only
        if (roomy && nt == 512) return launch_pairs_t<2, true, 1, kRoomyNT>(a, n_ctas, smem, stream);
        if (roomy) return launch_pairs_t<2, true, 1, kNT>(a, n_ctas, smem, stream);
        return launch_pairs_t<2, true, bulk_minb(2, true), kNT>(a, n_ctas, smem, stream);
    }
    if (dim == 2) {
        if (roomy && nt == 512) return launch_pairs_t<2, false, 1, kRoomyNT>(a, n_ctas, smem, stream);
        if (roomy) return launch_pairs_t<2, false, 1, kNT>(a, n_ctas, smem, stream);
        return launch_pairs_t<2, false, bulk_minb(2, false), kNT>(a, n_ctas, smem, stream);
    }
    if (roomy && nt == 512) return launch_pairs_t<3, false, 1, kRoomyNT>(a, n_ctas, smem, stream);
    if (roomy) return launch_pairs_t<3, false, 1, kNT>(a, n_ctas, smem, stream);
    return launch_pairs_t<3, false, bulk_minb(3, false), kNT>(a, n_ctas, smem, stream);
}

bool icp_cluster_variant(int dim, bool grid) { return dim == 2 && !grid; }

// The hand-over launch of the 2-D brute-force path: 512-thread CTAs in clusters of kCoopCluster (see icp_pairs_kernel).
// *n_ctas_out = the CTAs launched: as many whole clusters as the device can hold at once, at most max_ctas (the cluster
// shape may leave a few SMs unused -- the caller puts plain CTAs of the same queue on those).
int launch_icp_pairs_cluster(const IcpArgs& a, int max_ctas, size_t smem_min, cudaStream_t stream, int* n_ctas_out) {
    const size_t smem = std::max(smem_min, icp_pair_smem_bytes(2, a.cap_s, a.cap_t, kRoomyNT));
    auto kern = icp_pairs_kernel<2, false, 1, kRoomyNT, kCoopCluster>;
    ICPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(max_ctas / kCoopCluster * kCoopCluster));
    cfg.blockDim = dim3(kRoomyNT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCoopCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n_clusters = 0;
    ICPB_CUDA(cudaOccupancyMaxActiveClusters(&n_clusters, kern, &cfg));
    const int n_ctas = std::min(max_ctas / kCoopCluster, n_clusters) * kCoopCluster;
    *n_ctas_out = n_ctas;
    if (n_ctas <= 0) return ICPB200_OK;
    cfg.gridDim = dim3((unsigned)n_ctas);
    IcpArgs b = a;
    b.coop_ctas = n_ctas;
    ICPB_CUDA(cudaLaunchKernelEx(&cfg, kern, b));
    ICPB_LAUNCH_CHECK();
    return ICPB200_OK;
}

template <int DIM, bool GRID, int MINB>
static int max_ctas_t(size_t smem) {
    int n = 0;
    cudaFuncSetAttribute(icp_pairs_kernel<DIM, GRID, MINB, kNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, icp_pairs_kernel<DIM, GRID, MINB, kNT>, kNT, smem) != cudaSuccess || n < 1) n = 1;
    return n;
}

int icp_max_ctas_per_sm(int dim, bool grid, size_t smem) {
    if (grid) return max_ctas_t<2, true, bulk_minb(2, true)>(smem);
    if (dim == 2) return max_ctas_t<2, false, bulk_minb(2, false)>(smem);
    return max_ctas_t<3, false, bulk_minb(3, false)>(smem);
}

}  // namespace icpb
