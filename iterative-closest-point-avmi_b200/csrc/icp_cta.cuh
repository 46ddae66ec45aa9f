// Device building blocks shared by the per-pair ICP kernel and the standalone
// voxel-downsample kernel: deterministic block reductions, the CTA-local
// voxel-grid mean (utilities/icp.py:117-129) and its bitonic sort.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <type_traits>

namespace icpb {

constexpr int kNT = 256;               // threads per CTA
constexpr int kNW = kNT / 32;          // warps per CTA
constexpr int kRedMax = 12;            // widest block reduction (values)

// Shared-memory header: lives at the start of dynamic shared memory in every
// kernel that uses these helpers.
template <int NW_>
struct CtaSharedT {
    static constexpr int kWarps = NW_;
    double red[2][NW_][kRedMax];       // double-buffered warp partials
    double red_tot[2][kRedMax];        // ... and block totals
    int scan_tot[NW_ + 1];
    int bcast_i[4];
    unsigned int pair;
    // registration state (ICP kernel only)
    double r_tot[9], t_tot[3], r[9], t[3];
    double kab_v[9];                   // 3-D point-to-point: right singular basis of the last solve (warm start of the next)
    double center[3], lo_t[3], hi_t[3];
    double grid_h;
    int grid_nx, grid_ny;
    int n_s, n_t;
    int amb_n;
    unsigned int slab_evals;           // fp32 evaluations of the slab sweeps of this iteration (statistics)
    // split-sweep partials (K3): [warp][lane]
    float part_b1[NW_][32], part_b2[NW_][32], part_b3[NW_][32];
    int part_bt[NW_][32], part_bt2[NW_][32];
    float part_gap[NW_][32];           // slab sweep split by direction: the partner's x-gap to what it left unvisited
    int bins[65];                      // x-bins of the points to decide (K3: todo list ordered for the slab sweep)
    int front_n;                       // K3 far-field front set size
    int slab_off;                      // K3: sweeps left before the slab sweep is tried again (it pruned too little last time)
    long long ph_seen[3];              // per-pair profile: phase cycles already attributed to earlier pairs of this CTA
    // K3, cluster variant of the hand-over launch: a CTA that ran out of pairs helps a cluster mate that is still running
    // one.  These words are read and written by the mates through distributed shared memory.
    unsigned long long coop_pub;       // owner -> helpers, one atomic word per publication:
                                       //   epoch << 32 | helper mask << 20 | flags << 16 | points to decide
                                       //   flags: bit 0 the pair is finished, bit 1 slab sweep, bit 2 the list is todo2
    unsigned int coop_word;            // bit 31: a pair runs here and takes helpers; bits 8..30: its sequence number on this
                                       //   CTA; bits 0..7: the cluster ranks registered as helpers
    unsigned int coop_arrived;         // helpers done with the current publication
    unsigned int coop_idle;            // this CTA has no pair of its own any more
    unsigned int coop_pair;            // the pair running here and where its state was parked
    int coop_slot;
    int coop_parts;                    // owner, this iteration: CTAs sharing the sweep (1: nobody helps)
    int coop_load;                     // points the pair running here had to decide in its last iteration
    // helper side (local copies of what the owner published, broadcast to the CTA)
    int help_rank;                     // the mate this CTA is about to help (-1: none)
    unsigned int help_word;            // its coop_word when it was chosen
    int help_ok;
    int help_dry;                      // the queue gave this CTA nothing any more
    int coop_first;                    // this CTA has not taken its first pair yet (it is dealt statically)
    unsigned int coop_seq, coop_epoch; // owner: pairs opened here, publications made (kept here, not in registers)
    unsigned long long help_pub;
};
using CtaShared = CtaSharedT<kNW>;     // the 256-thread kernels (K1, K2, K8, bulk K3)

struct SumOp {
    __device__ static double f(double a, double b) { return a + b; }
    __device__ static double id() { return 0.0; }
};
struct MinOp {
    __device__ static double f(double a, double b) { return fmin(a, b); }
    __device__ static double id() { return INFINITY; }
};

// All-threads block reduction of NV doubles.  Fixed tree + fixed warp order => bitwise
// deterministic, and every thread returns the same values.
//   warp    recursive halving: at each of the first log2(NVP) shuffle distances a lane hands half of
//           its values to its partner and keeps the other half (which half is the lane's bit), so
//           NVP values cost NVP - 1 shuffles per lane instead of 5 * NVP; the remaining distances
//           are a plain butterfly.  Lane bits then spell the index of the one value a lane holds.
//   block   threads 0 .. NV-1 each add one value over the warps (in warp order), everybody reads
//           the NV totals back: NV + kNW shared loads per thread instead of NV * kNW.
// `phase` alternates the scratch buffer so a reduction may start while another is still read.
template <int NV, class Op, class SH>
__device__ __forceinline__ void block_reduce(double (&v)[NV], SH& sh, int& phase) {
    constexpr int NWB = SH::kWarps;
    static_assert(NV <= kRedMax, "reduction too wide");
    constexpr int NVP = NV <= 1 ? 1 : NV <= 2 ? 2 : NV <= 4 ? 4 : NV <= 8 ? 8 : 16;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    double x[NVP];
#pragma unroll
    for (int k = 0; k < NVP; ++k) x[k] = k < NV ? v[k] : Op::id();
    int held = 0;                                  // index of x[0] once the halving is over
    int c = NVP;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        if (c > 1) {
            const bool upper = (l & o) != 0;
            const int h = c / 2;
#pragma unroll
            for (int k = 0; k < NVP / 2; ++k) {
                if (k < h) {
                    const double send = upper ? x[k] : x[k + h];
                    const double keep = upper ? x[k + h] : x[k];
                    x[k] = Op::f(keep, __shfl_xor_sync(0xffffffffu, send, o));
                }
            }
            held += upper ? h : 0;
            c = h;
        } else {
            x[0] = Op::f(x[0], __shfl_xor_sync(0xffffffffu, x[0], o));
        }
    }
    double(*buf)[kRedMax] = sh.red[phase & 1];
    constexpr int kDup = 32 / NVP;                 // lanes that hold the same value after the butterfly tail
    if ((l & (kDup - 1)) == 0 && held < NV) buf[w][held] = x[0];
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = buf[0][threadIdx.x];
#pragma unroll
        for (int ww = 1; ww < NWB; ++ww) s = Op::f(s, buf[ww][threadIdx.x]);
        sh.red_tot[phase & 1][threadIdx.x] = s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = sh.red_tot[phase & 1][k];
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

// Register/shuffle bitonic sort for n_pad = kNT * E elements in a blocked layout (thread t holds
// elements t*E .. t*E+E-1): compare-exchange distances below E stay in registers, distances
// below 32*E go through warp shuffles, only the larger ones through shared memory (6 of the 66
// passes for 2048 elements).  Same order as bitonic_sort_pairs: ascending (key, idx).
template <int E>
__device__ __forceinline__ void bitonic_sort_blocked(unsigned long long (&key)[E], unsigned int (&idx)[E],
                                                     unsigned long long* skeys, unsigned int* sidx) {
    const int tid = threadIdx.x, lane = tid & 31;
    constexpr int n_pad = kNT * E;
    auto greater = [](unsigned long long ka, unsigned ia, unsigned long long kb, unsigned ib) {
        return ka > kb || (ka == kb && ia > ib);
    };
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j >= E; j >>= 1) {
            if (false) {
            } else if (j < 32 * E) {                       // partner lane in this warp
                const int lj = j / E;
                const bool lower = (lane & lj) == 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key[e], lj);
                    const unsigned oi = __shfl_xor_sync(0xffffffffu, idx[e], lj);
                    const bool asc = (((tid * E + e) & k) == 0);
                    // the lower element of the pair keeps the minimum when ascending
                    const bool mine_greater = greater(key[e], idx[e], ok, oi);
                    const bool take = (lower == asc) ? mine_greater : !mine_greater;
                    if (take) { key[e] = ok; idx[e] = oi; }
                }
            } else {                                       // partner in another warp: through shared memory
                __syncthreads();
#pragma unroll
                for (int e = 0; e < E; ++e) { skeys[tid * E + e] = key[e]; sidx[tid * E + e] = idx[e]; }
                __syncthreads();
                const int pt = tid ^ (j / E);
                const bool lower = (tid & (j / E)) == 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const unsigned long long ok = skeys[pt * E + e];
                    const unsigned oi = sidx[pt * E + e];
                    const bool asc = (((tid * E + e) & k) == 0);
                    const bool mine_greater = greater(key[e], idx[e], ok, oi);
                    const bool take = (lower == asc) ? mine_greater : !mine_greater;
                    if (take) { key[e] = ok; idx[e] = oi; }
                }
            }
        }
        // distances below E: partner in this thread's registers (all indices compile-time)
#pragma unroll
        for (int jj = E / 2; jj > 0; jj >>= 1) {
            if (jj >= k) continue;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int pe = e ^ jj;
                if (pe > e) {
                    const bool asc = (((tid * E + e) & k) == 0);
                    if (greater(key[e], idx[e], key[pe], idx[pe]) == asc) {
                        const unsigned long long tk = key[e]; key[e] = key[pe]; key[pe] = tk;
                        const unsigned ti = idx[e]; idx[e] = idx[pe]; idx[pe] = ti;
                    }
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < E; ++e) { skeys[tid * E + e] = key[e]; sidx[tid * E + e] = idx[e]; }
    __syncthreads();
}

// The same network on single 32-bit words (cell key and input index packed into one word, see
// cta_voxel_means): a compare-exchange is one min and one max.
template <int E>
__device__ __forceinline__ void bitonic_sort_blocked_u32(unsigned int (&key)[E], unsigned int* skeys) {
    const int tid = threadIdx.x, lane = tid & 31;
    constexpr int n_pad = kNT * E;
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j >= E; j >>= 1) {
            const int tj = j / E;
            const bool keep_min = (((tid & tj) == 0) == (((tid * E) & k) == 0));    // k >= 2E here: the direction is per thread
            if (j < 32 * E) {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const unsigned o = __shfl_xor_sync(0xffffffffu, key[e], tj);
                    key[e] = keep_min ? min(key[e], o) : max(key[e], o);
                }
            } else {
                __syncthreads();
#pragma unroll
                for (int e = 0; e < E; ++e) skeys[e * kNT + tid] = key[e];             // [e][tid]: conflict-free
                __syncthreads();
                const int pt = tid ^ tj;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const unsigned o = skeys[e * kNT + pt];
                    key[e] = keep_min ? min(key[e], o) : max(key[e], o);
                }
            }
        }
#pragma unroll
        for (int jj = E / 2; jj > 0; jj >>= 1) {
            if (jj >= k) continue;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int pe = e ^ jj;
                if (pe > e) {
                    const bool asc = (((tid * E + e) & k) == 0);
                    const unsigned lo_ = min(key[e], key[pe]), hi_ = max(key[e], key[pe]);
                    key[e] = asc ? lo_ : hi_;
                    key[pe] = asc ? hi_ : lo_;
                }
            }
        }
    }
    __syncthreads();
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

    auto make_key = [&](int i, unsigned long long& key, unsigned int& id) {
        key = ~0ull; id = ~0u;
        if (i < n) {
            key = 0ull;
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
                const double c = floor((raw[(size_t)i * DIM + a] - lo[a]) / voxel);
                key = key * ext[a] + (unsigned long long)c;
            }
            id = (unsigned int)i;
        }
    };
    auto sort_blocked = [&](auto e_tag) {
        constexpr int E = decltype(e_tag)::value;
        unsigned long long k[E];
        unsigned int id[E];
#pragma unroll
        for (int e = 0; e < E; ++e) make_key(threadIdx.x * E + e, k[e], id[e]);
        bitonic_sort_blocked<E>(k, id, keys, idx);
    };
    // Cell key and input index fit one 32-bit word together (always for lidar scans: ~600k cells x 2048
    // points): sort plain words.  Padding sorts last (all ones) in both forms.
    int idx_bits = 0;
    while ((1 << idx_bits) < n_pad) ++idx_bits;
    const bool pack32 = span * (double)n_pad < 4.0e9 && n_pad <= 8 * kNT;
    auto sort_packed = [&](auto e_tag) {
        constexpr int E = decltype(e_tag)::value;
        unsigned int k[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            unsigned long long key;
            unsigned int id;
            make_key(threadIdx.x * E + e, key, id);
            k[e] = id == ~0u ? ~0u : (unsigned int)((key << idx_bits) | id);
        }
        bitonic_sort_blocked_u32<E>(k, idx);               // idx[] doubles as the exchange buffer
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const bool pad = k[e] == ~0u;
            keys[threadIdx.x * E + e] = pad ? ~0ull : (unsigned long long)(k[e] >> idx_bits);
        }
        __syncthreads();                                   // idx[] is free again
#pragma unroll
        for (int e = 0; e < E; ++e) idx[threadIdx.x * E + e] = k[e] == ~0u ? ~0u : (k[e] & ((1u << idx_bits) - 1u));
        __syncthreads();
    };
    const int epl = n_pad / kNT;
    if (pack32 && epl == 1) sort_packed(std::integral_constant<int, 1>{});
    else if (pack32 && epl == 2) sort_packed(std::integral_constant<int, 2>{});
    else if (pack32 && epl == 4) sort_packed(std::integral_constant<int, 4>{});
    else if (pack32 && epl == 8) sort_packed(std::integral_constant<int, 8>{});
    else switch (epl) {
        case 1: sort_blocked(std::integral_constant<int, 1>{}); break;
        case 2: sort_blocked(std::integral_constant<int, 2>{}); break;
        case 4: sort_blocked(std::integral_constant<int, 4>{}); break;
        case 8: sort_blocked(std::integral_constant<int, 8>{}); break;
        default:
            for (int i = threadIdx.x; i < n_pad; i += kNT) make_key(i, keys[i], idx[i]);
            __syncthreads();
            bitonic_sort_pairs(keys, idx, n_pad);
    }

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
