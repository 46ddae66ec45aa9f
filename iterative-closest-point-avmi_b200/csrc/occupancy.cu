// Occupancy-grid log-odds raycast for sm_100a -- replaces
// /root/reference/utilities/mapping.py:103-141 (and the _rebuild_map replay,
// /root/reference/slam.py:271-277) for a batch of scans applied in order.
//
// Why it is not "one thread per ray with atomics on the grid": the update is
// x = f32(f64(x) + c) per event with one clamp per scan, so the value of a
// cell depends on the ORDER of scans touching it.  Within
// one scan only the counts matter: m hits then k misses then clamp.  So:
//
//   1. ray_setup      world -> cell for every endpoint / origin (fp64, exact)
//   2. bin_count      every ray is cut into runs, one per 64x64-cell tile it
//                     crosses (closed-form Bresenham, bres.cuh); count runs
//                     per (tile, scan)
//   3. exclusive scan over the (tile, scan) count matrix
//   4. bin_fill       write the runs, grouped by tile and, inside a tile, by scan
//   5. tile_apply     one CTA per tile keeps the tile in shared memory and
//                     replays ITS scans in order: count k/m per cell with
//                     shared-memory atomics, then run the fp64->fp32 add
//                     chain and the clamp on the touched cells only.
//
// Tiles are independent, so there is no grid-wide synchronisation per scan;
// HBM sees each active tile once in and once out per batch, the runs once,
// the endpoints twice.  Multi-GPU: a rank owns the tiles t with
// occ_owner(tile) == rank and skips all others in step 2.
#include "icp_b200.h"
#include "bres.cuh"
#include "common.cuh"
#include "occupancy.h"

#include <algorithm>

namespace icpb {

constexpr int TS = kOccTile;                 // tile edge in cells
constexpr int TCELLS = TS * TS;
constexpr int kOccNT = 512;
constexpr unsigned kHitUnit = 1u << 20;      // counter word: hits << 20 | misses
constexpr unsigned kMissMask = kHitUnit - 1u;

constexpr size_t kOccSmem = sizeof(float) * TCELLS * (1 + 8 /* kOccBatch */) + 3 * sizeof(unsigned) * 256 /* kOccList */;

// A run is self-contained: the tile kernel needs neither the ray's endpoint nor
// the scan's origin to walk it (no dependent loads on its critical path).
struct __align__(16) Run {                    // 32 bytes
    int n0;                                   // first step of the run
    int j0;                                   // minor steps at n0
    int len;                                  // cells in run; 0 marks a hit
    int idx0;                                 // local cell index at n0 (hit: the hit cell) | flags << 16
    int dmaj, dmin;                           // ray slope
    int pad0, pad1;
};
constexpr int kRunXMajor = 1 << 16, kRunMajPos = 1 << 17, kRunMinPos = 1 << 18;

// ---- 1. setup ----------------------------------------------------------------
__global__ void occ_scan_setup(const double* __restrict__ origins, int n_scans, double min_x,
                               double min_y, double res, int2* __restrict__ origin_cell) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_scans) return;
    // mapping.py:57-60: floor((w - min) / resolution)
    origin_cell[s] = make_int2(sat_cell(floor((origins[2 * s] - min_x) / res)),
                               sat_cell(floor((origins[2 * s + 1] - min_y) / res)));
}

__global__ void occ_ray_setup(const double2* __restrict__ hits, const long long* __restrict__ hit_off,
                              int n_scans, long long n_rays, double min_x, double min_y, double res,
                              int2* __restrict__ ray_cell, int* __restrict__ ray_scan) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const double2 h = hits[r];
    // mapping.py:94-98
    ray_cell[r] = make_int2(sat_cell(floor((h.x - min_x) / res)), sat_cell(floor((h.y - min_y) / res)));
    int lo = 0, hi = n_scans;                 // largest s with hit_off[s] <= r
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (hit_off[mid] <= r) lo = mid; else hi = mid;
    }
    ray_scan[r] = lo;
}

// ---- 2./4. binning -------------------------------------------------------------
struct BinArgs {
    const int2* ray_cell;
    const int* ray_scan;
    const int2* origin_cell;
    long long ray_begin, ray_end;             // rays of this scan chunk
    int scan_begin, chunk_scans;              // counts are indexed [tile][scan - scan_begin]
    int nx, ny, tiles_x;
    int rank, world;
    unsigned int* counts;                     // count pass: += 1 ; fill pass: cursor
    const unsigned int* offsets;              // fill pass only
    Run* runs;                                // fill pass only
    unsigned long long* stats;                // [rays, traversed, hits, runs]
};

// All lanes of a warp walk their rays' tile runs in lock step.  Runs of the
// warp that fall into the same (tile, scan) group are allocated as one block
// of consecutive slots in lane order, so rays that arrive sorted by angle (a
// lidar scan) stay sorted inside the group: the tile kernel relies on that
// for its neighbour-lane de-duplication (performance only, never correctness).
template <bool FILL>
__global__ void __launch_bounds__(256) occ_bin(const BinArgs a) {
    const long long r = a.ray_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned long long cells = 0, hits = 0, nruns = 0;
    TileRunIter<TS> it;
    it.init_empty();
    int sl = 0;
    bool hit_pending = false;
    int hit_tile = 0, hit_cell = 0;
    if (r < a.ray_end) {
        const int s = a.ray_scan[r];
        sl = s - a.scan_begin;
        const int2 o = a.origin_cell[s], h = a.ray_cell[r];
        it.init(make_ray(o.x, o.y, h.x, h.y), a.nx, a.ny, a.tiles_x);
        if (h.x >= 0 && h.x < a.nx && h.y >= 0 && h.y < a.ny) {       // mapping.py:124-127
            hit_pending = true;
            hit_tile = (h.y / TS) * a.tiles_x + (h.x / TS);
            hit_cell = (h.y % TS) * TS + (h.x % TS);
        }
    }
    for (;;) {
        TileRun t;
        bool has = it.next(t);
        if (!has && hit_pending) {                                    // the hit goes last
            hit_pending = false;
            has = true;
            t.tile = hit_tile; t.n0 = hit_cell; t.j0 = 0; t.len = 0;
        }
        if (!__any_sync(0xffffffffu, has)) break;
        const bool owned = has && occ_owner((t.tile % a.tiles_x) * TS, (t.tile / a.tiles_x) * TS, a.nx, a.ny, a.world) == a.rank;
        const unsigned long long gi = owned ? (unsigned long long)t.tile * a.chunk_scans + sl : ~0ull;
        const unsigned peers = __match_any_sync(0xffffffffu, gi);
        const int leader = __ffs(peers) - 1;
        unsigned base = 0;
        if (owned && lane == leader) base = atomicAdd(&a.counts[gi], (unsigned)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (owned) {
            if (FILL) {
                Run run;
                run.n0 = t.n0; run.j0 = t.j0; run.len = t.len;
                run.dmaj = it.g.dmaj; run.dmin = it.g.dmin; run.pad0 = 0; run.pad1 = 0;
                if (t.len == 0) {
                    run.idx0 = t.n0;                                  // hit: n0 carried the local cell
                    run.n0 = 0;
                } else {
                    int x, y;
                    cell_at(it.g, t.n0, t.j0, x, y);
                    run.idx0 = ((y % TS) * TS + (x % TS)) | (it.g.xmajor ? kRunXMajor : 0) |
                               (it.g.smaj > 0 ? kRunMajPos : 0) | (it.g.smin > 0 ? kRunMinPos : 0);
                }
                int4* dst = reinterpret_cast<int4*>(a.runs + (a.offsets[gi] + base + __popc(peers & lt_mask)));
                dst[0] = reinterpret_cast<const int4*>(&run)[0];
                dst[1] = reinterpret_cast<const int4*>(&run)[1];
            } else {
                cells += t.len;
                hits += t.len == 0;
                ++nruns;
            }
        }
    }
    if (!FILL) {
        // block totals -> three global atomics per block
        __shared__ unsigned long long part[3][8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cells += __shfl_xor_sync(0xffffffffu, cells, o);
            hits += __shfl_xor_sync(0xffffffffu, hits, o);
            nruns += __shfl_xor_sync(0xffffffffu, nruns, o);
        }
        const int w = threadIdx.x >> 5;
        if (lane == 0) { part[0][w] = cells; part[1][w] = hits; part[2][w] = nruns; }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long c = 0, h = 0, n = 0;
            for (int k = 0; k < 8; ++k) { c += part[0][k]; h += part[1][k]; n += part[2][k]; }
            if (c) atomicAdd(&a.stats[1], c);
            if (h) atomicAdd(&a.stats[2], h);
            if (n) atomicAdd(&a.stats[3], n);
        }
    }
}

// ---- 3. exclusive scan of a u32 array (n up to 2^31), three passes ---------------
constexpr int kScanBlock = 1024;              // elements per block (256 threads x 4)

__device__ __forceinline__ unsigned block_scan_u32(unsigned v, unsigned* warp_tot, unsigned& total) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (l >= o) inc += t;
    }
    __syncthreads();
    if (l == 31) warp_tot[w] = inc;
    __syncthreads();
    unsigned base = 0, tot = 0;
    for (int k = 0; k < 8; ++k) { const unsigned c = warp_tot[k]; if (k < w) base += c; tot += c; }
    total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(256) scan_block_sums(const unsigned* __restrict__ in, size_t n,
                                                       unsigned* __restrict__ sums) {
    __shared__ unsigned wt[8];
    const size_t base = (size_t)blockIdx.x * kScanBlock + threadIdx.x * 4;
    unsigned v = 0;
    for (int k = 0; k < 4; ++k) if (base + k < n) v += in[base + k];
    unsigned total;
    block_scan_u32(v, wt, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) scan_sums_inplace(unsigned* sums, int nb, unsigned* grand_total) {
    __shared__ unsigned wt[8];
    __shared__ unsigned carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += 256) {
        const int i = b0 + threadIdx.x;
        const unsigned v = i < nb ? sums[i] : 0u;
        unsigned total;
        const unsigned ex = block_scan_u32(v, wt, total);
        const unsigned carry = carry_s;
        if (i < nb) sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry_s;
}

__global__ void __launch_bounds__(256) scan_apply(const unsigned* __restrict__ in, size_t n,
                                                  const unsigned* __restrict__ sums,
                                                  unsigned* __restrict__ out /* n + 1 */,
                                                  const unsigned* __restrict__ grand_total) {
    __shared__ unsigned wt[8];
    const size_t base = (size_t)blockIdx.x * kScanBlock + threadIdx.x * 4;
    unsigned v[4], s = 0;
    for (int k = 0; k < 4; ++k) { v[k] = base + k < n ? in[base + k] : 0u; s += v[k]; }
    unsigned total;
    unsigned run = sums[blockIdx.x] + block_scan_u32(s, wt, total);
    for (int k = 0; k < 4; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *grand_total;
}

// Active tiles, heaviest first (bucketed by log2 of their run count).
__global__ void __launch_bounds__(1024) occ_order_tiles(const unsigned* __restrict__ offsets, int n_tiles,
                                                        int chunk_scans, int* __restrict__ order,
                                                        int* __restrict__ n_active) {
    __shared__ int hist[33];
    __shared__ int start[33];
    if (threadIdx.x < 33) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
        const unsigned tot = offsets[(size_t)(t + 1) * chunk_scans] - offsets[(size_t)t * chunk_scans];
        if (tot) atomicAdd(&hist[32 - __clz(tot)], 1);           // bucket 1..32
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int b = 32; b >= 1; --b) { start[b] = run; run += hist[b]; }
        *n_active = run;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
        const unsigned tot = offsets[(size_t)(t + 1) * chunk_scans] - offsets[(size_t)t * chunk_scans];
        if (tot) order[atomicAdd(&start[32 - __clz(tot)], 1)] = t;
    }
}

// ---- 5. tile-resident replay ------------------------------------------------------
struct ApplyArgs {
    float* grid;
    int nx, ny, tiles_x;
    const unsigned* offsets;                  // [tile][scan] exclusive, + total
    const Run* runs;
    const int2* ray_cell;                     // indexed by run.ray (relative to chunk ray_begin)
    const int2* origin_cell;                  // indexed by scan - scan_begin
    int chunk_scans;
    const int* order;
    const int* n_active;
    unsigned* queue;
    double l_hit, l_miss;
    float lo, hi;
    // clamp interval that excludes 0 (see occupancy.h): cells still exactly 0
    // are "virgin"; after the first non-empty scan they read as clamp(0)
    int virgin_after;                         // first chunk-local scan index at which virgin cells read clamp(0); INT_MAX = never
    float clamp0;
    int* error_flag;
    long long* tile_prof;                     // optional: per queue slot {tile, scans, runs, cycles}
};

__device__ __forceinline__ float chain(float x, unsigned m, unsigned k, double l_hit, double l_miss,
                                       float lo, float hi) {
    // steady state of free space: the cell already sits on the lower clamp and only misses arrive
    // (every add keeps it at or below the clamp, the final clamp returns it) -- likewise at the top
    if (m == 0u) {
        if (l_miss < 0.0 && x <= lo) return lo;
        if (l_miss > 0.0 && x >= hi) return hi;
    }
    // mapping.py:129 -- m hits, each x = f32(f64(x) + l_hit)
    for (unsigned i = 0; i < m; ++i) {
        x = (float)((double)x + l_hit);
        if (k == 0 && ((l_hit > 0.0 && x >= hi) || (l_hit < 0.0 && x <= lo))) break;   // saturated, clamp follows
    }
    // mapping.py:139 -- k misses.  The adds are monotone, so once past the
    // clamp bound in the direction of travel the final clamp decides the value.
    // Each add moves x by l_miss up to one float rounding (< 1e-6 for |x| <= 16),
    // so k adds certainly cross the bound when k * (|l_miss| - 1e-6) covers the gap.
    if (k > 4 && fabsf(x) <= 16.f && lo >= -16.f && hi <= 16.f) {
        if (l_miss < -1e-5 && (double)x + (double)k * (l_miss + 1e-6) <= (double)lo) return lo;
        if (l_miss > 1e-5 && (double)x + (double)k * (l_miss - 1e-6) >= (double)hi) return hi;
    }
    if (l_miss != 0.0) {
        for (unsigned i = 0; i < k; ++i) {
            x = (float)((double)x + l_miss);
            if ((l_miss < 0.0 && x <= lo) || (l_miss > 0.0 && x >= hi)) break;
        }
    }
    return fminf(fmaxf(x, lo), hi);           // mapping.py:141
}

// One CTA per tile; the tile stays in shared memory while ITS scans are replayed in order.
// Scans are taken kOccBatch at a time, each with its own counter plane, so that one pair
// of barriers covers a whole batch:
//   count  a task = 32 consecutive runs of one scan (one per lane) x one window of their
//          common step range; the warp walks it in lock step over the ray step index n.
//          Rays of one scan share their origin, so neighbouring lanes sit on the same cell
//          for long stretches: each maximal group of equal neighbours issues ONE shared
//          atomicAdd carrying the group size (no same-address conflicts, far fewer atomics
//          near the sensor, nothing waits on a return value).  All tasks of the batch are
//          independent, so the warps stay busy instead of waiting on a per-scan barrier.
//   apply  every thread owns four cells: it keeps them in registers and runs, scan by scan
//          in order, the fp64->fp32 add chain + clamp for the non-zero counters.
constexpr int kWin = 16;                       // lock-step block (unrolled)
constexpr int kOccBatch = 8;                   // scans per barrier pair
constexpr int kOccList = 256;                  // scans compacted per pass over the offsets row

// Shared-memory slot of local cell idx = y * TS + x.  Rays of one scan that are
// x-major sit in the same column at a given step; without the swizzle their
// counters would all fall into one bank.
__device__ __forceinline__ int swz(int idx) { return idx ^ ((idx / TS) & 31); }

// One task = 32 consecutive runs of the batch (the runs of a tile are contiguous across its
// scans, so a warp is full even when single scans bring only a few runs).  `bounds` are the
// run-index boundaries between the batch's scans; lane e belongs to counter plane
// #{bounds <= e}.  Lanes are merged only with neighbours of the same scan.
__device__ __forceinline__ void occ_count_task(const ApplyArgs& a, unsigned e, unsigned end, const unsigned* bounds,
                                               int n_bounds, unsigned* cnt, int lane) {
    const unsigned above = lane == 31 ? 0u : (0xffffffffu << (lane + 1));
    int n0 = 0x7fffffff, nend = 0, j0 = 0, idx0 = 0, dmaj = 1, dmin = 0, plane = 0;
    // a tile's runs are contiguous across its scans: pull the records far ahead into L2
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.runs + e + 2048));
    if (e < end) {
        for (int i = 0; i < n_bounds; ++i) plane += e >= bounds[i];
        const int4 ra = __ldg(reinterpret_cast<const int4*>(a.runs + e));
        const int2 rb = __ldg(reinterpret_cast<const int2*>(a.runs + e) + 2);
        if (ra.z == 0) {                                           // hit: one add of the hit unit
            const unsigned old = atomicAdd(&cnt[plane * TCELLS + swz(ra.w & 0xffff)], kHitUnit);
            if ((old >> 20) == 4095u) *a.error_flag = 1;
        } else {
            n0 = ra.x; j0 = ra.y; nend = ra.x + ra.z; idx0 = ra.w; dmaj = rb.x; dmin = rb.y;
        }
    }
    // Every lane walks its own run from its first cell (local step k); a run never exceeds TS
    // cells, so two unrolled blocks cover it.  Lanes that sit on the same cell of the same scan
    // in the same step -- always the case near the sensor, where all rays start together --
    // are merged with their neighbours into one add.
    const int len = nend - n0 > 0 && n0 != 0x7fffffff ? nend - n0 : 0;
    const int maxlen = (int)__reduce_max_sync(0xffffffffu, (unsigned)len);
    if (maxlen == 0) return;                                       // no miss runs in this chunk
    unsigned* my = cnt + plane * TCELLS;
    const int prev_plane = __shfl_up_sync(0xffffffffu, plane, 1);
    const bool same_scan = lane > 0 && prev_plane == plane;
    int idx = idx0 & 0xffff, step_maj = 0, step_both = 0, d = 0, inc = 0, dec = 0;
    if (len > 0) {
        const int maj = (idx0 & kRunXMajor) ? ((idx0 & kRunMajPos) ? 1 : -1) : ((idx0 & kRunMajPos) ? TS : -TS);
        const int mnr = (idx0 & kRunXMajor) ? ((idx0 & kRunMinPos) ? TS : -TS) : ((idx0 & kRunMinPos) ? 1 : -1);
        step_maj = maj; step_both = maj + mnr;
        inc = 2 * dmin; dec = 2 * dmaj;
        d = (int)((2ll * n0 + 2) * dmin - (long long)dec * j0 - dmaj);      // RunWalker::start
    }
    for (int b = 0; b < maxlen; b += kWin) {
        // cells of this lane for the next kWin steps (pure ALU), then the votes + atomics
        int cell[kWin];
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            cell[k] = idx;
            if (b + k < len) { const bool m = d > 0; idx += m ? step_both : step_maj; d += inc - (m ? dec : 0); }
        }
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const bool act = b + k < len;
            const unsigned am = __ballot_sync(0xffffffffu, act);
            if (am == 0u) break;                                   // runs only get shorter
            const int prev = __shfl_up_sync(0xffffffffu, cell[k], 1);
            const bool prev_act = same_scan && ((am >> (lane - 1)) & 1u);
            const bool lead = act && !(prev_act && prev == cell[k]);
            const unsigned lm = __ballot_sync(0xffffffffu, lead);
            if (lead) {
                // group = this lane and the active lanes right after it on the same cell of the same scan
                const unsigned stop = (lm | ~am) & above;
                const int nxt = stop ? __ffs(stop) - 1 : 32;
                atomicAdd(&my[swz(cell[k])], (unsigned)(nxt - lane));        // result unused
            }
        }
    }
}

__global__ void __launch_bounds__(kOccNT) occ_tile_apply(const ApplyArgs a) {
    extern __shared__ __align__(16) unsigned char occ_smem[];
    float* tile = reinterpret_cast<float*>(occ_smem);
    unsigned* cnt = reinterpret_cast<unsigned*>(occ_smem + sizeof(float) * TCELLS);      // [kOccBatch][TCELLS]
    unsigned* lst_beg = cnt + kOccBatch * TCELLS;
    unsigned* lst_end = lst_beg + kOccList;
    int* lst_scan = reinterpret_cast<int*>(lst_end + kOccList);
    __shared__ int wcount[kOccNT / 32];
    __shared__ int n_list;
    __shared__ int cur_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int n_active = *a.n_active;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const unsigned q = atomicAdd(a.queue, 1u);
            cur_tile = q < (unsigned)n_active ? a.order[q] : -1;
        }
        __syncthreads();
        const int t = cur_tile;
        if (t < 0) break;
        const long long t_start = clock64();
        const int tx0 = (t % a.tiles_x) * TS, ty0 = (t / a.tiles_x) * TS;
        for (int c = tid; c < TCELLS; c += kOccNT) {
            const int x = tx0 + (c & (TS - 1)), y = ty0 + (c / TS);
            tile[swz(c)] = (x < a.nx && y < a.ny) ? a.grid[(size_t)y * a.nx + x] : 0.f;
        }
        for (int c = tid; c < kOccBatch * TCELLS; c += kOccNT) cnt[c] = 0u;
        const unsigned* off = a.offsets + (size_t)t * a.chunk_scans;
        int scans_done = 0;
        for (int s0 = 0; s0 < a.chunk_scans; s0 += kOccList) {
            // ---- non-empty scans of this block of the offsets row, ascending
            if (tid == 0) n_list = 0;
            __syncthreads();
            for (int s1 = s0; s1 < min(s0 + kOccList, a.chunk_scans); s1 += kOccNT) {
                const int s = s1 + tid;
                unsigned ob = 0, oe = 0;
                if (s < min(s0 + kOccList, a.chunk_scans)) { ob = off[s]; oe = off[s + 1]; }
                const bool has = oe > ob;
                const unsigned bal = __ballot_sync(0xffffffffu, has);
                if (lane == 0) wcount[warp] = __popc(bal);
                __syncthreads();
                int base = n_list;
                for (int w = 0; w < warp; ++w) base += wcount[w];
                if (has) {
                    const int slot = base + __popc(bal & lt_mask);
                    lst_scan[slot] = s; lst_beg[slot] = ob; lst_end[slot] = oe;
                }
                __syncthreads();
                if (tid == 0) { int tot = 0; for (int w = 0; w < kOccNT / 32; ++w) tot += wcount[w]; n_list += tot; }
                __syncthreads();
            }
            const int n_lst = n_list;
            scans_done += n_lst;
            for (int b0 = 0; b0 < n_lst; b0 += kOccBatch) {
                const int nb = min(kOccBatch, n_lst - b0);
                // ---- count: the batch's runs are one contiguous range; 32 runs per warp task
                const unsigned rb = lst_beg[b0], re = lst_end[b0 + nb - 1];
                for (unsigned e0 = rb + warp * 32u; e0 < re; e0 += kOccNT)
                    occ_count_task(a, e0 + lane, re, lst_end + b0, nb - 1, cnt, lane);
                __syncthreads();
                // ---- apply: kCpt cells per thread kept in registers, scans in order
                constexpr int kCpt = TCELLS / kOccNT >= 4 ? 4 : 2;
                for (int c0 = tid * kCpt; c0 < TCELLS; c0 += kOccNT * kCpt) {
                    float xv[kCpt];
#pragma unroll
                    for (int k = 0; k < kCpt; ++k) xv[k] = tile[c0 + k];
                    bool dirty = false;
                    for (int b = 0; b < nb; ++b) {
                        unsigned* plane = cnt + b * TCELLS;
                        unsigned cv[kCpt];
                        if (kCpt == 4) {
                            const uint4 c4 = *reinterpret_cast<const uint4*>(&plane[c0]);
                            cv[0] = c4.x; cv[1] = c4.y; cv[kCpt - 2] = c4.z; cv[kCpt - 1] = c4.w;
                        } else {
                            const uint2 c2 = *reinterpret_cast<const uint2*>(&plane[c0]);
                            cv[0] = c2.x; cv[1] = c2.y;
                        }
                        unsigned any = 0u;
#pragma unroll
                        for (int k = 0; k < kCpt; ++k) any |= cv[k];
                        if (any == 0u) continue;
#pragma unroll
                        for (int k = 0; k < kCpt; ++k) plane[c0 + k] = 0u;
                        const bool virgin_fix = lst_scan[b0 + b] >= a.virgin_after;
#pragma unroll
                        for (int k = 0; k < kCpt; ++k) {
                            if (cv[k]) {
                                float x = xv[k];
                                if (virgin_fix && x == 0.0f) x = a.clamp0;
                                xv[k] = chain(x, cv[k] >> 20, cv[k] & kMissMask, a.l_hit, a.l_miss, a.lo, a.hi);
                            }
                        }
                        dirty = true;
                    }
                    if (dirty) {
#pragma unroll
                        for (int k = 0; k < kCpt; ++k) tile[c0 + k] = xv[k];
                    }
                }
                __syncthreads();
            }
        }
        for (int c = tid; c < TCELLS; c += kOccNT) {
            const int x = tx0 + (c & (TS - 1)), y = ty0 + (c / TS);
            if (x < a.nx && y < a.ny) a.grid[(size_t)y * a.nx + x] = tile[swz(c)];
        }
        if (a.tile_prof && tid == 0) {
            long long* rec = a.tile_prof + 4ll * t;
            rec[0] = t; rec[1] = scans_done;
            rec[2] = (long long)(off[a.chunk_scans] - off[0]);
            rec[3] = clock64() - t_start;
        }
    }
}

// cells still exactly 0 after the first non-empty scan read as clamp(0)
// (only when the clamp interval excludes 0); owned tiles only
__global__ void occ_finalize_virgin(float* grid, int nx, int ny, int tiles_x, int rank, int world, float clamp0) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nx * ny) return;
    const int x = (int)(i % nx), y = (int)(i / nx);
    const int tile = (y / TS) * tiles_x + (x / TS);
    (void)tile;
    if (occ_owner(x, y, nx, ny, world) != rank) return;
    if (grid[i] == 0.0f) grid[i] = clamp0;
}

// ---- host orchestration -------------------------------------------------------------
static int exclusive_scan_u32(const unsigned* d_in, size_t n, unsigned* d_out, DevBuf& sums_buf,
                              unsigned* d_total, cudaStream_t st) {
    const int nb = (int)((n + kScanBlock - 1) / kScanBlock);
    if (sums_buf.reserve(sizeof(unsigned) * (size_t)(nb + 1))) return ICPB200_ERR_CUDA;
    unsigned* sums = sums_buf.as<unsigned>();
    scan_block_sums<<<nb, 256, 0, st>>>(d_in, n, sums);
    ICPB_LAUNCH_CHECK();
    scan_sums_inplace<<<1, 256, 0, st>>>(sums, nb, d_total);
    ICPB_LAUNCH_CHECK();
    scan_apply<<<nb, 256, 0, st>>>(d_in, n, sums, d_out, d_total);
    ICPB_LAUNCH_CHECK();
    return ICPB200_OK;
}

// slam.py:46-50 `transform_points_2d`: points_2d @ R.T + t.  numpy evaluates the 2-term dot product of the matmul as
// fma(p1, r1, rn(p0 * r0)) (OpenBLAS dgemm on FMA hardware; checked against exact rational arithmetic in the build
// container, oracle/pin_rebuild.py pins the result against the reference) and then adds t: the same three roundings here.
__global__ void occ_transform_kernel(int n_scans, long long n_points, const double* __restrict__ poses,
                                     const double2* __restrict__ local, const long long* __restrict__ off,
                                     double2* __restrict__ world, double* __restrict__ origins) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_scans) {                                   // slam.py:275: origin = pose[:2, 2]
        origins[2 * i] = poses[9 * i + 2];
        origins[2 * i + 1] = poses[9 * i + 5];
    }
    if (i >= n_points) return;
    int lo = 0, hi = n_scans;                            // largest s with off[s] <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    const double* P = poses + 9 * (size_t)lo;
    const double2 p = local[i];
    double2 w;
    w.x = __dadd_rn(__fma_rn(p.y, P[1], __dmul_rn(p.x, P[0])), P[2]);
    w.y = __dadd_rn(__fma_rn(p.y, P[4], __dmul_rn(p.x, P[3])), P[5]);
    world[i] = w;
}

int occ_transform_history(int n_scans, long long n_points, const double* d_poses, const double* d_local,
                          const long long* d_off, double* d_world, double* d_origins, cudaStream_t st) {
    const long long n = std::max<long long>(n_points, n_scans);
    if (n <= 0) return ICPB200_OK;
    occ_transform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n_scans, n_points, d_poses,
                                                                      reinterpret_cast<const double2*>(d_local), d_off,
                                                                      reinterpret_cast<double2*>(d_world), d_origins);
    ICPB_LAUNCH_CHECK();
    return ICPB200_OK;
}

int occ_update_device(OccGrid& g, int n_scans, const double* d_origins, const double* d_hits,
                      const long long* d_hit_off, const long long* h_hit_off, cudaStream_t st) {
    if (!(g.use_fast && !g.zero_outside_clamp)) g.all_dirty = true;       // the ordered replay does not track touched tiles
    if (g.use_fast && !g.zero_outside_clamp) {
        const int rc = occ_update_fast(g, n_scans, d_origins, d_hits, d_hit_off, h_hit_off, h_hit_off[n_scans] - h_hit_off[0],
                                       false, st);
        if (rc < 0) { g.slotmap.release(); g.ord.release(); g.ncount.release(); }      // claims may be left behind: start clean next time
        return rc;
    }
    return occ_update_ordered(g, n_scans, d_origins, d_hits, d_hit_off, h_hit_off, st);
}

int occ_update_ordered(OccGrid& g, int n_scans, const double* d_origins, const double* d_hits,
                       const long long* d_hit_off, const long long* h_hit_off, cudaStream_t st) {
    const long long n_rays = h_hit_off[n_scans] - h_hit_off[0];
    g.stats[0] = n_rays; g.stats[1] = g.stats[2] = g.stats[3] = 0;
    if (n_rays <= 0) return ICPB200_OK;                                   // mapping.py:113-114
    if (n_rays > 0x7fffffffLL) { set_error("grid_update: more than 2^31-1 rays in one call"); return ICPB200_ERR_LIMIT; }
    for (int s = 0; s < n_scans; ++s)
        if (h_hit_off[s + 1] - h_hit_off[s] >= (long long)kHitUnit) {
            set_error("grid_update: scan %d has %lld rays; this build supports < %u rays per scan", s,
                      (long long)(h_hit_off[s + 1] - h_hit_off[s]), kHitUnit);
            return ICPB200_ERR_LIMIT;
        }
    const int n_tiles = g.tiles_x * g.tiles_y;
    if (g.origin_cell.reserve(sizeof(int2) * (size_t)n_scans) || g.ray_cell.reserve(sizeof(int2) * (size_t)n_rays) ||
        g.ray_scan.reserve(sizeof(int) * (size_t)n_rays) || g.order.reserve(sizeof(int) * (size_t)n_tiles) ||
        g.small.reserve(256))
        return ICPB200_ERR_CUDA;
    // small: [0] grand total (u32) [1] n_active [2] queue [3] error flag ; stats at +64 bytes (4 x u64)
    unsigned* d_small = g.small.as<unsigned>();
    unsigned long long* d_stats = reinterpret_cast<unsigned long long*>(g.small.as<unsigned char>() + 64);
    ICPB_CUDA(cudaMemsetAsync(g.small.p, 0, 256, st));

    occ_scan_setup<<<(n_scans + 255) / 256, 256, 0, st>>>(d_origins, n_scans, g.min_x, g.min_y, g.res,
                                                          g.origin_cell.as<int2>());
    ICPB_LAUNCH_CHECK();
    occ_ray_setup<<<(unsigned)((n_rays + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const double2*>(d_hits), d_hit_off, n_scans, n_rays, g.min_x, g.min_y, g.res,
        g.ray_cell.as<int2>(), g.ray_scan.as<int>());
    ICPB_LAUNCH_CHECK();

    // scan chunks bound the (tile, scan) matrix
    int chunk = (int)std::min<long long>(kOccMaxChunkScans, std::max<long long>(1, kOccMaxMatrix / n_tiles));
    for (int s0 = 0; s0 < n_scans; s0 += chunk) {
        const int cs = std::min(chunk, n_scans - s0);
        const long long rb = h_hit_off[s0] - h_hit_off[0], re = h_hit_off[s0 + cs] - h_hit_off[0];
        if (re == rb) continue;
        const size_t cells = (size_t)n_tiles * cs;
        if (g.counts.reserve(sizeof(unsigned) * cells) || g.offsets.reserve(sizeof(unsigned) * (cells + 1)))
            return ICPB200_ERR_CUDA;
        ICPB_CUDA(cudaMemsetAsync(g.counts.p, 0, sizeof(unsigned) * cells, st));
        BinArgs b;
        b.ray_cell = g.ray_cell.as<int2>() + rb;       // run.ray is relative to the chunk
        b.ray_scan = g.ray_scan.as<int>() + rb;
        b.origin_cell = g.origin_cell.as<int2>();
        b.ray_begin = 0; b.ray_end = re - rb;
        b.scan_begin = s0; b.chunk_scans = cs;
        b.nx = g.nx; b.ny = g.ny; b.tiles_x = g.tiles_x;
        b.rank = g.rank; b.world = g.world;
        b.counts = g.counts.as<unsigned>();
        b.offsets = nullptr; b.runs = nullptr;
        b.stats = d_stats;
        const unsigned nblk = (unsigned)((re - rb + 255) / 256);
        occ_bin<false><<<nblk, 256, 0, st>>>(b);
        ICPB_LAUNCH_CHECK();
        int rc = exclusive_scan_u32(g.counts.as<unsigned>(), cells, g.offsets.as<unsigned>(), g.sums, d_small, st);
        if (rc) return rc;
        unsigned total_runs = 0;
        ICPB_CUDA(cudaMemcpyAsync(&total_runs, d_small, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        ICPB_CUDA(cudaStreamSynchronize(st));
        if (total_runs == 0) continue;
        if (g.runs.reserve(sizeof(Run) * ((size_t)total_runs + 4096))) return ICPB200_ERR_CUDA;   // + prefetch slack
        ICPB_CUDA(cudaMemsetAsync(g.counts.p, 0, sizeof(unsigned) * cells, st));
        b.offsets = g.offsets.as<unsigned>();
        b.runs = g.runs.as<Run>();
        occ_bin<true><<<nblk, 256, 0, st>>>(b);
        ICPB_LAUNCH_CHECK();
        occ_order_tiles<<<1, 1024, 0, st>>>(g.offsets.as<unsigned>(), n_tiles, cs, g.order.as<int>(),
                                            reinterpret_cast<int*>(d_small + 1));
        ICPB_LAUNCH_CHECK();
        ICPB_CUDA(cudaMemsetAsync(d_small + 2, 0, sizeof(unsigned), st));
        ApplyArgs ap;
        ap.grid = g.grid.as<float>();
        ap.nx = g.nx; ap.ny = g.ny; ap.tiles_x = g.tiles_x;
        ap.offsets = g.offsets.as<unsigned>();
        ap.runs = g.runs.as<Run>();
        ap.ray_cell = g.ray_cell.as<int2>() + rb;
        ap.origin_cell = g.origin_cell.as<int2>() + s0;
        ap.chunk_scans = cs;
        ap.order = g.order.as<int>();
        ap.n_active = reinterpret_cast<int*>(d_small + 1);
        ap.queue = d_small + 2;
        ap.l_hit = g.l_hit; ap.l_miss = g.l_miss;
        ap.lo = (float)g.lo_min; ap.hi = (float)g.lo_max;
        ap.clamp0 = fminf(fmaxf(0.f, ap.lo), ap.hi);
        ap.virgin_after = 0x7fffffff;
        if (g.zero_outside_clamp) {
            if (g.seen_nonempty_scan) ap.virgin_after = 0;
            else {
                // the first non-empty scan of this chunk still sees true zeros
                int first = 0;
                while (first < cs && h_hit_off[s0 + first + 1] == h_hit_off[s0 + first]) ++first;
                ap.virgin_after = first + 1;
            }
        }
        ap.error_flag = reinterpret_cast<int*>(d_small + 3);
        ap.tile_prof = nullptr;
        if (g.profile_tiles) {
            if (g.tile_prof.reserve(sizeof(long long) * 4 * (size_t)n_tiles)) return ICPB200_ERR_CUDA;
            ICPB_CUDA(cudaMemsetAsync(g.tile_prof.p, 0, sizeof(long long) * 4 * (size_t)n_tiles, st));
            ap.tile_prof = g.tile_prof.as<long long>();
        }
        ICPB_CUDA(cudaFuncSetAttribute(occ_tile_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOccSmem));
        occ_tile_apply<<<g.apply_ctas, kOccNT, kOccSmem, st>>>(ap);
        ICPB_LAUNCH_CHECK();
        g.seen_nonempty_scan = true;
    }
    if (g.zero_outside_clamp && !g.virgin_finalised && g.seen_nonempty_scan) {
        const size_t n = (size_t)g.nx * g.ny;
        const float lo = (float)g.lo_min, hi = (float)g.lo_max;
        occ_finalize_virgin<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g.grid.as<float>(), g.nx, g.ny, g.tiles_x,
                                                                         g.rank, g.world, fminf(fmaxf(0.f, lo), hi));
        ICPB_LAUNCH_CHECK();
        g.virgin_finalised = true;
    }
    // stats + error flag (one small readback; also orders the host after the work)
    unsigned char host_small[256];
    ICPB_CUDA(cudaMemcpyAsync(host_small, g.small.p, 256, cudaMemcpyDeviceToHost, st));
    ICPB_CUDA(cudaStreamSynchronize(st));
    const unsigned long long* hs = reinterpret_cast<const unsigned long long*>(host_small + 64);
    g.stats[1] = (long long)hs[1]; g.stats[2] = (long long)hs[2]; g.stats[3] = (long long)hs[3];
    if (reinterpret_cast<const int*>(host_small)[3]) {
        set_error("grid_update: more than 4095 hits landed in one cell within one scan (unsupported)");
        return ICPB200_ERR_LIMIT;
    }
    return ICPB200_OK;
}

int occ_apply_ctas(int sm_count) {
    int per_sm = 0;
    cudaFuncSetAttribute(occ_tile_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOccSmem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, occ_tile_apply, kOccNT, kOccSmem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    return sm_count * per_sm;
}

}  // namespace icpb
