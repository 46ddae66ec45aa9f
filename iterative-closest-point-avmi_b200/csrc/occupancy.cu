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
// t % world == rank and skips all others in step 2.
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
constexpr size_t kOccSmem = sizeof(float) * TCELLS + 2 * sizeof(unsigned) * TCELLS + sizeof(unsigned short) * kOccMaxChunkScans;

struct Run {                                  // 16 bytes
    unsigned int ray;                         // global ray index
    int n0;                                   // first step (miss run) / local cell (hit)
    int j0;                                   // minor steps at n0
    int len;                                  // cells in run; 0 marks a hit
};

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

template <bool FILL>
__global__ void __launch_bounds__(256) occ_bin(const BinArgs a) {
    const long long r = a.ray_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cells = 0, hits = 0, nruns = 0;
    if (r < a.ray_end) {
        const int s = a.ray_scan[r];
        const int sl = s - a.scan_begin;
        const int2 o = a.origin_cell[s], h = a.ray_cell[r];
        const RayGeom g = make_ray(o.x, o.y, h.x, h.y);
        for_each_tile_run<TS>(g, a.nx, a.ny, a.tiles_x, [&](const TileRun& t) {
            if (t.tile % a.world != a.rank) return;
            const size_t gi = (size_t)t.tile * a.chunk_scans + sl;
            if (FILL) {
                const unsigned slot = atomicAdd(&a.counts[gi], 1u);
                Run run;
                run.ray = (unsigned)(r - a.ray_begin);
                run.n0 = t.n0; run.j0 = t.j0; run.len = t.len;
                reinterpret_cast<int4*>(a.runs)[a.offsets[gi] + slot] = *reinterpret_cast<int4*>(&run);
            } else {
                atomicAdd(&a.counts[gi], 1u);
                cells += t.len;
                ++nruns;
            }
        });
        if (h.x >= 0 && h.x < a.nx && h.y >= 0 && h.y < a.ny) {       // mapping.py:124-127
            const int tile = (h.y / TS) * a.tiles_x + (h.x / TS);
            if (tile % a.world == a.rank) {
                const size_t gi = (size_t)tile * a.chunk_scans + sl;
                if (FILL) {
                    const unsigned slot = atomicAdd(&a.counts[gi], 1u);
                    Run run;
                    run.ray = (unsigned)(r - a.ray_begin);
                    run.n0 = (h.y % TS) * TS + (h.x % TS); run.j0 = 0; run.len = 0;
                    reinterpret_cast<int4*>(a.runs)[a.offsets[gi] + slot] = *reinterpret_cast<int4*>(&run);
                } else {
                    atomicAdd(&a.counts[gi], 1u);
                    ++hits;
                    ++nruns;
                }
            }
        }
    }
    if (!FILL) {
        // block totals -> four global atomics per block
        __shared__ unsigned long long part[3][8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cells += __shfl_xor_sync(0xffffffffu, cells, o);
            hits += __shfl_xor_sync(0xffffffffu, hits, o);
            nruns += __shfl_xor_sync(0xffffffffu, nruns, o);
        }
        const int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) { part[0][w] = cells; part[1][w] = hits; part[2][w] = nruns; }
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
};

__device__ __forceinline__ float chain(float x, unsigned m, unsigned k, double l_hit, double l_miss,
                                       float lo, float hi) {
    // mapping.py:129 -- m hits, each x = f32(f64(x) + l_hit)
    for (unsigned i = 0; i < m; ++i) {
        x = (float)((double)x + l_hit);
        if (k == 0 && ((l_hit > 0.0 && x >= hi) || (l_hit < 0.0 && x <= lo))) break;   // saturated, clamp follows
    }
    // mapping.py:139 -- k misses.  The adds are monotone, so once past the
    // clamp bound in the direction of travel the final clamp decides the value.
    if (l_miss != 0.0) {
        for (unsigned i = 0; i < k; ++i) {
            x = (float)((double)x + l_miss);
            if ((l_miss < 0.0 && x <= lo) || (l_miss > 0.0 && x >= hi)) break;
        }
    }
    return fminf(fmaxf(x, lo), hi);           // mapping.py:141
}

// Work item = one "slot" of a run: slot k of a miss run covers its cells
// [8k, 8k+8); a hit is slot 0 of a zero-length run.  A run has TS/8 slots, most
// of them empty for short runs; consecutive lanes share a run, so its 16-byte
// record is one broadcast load.
constexpr int kSub = 8;
constexpr int kSlots = TS / kSub;

struct Slot {
    int idx;            // first local cell index
    int ncell;          // cells in this slot (0: nothing to do)
    int maj_stride, min_stride;
    RunWalker wk;
    bool hit;
};

__device__ __forceinline__ Slot load_slot(const ApplyArgs& a, unsigned item, int2 o, int tx0, int ty0) {
    Slot s;
    s.ncell = 0; s.hit = false; s.idx = 0; s.maj_stride = 0; s.min_stride = 0;
    const int4 raw = __ldg(reinterpret_cast<const int4*>(a.runs) + item / kSlots);
    const int sub = (int)(item % kSlots);
    if (raw.w == 0) {                                      // hit: n0 is the local cell
        if (sub == 0) { s.hit = true; s.ncell = 1; s.idx = raw.y; }
        return s;
    }
    const int first = sub * kSub;
    if (first >= raw.w) return s;
    s.ncell = min(kSub, raw.w - first);
    const int2 h = __ldg(&a.ray_cell[(unsigned)raw.x]);
    const RayGeom g = make_ray(o.x, o.y, h.x, h.y);
    const int n = raw.y + first;
    int j = raw.z;
    if (sub != 0) {
        // minor_steps(g, n); 32-bit division whenever the numerator fits
        const unsigned long long num = 2ull * (unsigned)n * (unsigned)g.dmin + (unsigned)g.dmaj - 1u;
        j = (num >> 32) == 0 ? (int)((unsigned)num / (2u * (unsigned)g.dmaj)) : (int)(num / (2ull * (unsigned)g.dmaj));
    }
    int x, y;
    cell_at(g, n, j, x, y);
    s.idx = (y - ty0) * TS + (x - tx0);
    s.maj_stride = g.xmajor ? g.smaj : g.smaj * TS;
    s.min_stride = g.xmajor ? g.smin * TS : g.smin;
    s.wk.start(g, n, j);
    return s;
}

__global__ void __launch_bounds__(kOccNT) occ_tile_apply(const ApplyArgs a) {
    extern __shared__ __align__(16) unsigned char occ_smem[];
    float* tile = reinterpret_cast<float*>(occ_smem);
    // double-buffered counters: count scan s+1 while applying scan s
    unsigned (*cnt)[TCELLS] = reinterpret_cast<unsigned (*)[TCELLS]>(occ_smem + sizeof(float) * TCELLS);
    unsigned short* scan_list = reinterpret_cast<unsigned short*>(occ_smem + sizeof(float) * TCELLS + 2 * sizeof(unsigned) * TCELLS);
    __shared__ int wcount[kOccNT / 32];
    __shared__ int n_list;
    __shared__ int cur_tile;
    const int tid = threadIdx.x;
    const int n_active = *a.n_active;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const unsigned q = atomicAdd(a.queue, 1u);
            cur_tile = q < (unsigned)n_active ? a.order[q] : -1;
            n_list = 0;
        }
        __syncthreads();
        const int t = cur_tile;
        if (t < 0) break;
        const int tx0 = (t % a.tiles_x) * TS, ty0 = (t / a.tiles_x) * TS;
        for (int c = tid; c < TCELLS; c += kOccNT) {
            const int x = tx0 + (c & (TS - 1)), y = ty0 + (c / TS);
            tile[c] = (x < a.nx && y < a.ny) ? a.grid[(size_t)y * a.nx + x] : 0.f;
            cnt[0][c] = 0u;
            cnt[1][c] = 0u;
        }
        // scans that have runs in this tile, ascending
        const unsigned* off = a.offsets + (size_t)t * a.chunk_scans;
        for (int s0 = 0; s0 < a.chunk_scans; s0 += kOccNT) {
            const int s = s0 + tid;
            const bool has = s < a.chunk_scans && off[s + 1] > off[s];
            const unsigned bal = __ballot_sync(0xffffffffu, has);
            if ((tid & 31) == 0) wcount[tid >> 5] = __popc(bal);
            __syncthreads();
            int base = n_list;
            for (int w = 0; w < (tid >> 5); ++w) base += wcount[w];
            if (has) scan_list[base + __popc(bal & ((1u << (tid & 31)) - 1u))] = (unsigned short)s;
            __syncthreads();
            if (tid == 0) { int tot = 0; for (int w = 0; w < kOccNT / 32; ++w) tot += wcount[w]; n_list += tot; }
            __syncthreads();
        }
        const int n_scans_here = n_list;

        auto count_scan = [&](int li) {
            const int s = scan_list[li];
            unsigned* c = cnt[li & 1];
            const unsigned beg = off[s] * kSlots, end = off[s + 1] * kSlots;
            const int2 o = a.origin_cell[s];
            for (unsigned item = beg + tid; item < end; item += kOccNT) {
                Slot sl = load_slot(a, item, o, tx0, ty0);
                if (sl.hit) {
                    const unsigned old = atomicAdd(&c[sl.idx], kHitUnit);
                    if ((old >> 20) == 4095u) *a.error_flag = 1;
                } else {
                    int idx = sl.idx;
                    for (int k = 0; k < sl.ncell; ++k) {
                        atomicAdd(&c[idx], 1u);                       // result unused: fire and forget
                        idx += sl.maj_stride + (sl.wk.step() ? sl.min_stride : 0);
                    }
                }
            }
        };
        auto apply_scan = [&](int li) {
            const int s = scan_list[li];
            unsigned* c = cnt[li & 1];
            const unsigned beg = off[s] * kSlots, end = off[s + 1] * kSlots;
            const int2 o = a.origin_cell[s];
            const bool virgin_fix = s >= a.virgin_after;
            for (unsigned item = beg + tid; item < end; item += kOccNT) {
                Slot sl = load_slot(a, item, o, tx0, ty0);
                if (sl.ncell == 0) continue;
                int idxs[kSub];
                unsigned got[kSub];
                int idx = sl.idx;
#pragma unroll
                for (int k = 0; k < kSub; ++k) {
                    idxs[k] = idx;
                    if (k < sl.ncell && !sl.hit) idx += sl.maj_stride + (sl.wk.step() ? sl.min_stride : 0);
                }
                // the thread whose exchange returns a non-zero count owns the cell for this scan
#pragma unroll
                for (int k = 0; k < kSub; ++k) got[k] = k < sl.ncell ? atomicExch(&c[idxs[k]], 0u) : 0u;
#pragma unroll
                for (int k = 0; k < kSub; ++k) {
                    if (got[k]) {
                        float x = tile[idxs[k]];
                        if (virgin_fix && x == 0.0f) x = a.clamp0;
                        tile[idxs[k]] = chain(x, got[k] >> 20, got[k] & kMissMask, a.l_hit, a.l_miss, a.lo, a.hi);
                    }
                }
            }
        };

        if (n_scans_here > 0) count_scan(0);
        __syncthreads();
        for (int li = 0; li < n_scans_here; ++li) {
            if (li + 1 < n_scans_here) count_scan(li + 1);    // other counter buffer
            apply_scan(li);
            __syncthreads();
        }
        for (int c = tid; c < TCELLS; c += kOccNT) {
            const int x = tx0 + (c & (TS - 1)), y = ty0 + (c / TS);
            if (x < a.nx && y < a.ny) a.grid[(size_t)y * a.nx + x] = tile[c];
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
    if (tile % world != rank) return;
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

int occ_update_device(OccGrid& g, int n_scans, const double* d_origins, const double* d_hits,
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
        if (g.runs.reserve(sizeof(Run) * (size_t)total_runs)) return ICPB200_ERR_CUDA;
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
