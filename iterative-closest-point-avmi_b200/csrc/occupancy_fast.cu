// Occupancy-grid log-odds raycast, order-free formulation (sm_100a) -- the
// default path behind icpb200_grid_update for
// /root/reference/utilities/mapping.py:103-141 applied to a batch of scans.
//
// The update is x = f32(f64(x) + c) per event with one clip per scan
// (mapping.py:129, 139, 141), so in general the value of a cell depends on the
// ORDER of the scans that touch it.  But a cell that receives NO hit in the
// batch only ever sees adds of l_miss: the chain is monotone, the clip can only
// pin it at the bound it is moving towards, and
//
//        final = clip(chain_N(x0))        N = total misses on the cell
//
// whatever way the N misses are spread over the scans (proof in DESIGN.md 3.4).
// So only the cells that are some ray's endpoint ("hit cells", the walls: 0.2 %
// of the touched cells on the C4 workload, 0.5 % of the ray-cell events) need an
// ordered replay.  The pipeline per chunk of <= 2048 scans:
//
//   1. occ_fast_count   per ray: world -> cell (fp64, exact), claim a slot for the
//                       endpoint's cell, count the ray's tile crossings per tile
//   2. occ_tile_scan    exclusive scan of the per-tile run counts, active tiles
//                       ordered heaviest first
//   3. occ_fast_fill    per ray: hit -> ord[slot][scan] (global atomic), tile runs
//                       (16-byte self-contained records) grouped by tile
//   4. occ_fast_tiles   one CTA per tile: every run is walked, in any order, by
//                       one lane; misses on ordinary cells are counted in shared
//                       memory, misses on hit cells go to ord[slot][scan]; then
//                       x = clip(chain_N(x)) on the counted cells
//   5. occ_fast_replay  one warp per hit cell: its row of ord is replayed scan by
//                       scan (m hits, k misses, clip) and cleared
//
// No barrier per scan, no (tile x scan) matrix, nothing sequential but the
// per-hit-cell chains.  Clamp intervals that exclude 0 (every untouched cell
// moves on the first scan) and hit-cell tables beyond kOrdBudget take the
// ordered tile replay in occupancy.cu instead.
#include "icp_b200.h"
#include "bres.cuh"
#include "common.cuh"
#include "occupancy.h"

#include <stdlib.h>

#include <algorithm>
#include <vector>

namespace icpb {

namespace {

constexpr int TS = kOccOwnTile;               // 64-cell tiles: an owned block is exactly one tile
constexpr int TCELLS = TS * TS;
constexpr unsigned kNone = 0xffffffffu;      // slot map: not a hit cell
constexpr unsigned kClaimed = 0xfffffffeu;   // slot map: being assigned
constexpr unsigned kHitUnit = 1u << 20;      // ord word: hits << 20 | misses
constexpr unsigned kMissMask = kHitUnit - 1u;
constexpr int kTileNT = 256;
// words of `small` beyond the per-chunk part: written by occ_fast_origins when the host has not seen the offsets
constexpr int kFlagBigScan = 32, kFlagBadOffsets = 33;
constexpr int kLenShift = 3, kLenClasses = TS >> kLenShift;      // runs are grouped by length: 1-8, 9-16, ... 57-64

// 16-byte run record: everything the tile kernel needs to walk the run.
//   w0 = local cell of the first step (12) | x-major (1) | major + (1) | minor + (1) | len-1 (6) | chunk-local scan (11)
//   w1 = Bresenham error term at the first step   w2 = dmaj   w3 = dmin
constexpr int kRunXMajor = 1 << 12, kRunMajPos = 1 << 13, kRunMinPos = 1 << 14;
static_assert(TS == 64 && kOccMaxChunkScans <= 2048, "run record bit layout");
static_assert(kLenClasses == 8, "occ_tile_scan reads the class counters as two uint4");

struct FastArgs {
    // rays
    const double2* hits;
    const long long* hit_off;                 // whole call, n_scans + 1
    const double* origins;
    long long ray_begin, ray_end;             // rays of this chunk (absolute)
    int scan_begin, chunk_scans, n_scans;
    double min_x, min_y, res;
    int nx, ny, tiles_x, n_tiles;
    int rank, world;                          // this rank walks the bands b with b % world == rank (occ_owner)
    int2* ray_cell;                           // chunk-relative
    int* ray_scan;
    int2* origin_cell;                        // absolute scan index
    // hit cells
    unsigned* slotmap;                        // ny * nx
    unsigned* slot_cell;                      // slot -> cell
    unsigned* ord;                            // [slot][ord_stride]
    int ord_stride;
    // binning
    unsigned* tile_count;                     // count pass: += 1 ; fill pass: cursor
    unsigned* tile_flag;                      // count pass: 1 = the tile holds a hit cell
    const unsigned* tile_off;                 // fill pass: start of every (tile, length class) segment
    uint4* runs;
    unsigned long long runs_cap, ord_cap;     // capacities (records / words): a fill launched before the host knows the
                                              // totals leaves without a write when they do not fit
    // small: [0] total runs [1] items [2] queue A [3] error flag [4] n_slots [5] multi-item tiles
    //        [6] items of tiles with hit cells (they come first) [7] queue B ; stats (u64 x 4) at +64 bytes
    unsigned* small;
    unsigned long long* stats;
};


// Tile crossings of one ray with the divisions done in 32 bits whenever the ray
// is short enough (always, for endpoints inside a <= 16k-cell grid); same runs
// in the same order as TileRunIter<TS>.
// SHARDED = false: one GPU, the band is the whole grid and none of the band state exists (the ray passes run at 32
// registers per thread).
template <bool SHARDED>
struct FastTileIter {
    RayGeom g0;                                 // the ray in grid coordinates (SHARDED only)
    RayGeom g;                                  // ... in the coordinates of the band being walked
    int n, nb;
    bool small;
    int tiles_x, tile_row0;
    int nx, ny, world;
    int band, band_last;                        // owned bands (tile rows) still to visit: band, band + world, ... <= band_last
    __device__ __forceinline__ void init_empty() { n = 0; nb = 0; small = true; tile_row0 = 0; band = 1; band_last = 0; world = 1; }
    // rows [ya, yb) are one band (the whole grid on one GPU): the ray is clipped to it in closed form
    __device__ __forceinline__ void init_band(int ya, int yb) {
        if (SHARDED) g = g0;
        g.oy -= ya;                                            // band coordinates; ya is a multiple of TS
        tile_row0 = ya / TS;
        int64_t a = 0, b = 0;
        if (g.dmaj != 0) clip_to_grid(g, nx, yb - ya, a, b);
        n = (int)a; nb = (int)b;
        small = g.dmaj < (1 << 14) && nb < (1 << 14);          // 2*n*dmin + dmaj < 2^30
    }
    // this rank walks the bands b with b % world == rank (occ_owner); one GPU: a single band, the whole grid
    __device__ __forceinline__ void init(const RayGeom& geom, int nx_, int ny_, int rank, int world_, int tiles_x_) {
        tiles_x = tiles_x_; nx = nx_; ny = ny_;
        n = 0; nb = 0; small = true; tile_row0 = 0;
        if (!SHARDED) { g = geom; init_band(0, ny); return; }
        g0 = geom; world = world_;
        // rows the ray can touch: between its origin and its endpoint (the endpoint itself is not walked, no matter)
        const int y_end = g0.xmajor ? g0.oy + g0.smin * g0.dmin : g0.oy + g0.smaj * g0.dmaj;
        int ylo = min(g0.oy, y_end), yhi = max(g0.oy, y_end);
        if (yhi < 0 || ylo >= ny || g0.dmaj == 0) { band = 1; band_last = 0; return; }
        ylo = max(ylo, 0); yhi = min(yhi, ny - 1);
        const int blo = ylo / TS;
        band_last = yhi / TS;
        band = blo + occ_mod_world(rank - blo + (blo / world + 1) * world, world);      // first owned band at or after blo
    }
    __device__ __forceinline__ int minor_at(int nn) const {
        if (small) return (int)((2u * (unsigned)nn * (unsigned)g.dmin + (unsigned)g.dmaj - 1u) / (2u * (unsigned)g.dmaj));
        return minor_steps(g, nn);
    }
    __device__ __forceinline__ long long reach(int j) const {
        if (small && j < (1 << 14)) {
            const unsigned num = 2u * (unsigned)g.dmaj * (unsigned)j - (unsigned)g.dmaj + 1u, den = 2u * (unsigned)g.dmin;
            return (long long)((num + den - 1u) / den);
        }
        return first_step_reaching(g, j);
    }
    __device__ __forceinline__ bool next(TileRun& r) {
        while (n >= nb) {                                      // this band is done: on to the next one this rank owns
            if (!SHARDED || band > band_last) return false;
            init_band(band * TS, min(band * TS + TS, ny));
            band += world;
        }
        const int j = minor_at(n);
        int x, y;
        cell_at(g, n, j, x, y);
        const int cmaj = g.xmajor ? x : y, cmin = g.xmajor ? y : x;
        const int inmaj = cmaj & (TS - 1);
        long long end = (long long)n + (g.smaj > 0 ? (TS - inmaj) : (inmaj + 1));
        if (g.dmin != 0) {
            const int inmin = cmin & (TS - 1);
            const int jleave = j + (g.smin > 0 ? (TS - inmin) : (inmin + 1));
            const long long nleave = reach(jleave);
            if (nleave < end) end = nleave;
        }
        if (end > nb) end = nb;
        r.tile = (y / TS + tile_row0) * tiles_x + (x / TS);
        r.n0 = n; r.j0 = j; r.len = (int)(end - n);
        n = (int)end;
        return true;
    }
};

// ---- 1./3. per-ray passes --------------------------------------------------------
__global__ void occ_fast_origins(const double* __restrict__ origins, int n_scans, double min_x, double min_y,
                                 double res, int2* __restrict__ origin_cell, const long long* __restrict__ hit_off,
                                 long long total_hits, unsigned* __restrict__ flags) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_scans) return;
    // mapping.py:57-60: floor((w - min) / resolution)
    origin_cell[s] = make_int2(sat_cell(floor((origins[2 * s] - min_x) / res)),
                               sat_cell(floor((origins[2 * s + 1] - min_y) / res)));
    if (flags) {
        // the host has not seen the offsets (device-resident entry point): check them here.
        // flags[0] = some scan has more than 4095 rays (the fill pass then needs the hit atomic's return value)
        // flags[1] = offsets inconsistent (hit_off[0] != 0, decreasing, or hit_off[n_scans] != total_hits)
        const long long b = hit_off[s], e = hit_off[s + 1];
        if (e - b > 4095) flags[0] = 1u;
        // (a scan of 2^20 rays or more does not fit the event word either: refused like inconsistent offsets -- the
        // host-buffer entry point sends such a scan through the ordered replay instead)
        if (e < b || (s == 0 && b != 0) || (s == n_scans - 1 && e != total_hits) || e - b >= (long long)kHitUnit) flags[1] = 1u;
    }
}

// CHECK: some scan has more than 4095 rays, so a cell could collect more hits in one scan than
// the 12-bit field holds; only then the hit atomic needs its return value.
template <bool FILL, bool CHECK, bool SHARDED>
__global__ void __launch_bounds__(256, 8) occ_fast_rays(const FastArgs a) {
    if (a.small[kFlagBadOffsets]) return;                     // offsets checked on the device: the host reports the error
    if (FILL && ((unsigned long long)a.small[0] + 64ull > a.runs_cap ||
                 (unsigned long long)a.small[4] * (unsigned long long)a.ord_stride > a.ord_cap ||
                 (!CHECK && a.small[kFlagBigScan])))
        return;                                               // speculative launch, buffers too small (or the CHECK variant is needed): the host repeats it
    const long long rl = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // chunk-relative ray
    const long long r = a.ray_begin + rl;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    __shared__ int cta_scan0_s;
    int cta_scan0 = 0;
    if (!FILL) {
        if (threadIdx.x == 0) {
            const long long r0 = a.ray_begin + (long long)blockIdx.x * blockDim.x;
            int lo = a.scan_begin, hi = a.scan_begin + a.chunk_scans;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (a.hit_off[mid] <= r0) lo = mid; else hi = mid;
            }
            cta_scan0_s = lo;
        }
        __syncthreads();
        cta_scan0 = cta_scan0_s;
    }
    unsigned long long cells = 0, nhits = 0, nruns = 0;
    FastTileIter<SHARDED> it;
    it.init_empty();
    int sl = 0;
    unsigned hit_old = 0u;
    if (r < a.ray_end) {
        int2 h;
        int s;
        if (!FILL) {
            const double2 p = a.hits[r];
            // mapping.py:94-98
            h = make_int2(sat_cell(floor((p.x - a.min_x) / a.res)), sat_cell(floor((p.y - a.min_y) / a.res)));
            // largest s with hit_off[s] <= r: one binary search per CTA (its first ray), then a short walk -- the rays of
            // a CTA belong to one or two scans
            s = cta_scan0;
            const int s_end = a.scan_begin + a.chunk_scans;
            while (s + 1 < s_end && a.hit_off[s + 1] <= r) ++s;
            a.ray_cell[rl] = h;
            a.ray_scan[rl] = s;
        } else {
            h = a.ray_cell[rl];
            s = a.ray_scan[rl];
        }
        sl = s - a.scan_begin;
        const int2 o = a.origin_cell[s];
        it.init(make_ray(o.x, o.y, h.x, h.y), a.nx, a.ny, a.rank, a.world, a.tiles_x);
        if (h.x >= 0 && h.x < a.nx && h.y >= 0 && h.y < a.ny) {                 // mapping.py:124-127
            const int tile = (h.y / TS) * a.tiles_x + (h.x / TS);
            if (occ_owner(h.x, h.y, a.nx, a.ny, a.world) == a.rank) {
                const size_t cell = (size_t)h.y * a.nx + h.x;
                if (!FILL) {
                    ++nhits;
                    if (atomicCAS(&a.slotmap[cell], kNone, kClaimed) == kNone) {
                        const unsigned slot = atomicAdd(&a.small[4], 1u);
                        a.slot_cell[slot] = (unsigned)cell;
                        a.slotmap[cell] = slot;                                  // read by later kernels only
                        a.tile_flag[tile] = 1u;
                    }
                } else {
                    const unsigned slot = a.slotmap[cell];
                    if (CHECK) hit_old = atomicAdd(&a.ord[(size_t)slot * a.ord_stride + sl], kHitUnit);   // looked at after the walk
                    else atomicAdd(&a.ord[(size_t)slot * a.ord_stride + sl], kHitUnit);
                }
            }
        }
    }
    // All lanes walk their rays' tile runs in lock step; runs of the warp that fall into the same
    // (tile, length class) take one atomic and consecutive slots.  Inside a tile the runs are laid
    // out by length class, so the 32 runs of a tile-kernel task have similar lengths.  The fill
    // pass stores a run one iteration late: its slot comes back from the atomic while the next
    // tile crossing is being computed.
    bool p_valid = false;
    unsigned p_base = 0, p_off = 0, p_rank = 0;
    int p_leader = 0;
    uint4 p_run = make_uint4(0u, 0u, 0u, 0u);
    for (;;) {
        TileRun t;
        const bool has = it.next(t);
        const bool more = __any_sync(0xffffffffu, has);
        const bool owned = has;                                // the walk is already clipped to the rank's bands
        const int seg = owned ? t.tile * kLenClasses + ((t.len - 1) >> kLenShift) : -1 - lane;
        unsigned base = 0, off = 0, peers = 0;
        int leader = 0;
        if (more) {
            peers = __match_any_sync(0xffffffffu, seg);
            leader = __ffs(peers) - 1;
            if (owned && lane == leader) base = atomicAdd(&a.tile_count[seg], (unsigned)__popc(peers));
            if (FILL && owned) off = a.tile_off[seg];
        }
        if (FILL) {
            const unsigned b = __shfl_sync(0xffffffffu, p_base, p_leader);
            if (p_valid) a.runs[p_off + b + p_rank] = p_run;
            p_valid = owned; p_base = base; p_leader = leader; p_off = off; p_rank = (unsigned)__popc(peers & lt_mask);
            if (owned) {
                int x, y;
                cell_at(it.g, t.n0, t.j0, x, y);
                p_run.x = (unsigned)((y % TS) * TS + (x % TS)) | (it.g.xmajor ? kRunXMajor : 0) |
                          (it.g.smaj > 0 ? kRunMajPos : 0) | (it.g.smin > 0 ? kRunMinPos : 0) |
                          ((unsigned)(t.len - 1) << 15) | ((unsigned)sl << 21);
                // RunWalker::start: (2n+2)*dmin - 2*dmaj*j - dmaj, bounded by 2*dmaj + 2*dmin
                p_run.y = (unsigned)(int)((2ll * t.n0 + 2) * it.g.dmin - 2ll * it.g.dmaj * t.j0 - it.g.dmaj);
                p_run.z = (unsigned)it.g.dmaj;
                p_run.w = (unsigned)it.g.dmin;
            }
        } else if (owned) {
            cells += t.len;
            ++nruns;
        }
        if (!more) break;
    }
    if (FILL && CHECK && (hit_old >> 20) == 4095u) a.small[3] = 1u;
    if (!FILL) {
        __shared__ unsigned long long part[3][8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cells += __shfl_xor_sync(0xffffffffu, cells, o);
            nhits += __shfl_xor_sync(0xffffffffu, nhits, o);
            nruns += __shfl_xor_sync(0xffffffffu, nruns, o);
        }
        const int w = threadIdx.x >> 5;
        if (lane == 0) { part[0][w] = cells; part[1][w] = nhits; part[2][w] = nruns; }
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

// ---- 2. scan of the per-tile run counts + active tiles, heaviest first ---------------
// Work items: a tile's runs are cut into pieces of at most kItemRuns, so that the busiest
// tiles (tens of thousands of runs) do not set the length of the tile kernel.  Tiles with
// several items combine their partial counts in a global per-cell counter (`multi` lists them).
constexpr unsigned kItemRuns = 2048;         // one GPU; a rank of a sharded grid holds 1 / world of the tiles and cuts them finer (item_runs)
__host__ __device__ inline unsigned occ_item_runs(int world) { return world >= 8 ? 512u : world >= 2 ? 1024u : kItemRuns; }

__global__ void __launch_bounds__(1024) occ_tile_scan(const unsigned* __restrict__ counts, int n_tiles,
                                                      unsigned* __restrict__ offsets /* n_tiles + 1 */,
                                                      unsigned* __restrict__ class_off /* [tile][class] */,
                                                      const unsigned* __restrict__ tile_flag, uint2* __restrict__ items, int* __restrict__ multi,
                                                      unsigned* __restrict__ small, unsigned char* __restrict__ dirty, unsigned item_runs) {
    __shared__ unsigned wsum[32];
    __shared__ unsigned carry;
    __shared__ int hist[2][33], start[2][33];      // [0] tiles with hit cells, [1] the others
    __shared__ unsigned n_multi;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry = 0; n_multi = 0; }
    if (tid < 66) hist[tid / 33][tid % 33] = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n_tiles; b0 += 1024) {
        const int t = b0 + tid;
        uint4 cc = make_uint4(0u, 0u, 0u, 0u), cd = cc;
        if (t < n_tiles) { cc = reinterpret_cast<const uint4*>(counts)[2 * t]; cd = reinterpret_cast<const uint4*>(counts)[2 * t + 1]; }
        const unsigned v = cc.x + cc.y + cc.z + cc.w + cd.x + cd.y + cd.z + cd.w;
        if (t < n_tiles && (v || tile_flag[t])) dirty[t] = 3;       // bit 0: touched since the last reset (read-out); bit 1: since the last push to the peers
        if (v) atomicAdd(&hist[tile_flag[t] ? 0 : 1][32 - __clz(v)], (int)((v + item_runs - 1) / item_runs));
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        unsigned base = carry;
        for (int w = 0; w < warp; ++w) base += wsum[w];
        if (t < n_tiles) {
            const unsigned o = base + inc - v;
            offsets[t] = o;
            const unsigned o4 = o + cc.x + cc.y + cc.z + cc.w;
            reinterpret_cast<uint4*>(class_off)[2 * t] = make_uint4(o, o + cc.x, o + cc.x + cc.y, o + cc.x + cc.y + cc.z);
            reinterpret_cast<uint4*>(class_off)[2 * t + 1] = make_uint4(o4, o4 + cd.x, o4 + cd.x + cd.y, o4 + cd.x + cd.y + cd.z);
        }
        __syncthreads();
        if (tid == 1023) carry = base + inc;
        __syncthreads();
    }
    if (tid == 0) {
        offsets[n_tiles] = carry;
        small[0] = carry;
        int run = 0;
        for (int g = 0; g < 2; ++g) {
            for (int b = 32; b >= 1; --b) { start[g][b] = run; run += hist[g][b]; }   // heaviest tiles first
            if (g == 0) small[6] = (unsigned)run;
        }
        small[1] = (unsigned)run;                                               // number of items
    }
    __syncthreads();
    for (int t = tid; t < n_tiles; t += 1024) {
        const uint4 cc = reinterpret_cast<const uint4*>(counts)[2 * t], cd = reinterpret_cast<const uint4*>(counts)[2 * t + 1];
        const unsigned v = cc.x + cc.y + cc.z + cc.w + cd.x + cd.y + cd.z + cd.w;
        if (v) {
            const int n_it = (int)((v + item_runs - 1) / item_runs);
            const int base = atomicAdd(&start[tile_flag[t] ? 0 : 1][32 - __clz(v)], n_it);
            for (int i = 0; i < n_it; ++i) items[base + i] = make_uint2((unsigned)t, (unsigned)i);
            if (n_it > 1) multi[atomicAdd(&n_multi, 1u)] = t;
        }
    }
    __syncthreads();
    if (tid == 0) small[5] = n_multi;
}

// ---- the add chain (mapping.py:129, 139, 141) ------------------------------------------
__device__ __forceinline__ float fast_chain(float x, unsigned m, unsigned k, double l_hit, double l_miss,
                                            float lo, float hi) {
    // Nothing is clipped inside a scan, so the value before the clip is x + m*l_hit + k*l_miss up to
    // the float rounding of each add (<= 2^-24 of the running magnitude, itself <= mag).  When that
    // cannot bring it back inside the clamp interval the clip alone decides -- the steady state of
    // free space (pinned low, misses only) and of walls (pinned high, hits outweigh the misses).
    {
        const double dm = (double)m, dk = (double)k;
        const double t = (double)x + dm * l_hit + dk * l_miss;
        const double mag = fabs((double)x) + dm * fabs(l_hit) + dk * fabs(l_miss);
        const double err = (dm + dk) * mag * 1.2e-7;
        if (t - err >= (double)hi) return hi;
        if (t + err <= (double)lo) return lo;
    }
    for (unsigned i = 0; i < m; ++i) {        // mapping.py:129 -- m hits, each x = f32(f64(x) + l_hit)
        const float y = (float)((double)x + l_hit);
        if (y == x) break;                    // absorbed: every further add is too
        x = y;
        if (k == 0 && ((l_hit > 0.0 && x >= hi) || (l_hit < 0.0 && x <= lo))) break;   // the clip decides
    }
    // mapping.py:139 -- k misses.  Each add moves x by l_miss up to one float rounding (< 1e-6 for
    // |x| <= 16), so k adds certainly cross the bound when k * (|l_miss| - 1e-6) covers the gap.
    if (k > 4 && fabsf(x) <= 16.f && lo >= -16.f && hi <= 16.f) {
        if (l_miss < -1e-5 && (double)x + (double)k * (l_miss + 1e-6) <= (double)lo) return lo;
        if (l_miss > 1e-5 && (double)x + (double)k * (l_miss - 1e-6) >= (double)hi) return hi;
    }
    for (unsigned i = 0; i < k; ++i) {
        const float y = (float)((double)x + l_miss);
        if (y == x) break;
        x = y;
        if ((l_miss < 0.0 && x <= lo) || (l_miss > 0.0 && x >= hi)) break;
    }
    return fminf(fmaxf(x, lo), hi);           // mapping.py:141
}

// The same chain written literally, for the ordered replay where m + k is a handful: no early
// exits to evaluate, three dependent instructions per add.
__device__ __forceinline__ float lean_chain(float x, unsigned m, unsigned k, double l_hit, double l_miss,
                                            float lo, float hi) {
    if (m + k > 8u) return fast_chain(x, m, k, l_hit, l_miss, lo, hi);
    for (unsigned i = 0; i < m; ++i) x = (float)((double)x + l_hit);      // mapping.py:129
    for (unsigned i = 0; i < k; ++i) x = (float)((double)x + l_miss);     // mapping.py:139
    return fminf(fmaxf(x, lo), hi);                                       // mapping.py:141
}

// ---- 4. tiles ----------------------------------------------------------------------------
struct TileArgs {
    float* grid;
    int nx, ny, tiles_x;
    const unsigned* tile_off;
    const uint4* runs;
    const uint2* items;                       // (tile, piece); tiles with hit cells first
    unsigned n_hit_items, item_first, item_end, item_runs;
    int persistent;
    const int* multi;                         // tiles cut into several items
    unsigned* small;                          // [1] n_items [2] queue [5] n_multi
    unsigned* ncount;                         // ny * nx partial miss counts of multi-item tiles (kept zero)
    const unsigned* slotmap;
    unsigned* ord;
    int ord_stride;
    double l_hit, l_miss;
    float lo, hi;
};

// shared-memory accesses by 32-bit shared address (no generic -> shared conversion per step)
__device__ __forceinline__ void red_shared_inc(unsigned addr) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ unsigned ld_shared_u32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int imad(int a, int b, int c) {      // a * b + c on the FMA pipe
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// Every lane walks one run; a task is 32 consecutive runs of the item.  The counters live in a
// padded layout (row stride TS + 1 words): lanes of a fan sit in the same column at different
// rows, which would all be one bank; the cell index is kept in padded space, so the padding
// costs nothing per step.  The Bresenham error term is kept negated (e = -d): "the minor axis
// steps too" is then its sign bit and both updates are multiply-adds by that bit (FMA pipe)
// instead of compare + select chains (the ALU pipe is the busy one).  A lane whose run is over
// keeps stepping a stale index but counts into a private dummy word: one select, no branch.
constexpr int TSP = TS + 1;
constexpr int kPadCells = TS * TSP;
constexpr int kStepsPerRound = 4;

__device__ __forceinline__ int pad_cell(int idx) { return idx + (idx >> 6); }

template <bool HITS>
__device__ __forceinline__ void walk_runs(const TileArgs& a, unsigned beg, unsigned end, unsigned cnt_base,
                                          unsigned slot_base, unsigned dummy_addr, int warp, int lane) {
    const unsigned full = 0xffffffffu;
    unsigned e0 = beg + (unsigned)warp * 32u;
    if (e0 >= end) return;
    uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
    if (e0 + lane < end) nxt = __ldg(a.runs + e0 + lane);
    for (; e0 < end; e0 += kTileNT) {
        const uint4 w = nxt;
        const bool have = e0 + lane < end;
        if (e0 + kTileNT + lane < end) nxt = __ldg(a.runs + e0 + kTileNT + lane);       // next task's record, in flight
        int idx = pad_cell((int)(w.x & 4095u));
        int rem = have ? (int)((w.x >> 15) & 63u) + 1 : 0;
        const unsigned sl = w.x >> 21;
        const int step_maj = (w.x & kRunXMajor) ? ((w.x & kRunMajPos) ? 1 : -1) : ((w.x & kRunMajPos) ? TSP : -TSP);
        const int nmnr = (w.x & kRunXMajor) ? ((w.x & kRunMinPos) ? -TSP : TSP) : ((w.x & kRunMinPos) ? -1 : 1);
        int e = -(int)w.y;
        const int ndec = -2 * (int)w.z, ninc = -2 * (int)w.w;
        const int maxlen = (int)__reduce_max_sync(full, (unsigned)rem);
        for (int k = 0; k < maxlen; k += kStepsPerRound) {
#pragma unroll
            for (int u = 0; u < kStepsPerRound; ++u) {
                const bool act = rem > 0;
                bool count_here = act;
                if (HITS) {
                    const unsigned sidx = ld_shared_u32(act ? (unsigned)imad(idx, 4, (int)slot_base) : slot_base);
                    const bool on_hit = act && sidx != kNone;
                    if (__any_sync(full, on_hit)) {
                        if (on_hit) atomicAdd(&a.ord[(size_t)sidx * a.ord_stride + sl], 1u);
                        count_here = act && !on_hit;
                    }
                }
                red_shared_inc(count_here ? (unsigned)imad(idx, 4, (int)cnt_base) : dummy_addr);
                const int mneg = e >> 31;                                        // -1 iff the minor axis steps too
                idx = imad(mneg, nmnr, idx + step_maj);
                e = imad(mneg, ndec, e + ninc);
                --rem;
            }
        }
    }
}

__global__ void __launch_bounds__(kTileNT) occ_fast_tiles(const TileArgs a) {
    __shared__ unsigned cnt[kPadCells];
    __shared__ unsigned slot[kPadCells];
    __shared__ unsigned dummy[kTileNT];           // where idle lanes count
    __shared__ uint2 cur_item;
    __shared__ unsigned cur_q;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Two launch shapes over the item range [item_first, item_end): persistent CTAs pulling from a
    // queue (the hit tiles, heaviest first), or one item per CTA (the rest: CTAs retire all the
    // time, so the hit cells' replay on the second stream finds SM slots while this runs).
    for (bool first = true;; first = false) {
        __syncthreads();
        if (tid == 0) {
            unsigned q;
            if (a.persistent) q = a.item_first + atomicAdd(&a.small[2], 1u);
            else q = first ? a.item_first + blockIdx.x : a.item_end;
            cur_q = q;
            cur_item = q < a.item_end ? a.items[q] : make_uint2(kNone, 0u);
        }
        __syncthreads();
        if (cur_item.x == kNone) break;
        const int t = (int)cur_item.x;
        const bool HITS = cur_q < a.n_hit_items;          // items of tiles that hold hit cells come first
        const int tx0 = (t % a.tiles_x) * TS, ty0 = (t / a.tiles_x) * TS;
        for (int c = tid; c < TCELLS; c += kTileNT) {
            if (HITS) {
                const int x = tx0 + (c & (TS - 1)), y = ty0 + (c / TS);
                slot[pad_cell(c)] = (x < a.nx && y < a.ny) ? a.slotmap[(size_t)y * a.nx + x] : kNone;
            }
            cnt[pad_cell(c)] = 0u;
        }
        __syncthreads();
        const unsigned t_beg = a.tile_off[t], t_end = a.tile_off[t + 1];
        const unsigned beg = t_beg + cur_item.y * a.item_runs, end = min(beg + a.item_runs, t_end);
        unsigned cnt_base = (unsigned)__cvta_generic_to_shared(cnt), slot_base = (unsigned)__cvta_generic_to_shared(slot);
        unsigned dummy_addr = (unsigned)__cvta_generic_to_shared(&dummy[tid]);
        asm volatile("" : "+r"(cnt_base), "+r"(slot_base), "+r"(dummy_addr));     // keep them in registers: no re-derivation per step
        if (HITS) walk_runs<true>(a, beg, end, cnt_base, slot_base, dummy_addr, warp, lane);
        else      walk_runs<false>(a, beg, end, cnt_base, slot_base, dummy_addr, warp, lane);
        __syncthreads();
        const bool whole = t_end - t_beg <= a.item_runs;
        for (int c = tid; c < TCELLS; c += kTileNT) {
            const unsigned n = cnt[pad_cell(c)];
            if (n) {
                const int x = tx0 + (c & (TS - 1)), y = ty0 + (c / TS);
                const size_t cell = (size_t)y * a.nx + x;
                if (whole) a.grid[cell] = fast_chain(a.grid[cell], 0u, n, a.l_hit, a.l_miss, a.lo, a.hi);
                else atomicAdd(&a.ncount[cell], n);
            }
        }
    }
}

// tiles that were cut into several items: total misses per cell -> one chain (a CTA per quarter tile)
__global__ void __launch_bounds__(256) occ_fast_apply_multi(const TileArgs a) {
    const int t = a.multi[blockIdx.x >> 2];
    const int tx0 = (t % a.tiles_x) * TS, ty0 = (t / a.tiles_x) * TS + (blockIdx.x & 3) * (TS / 4);
    unsigned n[4];
    size_t cell[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int c = threadIdx.x + u * 256;
        const int x = tx0 + (c & (TS - 1)), y = ty0 + (c / TS);
        cell[u] = (size_t)y * a.nx + x;
        n[u] = (x < a.nx && y < a.ny) ? a.ncount[cell[u]] : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (n[u]) {
            a.ncount[cell[u]] = 0u;
            a.grid[cell[u]] = fast_chain(a.grid[cell[u]], 0u, n[u], a.l_hit, a.l_miss, a.lo, a.hi);
        }
}

// ---- 5. ordered replay of the hit cells ---------------------------------------------------
// The rows of `ord` are sparse (a wall cell is seen by ~7 % of the scans), so a lane that
// walked its own row would idle most of the time, and a warp that walked one row would run
// the chain 32 times redundantly.  Two kernels instead:
//   occ_fast_compact  one warp per hit cell, coalesced: the non-zero words of its row are
//                     moved, in scan order, to the front of the matching row of `ev`
//                     (and cleared in `ord`, which is thereby left clean for the next chunk)
//   occ_fast_chain    one lane per hit cell: runs the add chain over its event list --
//                     in scan order, which is all the reference's semantics need.
__global__ void __launch_bounds__(256) occ_fast_compact(unsigned* __restrict__ ord, unsigned* __restrict__ ev,
                                                        unsigned* __restrict__ ev_count, const unsigned* __restrict__ small,
                                                        int chunk_scans, int stride) {
    const unsigned n_slots = small[4];
    const unsigned slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (slot >= n_slots) return;
    unsigned* row = ord + (size_t)slot * stride;
    unsigned* out = ev + (size_t)slot * stride;
    const unsigned lt = (1u << lane) - 1u;
    int pos = 0;
    for (int s0 = 0; s0 < chunk_scans; s0 += 256) {
        unsigned v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int s = s0 + u * 32 + lane;
            v[u] = s < chunk_scans ? __ldcs(row + s) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned nz = __ballot_sync(0xffffffffu, v[u] != 0u);
            if (v[u]) {
                row[s0 + u * 32 + lane] = 0u;
                out[pos + __popc(nz & lt)] = v[u];
            }
            pos += __popc(nz);
        }
    }
    if (lane == 0) ev_count[slot] = (unsigned)pos;
}

// A hit cell that sits on a clamp bound usually stays there: walls are pinned high, each event adds more hits than
// misses.  For an event of at most 8 adds starting AT a bound, the literal chain ends at x + m*l_hit + k*l_miss up to
// the float rounding of each add (<= 8 * 2^-24 * (|bound| + 8 * max|l|)), so when that sum leaves the interval by more
// than `margin` (4e-6 * (|bound| + 8 * max|l|), which also covers the float evaluation of the sum below) the clip
// returns the bound again and the three-instruction-deep fp64 adds are skipped.  The longest event lists belong to
// exactly such cells, and the longest list is the critical path of the whole replay.
__device__ __forceinline__ float chain_event(float x, unsigned w, double l_hit, double l_miss, float lhf, float lmf,
                                             float margin, float lo, float hi) {
    const unsigned m = w >> 20, k = w & kMissMask;
    if (m + k <= 8u) {
        const float t = fmaf((float)m, lhf, (float)k * lmf);
        if ((x == hi && t > margin) || (x == lo && t < -margin)) return x;
    }
    return lean_chain(x, m, k, l_hit, l_miss, lo, hi);
}

// One GROUP of kChainG lanes per hit cell.  The event list of a cell is cut into kChainG consecutive segments, one per
// lane.  Lane 0 runs its segment from the cell's value; every other lane does not know its input yet and runs its
// segment from BOTH ends of the clamp interval.  Every event is a monotone map of [lo, hi] into itself (a rounded add
// never reorders two values, neither does the clip), so a segment is one too: when its two trajectories meet, its output
// is that value whatever the input was -- and for wall and free cells they meet within a dozen events.  The group then
// takes the output of its last such segment and replays only the segments behind it (usually none, or empty ones) in
// order.  Exact (the literal adds are still what produces every value); the longest dependent chain of the replay drops
// from the longest list (373 events on C4) to an eighth of it plus the unresolved tail.

__device__ __forceinline__ float chain_segment(float x, const uint4* __restrict__ row, int gb, int ge, int n, double l_hit,
                                               double l_miss, float lhf, float lmf, float margin, float lo, float hi) {
    for (int q = gb; q < ge; ++q) {
        const uint4 w = __ldcs(row + q);
        const int p = 4 * q;
        x = chain_event(x, w.x, l_hit, l_miss, lhf, lmf, margin, lo, hi);
        if (p + 1 < n) x = chain_event(x, w.y, l_hit, l_miss, lhf, lmf, margin, lo, hi);
        if (p + 2 < n) x = chain_event(x, w.z, l_hit, l_miss, lhf, lmf, margin, lo, hi);
        if (p + 3 < n) x = chain_event(x, w.w, l_hit, l_miss, lhf, lmf, margin, lo, hi);
    }
    return x;
}

template <int kChainG>
__global__ void __launch_bounds__(128, 4) occ_fast_chain(float* __restrict__ grid, unsigned* __restrict__ slotmap,
                                                         const unsigned* __restrict__ slot_cell, const unsigned* __restrict__ ev,
                                                         const unsigned* __restrict__ ev_count, const unsigned* __restrict__ small,
                                                         int stride, double l_hit, double l_miss, float lo, float hi) {
    static_assert(128 % kChainG == 0 && 32 % kChainG == 0, "groups must not straddle warps");
    const unsigned full = 0xffffffffu;
    const unsigned n_slots = small[4];
    const unsigned slot = (blockIdx.x * blockDim.x + threadIdx.x) / kChainG;
    const int lane = threadIdx.x & 31, j = lane % kChainG, base = lane - j;
    const bool valid = slot < n_slots;
    const unsigned cell = valid ? slot_cell[slot] : 0u;
    const int n = valid ? (int)ev_count[slot] : 0;
    const float x0 = valid ? grid[cell] : 0.f;
    const float lhf = (float)l_hit, lmf = (float)l_miss;
    const float margin = 4e-6f * (fmaxf(fabsf(lo), fabsf(hi)) + 8.f * fmaxf(fabsf(lhf), fabsf(lmf))) * 1.0001f;
    const uint4* row = reinterpret_cast<const uint4*>(ev + (size_t)slot * stride);      // stride is a multiple of 4
    // segments in units of four events (one 16-byte load); short lists stay with lane 0
    const int quads = (n + 3) >> 2;
    const int per = quads <= 4 ? quads : (quads + kChainG - 1) / kChainG;
    const int gb = min(j * per, quads), ge = min(gb + per, quads);
    // x0 must lie inside the clamp interval for the two-ended argument (it does: the interval contains 0 on this path
    // and every earlier update ended with the clip); a value outside is replayed by lane 0 alone
    const bool solo = !(x0 >= lo && x0 <= hi);
    float a = j == 0 ? x0 : lo, b = j == 0 ? x0 : hi;
    const int my_gb = solo ? (j == 0 ? 0 : quads) : gb, my_ge = solo ? quads : ge;
    const int steps = (int)__reduce_max_sync(full, (unsigned)(my_ge - my_gb));
    uint4 w_next = my_gb < my_ge ? __ldcs(row + my_gb) : make_uint4(0u, 0u, 0u, 0u);
    for (int q = 0; q < steps; ++q) {
        const int g4 = my_gb + q;
        const bool act = g4 < my_ge;
        const uint4 w = w_next;
        if (g4 + 1 < my_ge) w_next = __ldcs(row + g4 + 1);                       // fetched while this group is replayed
        const int p = 4 * g4;
        const bool two = act && __float_as_uint(a) != __float_as_uint(b);
        const bool any_two = __any_sync(full, two);
        if (act) {
            a = chain_event(a, w.x, l_hit, l_miss, lhf, lmf, margin, lo, hi);
            if (p + 1 < n) a = chain_event(a, w.y, l_hit, l_miss, lhf, lmf, margin, lo, hi);
            if (p + 2 < n) a = chain_event(a, w.z, l_hit, l_miss, lhf, lmf, margin, lo, hi);
            if (p + 3 < n) a = chain_event(a, w.w, l_hit, l_miss, lhf, lmf, margin, lo, hi);
        }
        if (any_two) {
            if (two) {
                b = chain_event(b, w.x, l_hit, l_miss, lhf, lmf, margin, lo, hi);
                if (p + 1 < n) b = chain_event(b, w.y, l_hit, l_miss, lhf, lmf, margin, lo, hi);
                if (p + 2 < n) b = chain_event(b, w.z, l_hit, l_miss, lhf, lmf, margin, lo, hi);
                if (p + 3 < n) b = chain_event(b, w.w, l_hit, l_miss, lhf, lmf, margin, lo, hi);
            } else if (act) {
                b = a;
            }
        } else if (act) {
            b = a;
        }
    }
    // the last segment whose output does not depend on its input (lane 0 always qualifies: it knew its input)
    const bool met = __float_as_uint(a) == __float_as_uint(b);
    const unsigned met_mask = (__ballot_sync(full, met) >> base) & (kChainG == 32 ? 0xffffffffu : ((1u << (kChainG & 31)) - 1u));
    const int last = 31 - __clz((int)(met_mask | 1u));
    float x = __shfl_sync(full, a, base + last);
#pragma unroll 1
    for (int jj = 1; jj < kChainG; ++jj) {
        // segments behind `last` are replayed in order from the value that reaches them
        if (j == jj && jj > last && my_gb < my_ge)
            a = chain_segment(x, row, my_gb, my_ge, n, l_hit, l_miss, lhf, lmf, margin, lo, hi);
        const float y = __shfl_sync(full, a, base + jj);
        if (jj > last && __shfl_sync(full, (int)(my_gb < my_ge), base + jj)) x = y;
    }
    if (valid && j == 0) {
        grid[cell] = x;
        slotmap[cell] = kNone;
    }
}

// group size: 4 and 8 lanes per cell measure the same on C4 (696 us per update), 16 and 32 are slower (721 / 754 us: more
// lanes replay their segment twice); one lane per cell was 0.71 ms
constexpr int kChainLanes = 8;

static void launch_chain(unsigned n_slots, cudaStream_t st, float* grid, unsigned* slotmap, const unsigned* slot_cell,
                         const unsigned* ev, const unsigned* ev_count, const unsigned* small, int stride, double l_hit,
                         double l_miss, float lo, float hi) {
    const unsigned blocks = (unsigned)(((unsigned long long)n_slots * kChainLanes + 127ull) / 128ull);
    occ_fast_chain<kChainLanes><<<blocks, 128, 0, st>>>(grid, slotmap, slot_cell, ev, ev_count, small, stride, l_hit, l_miss, lo, hi);
}

}  // namespace

// Optional stage timing (ICPB200_OCC_TIMING=1): events on the main stream at the stage boundaries,
// printed to stderr after the chunk.  A profiling aid; never on in the timed runs.
struct StageTimer {
    bool on = false;
    cudaEvent_t ev[12];
    const char* name[12];
    int n = 0;
    cudaStream_t st = nullptr;
    void init(cudaStream_t s) {
        static const bool want = getenv("ICPB200_OCC_TIMING") != nullptr;
        on = want; st = s;
    }
    // events on the second stream, reported relative to the main stream's mark `ref` ("fill")
    cudaEvent_t aux_ev[4];
    const char* aux_name[4];
    int n_aux = 0, ref = 0;
    void mark_aux(const char* what, cudaStream_t aux) {
        if (!on || n_aux >= 4) return;
        cudaEventCreate(&aux_ev[n_aux]);
        cudaEventRecord(aux_ev[n_aux], aux);
        aux_name[n_aux++] = what;
    }
    void mark(const char* what) {
        if (!on || n >= 12) return;
        cudaEventCreate(&ev[n]);
        cudaEventRecord(ev[n], st);
        name[n++] = what;
    }
    void report() {
        if (!on) return;
        cudaStreamSynchronize(st);
        for (int i = 1; i < n; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
            fprintf(stderr, "[occ stage] %-12s %8.1f us\n", name[i], ms * 1e3f);
        }
        for (int i = 0; i < n_aux; ++i) {
            float ms = 0.f;
            cudaEventSynchronize(aux_ev[i]);
            cudaEventElapsedTime(&ms, ev[ref], aux_ev[i]);
            fprintf(stderr, "[occ stage]   second stream: %-12s done %8.1f us after %s\n", aux_name[i], ms * 1e3f, name[ref]);
            cudaEventDestroy(aux_ev[i]);
        }
        n_aux = 0;
        float tot = 0.f;
        if (n > 1) cudaEventElapsedTime(&tot, ev[0], ev[n - 1]);
        fprintf(stderr, "[occ stage] %-12s %8.1f us\n", "total", tot * 1e3f);
        for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
        n = 0;
    }
};

// Returns ICPB200_OK, an error, or 1 when this chunk must take the ordered path instead.
// [rb, re) are the chunk's rays; big_scan: 1 = some scan of the chunk has more than 4095 rays, 0 = none has,
// -1 = the host does not know (the device flag kFlagBigScan decides; comes back with the totals).
// *bad_offsets is set when the device-side check of the offsets failed (nothing was written then).
static int fast_chunk(OccGrid& g, int n_scans, int s0, int cs, const double* d_origins, const double* d_hits,
                      const long long* d_hit_off, long long rb, long long re, int big_scan_in, bool* bad_offsets,
                      cudaStream_t st) {
    const long long nr = re - rb;
    if (nr == 0) return ICPB200_OK;
    const int tiles_x = (g.nx + TS - 1) / TS, tiles_y = (g.ny + TS - 1) / TS;
    const int n_tiles = tiles_x * tiles_y;
    // a ray crosses at most tiles_x + tiles_y + 1 tiles, which bounds the number of work items
    const size_t max_items = (size_t)n_tiles + (size_t)(((unsigned long long)nr * (tiles_x + tiles_y + 1)) / occ_item_runs(g.world)) + 1;
    if (g.tile_count.reserve(sizeof(unsigned) * kLenClasses * (size_t)n_tiles) || g.offsets.reserve(sizeof(unsigned) * ((size_t)n_tiles + 1)) ||
        g.class_off.reserve(sizeof(unsigned) * kLenClasses * (size_t)n_tiles) ||
        g.slot_cell.reserve(sizeof(unsigned) * (size_t)nr) || g.items.reserve(sizeof(uint2) * max_items) ||
        g.multi.reserve(sizeof(int) * (size_t)n_tiles) || g.tile_flag.reserve(sizeof(unsigned) * (size_t)n_tiles))
        return ICPB200_ERR_CUDA;
    if (!g.aux_stream) {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        ICPB_CUDA(cudaStreamCreateWithPriority(&g.aux_stream, cudaStreamNonBlocking, prio_hi));
        ICPB_CUDA(cudaEventCreateWithFlags(&g.ev_fork, cudaEventDisableTiming));
        ICPB_CUDA(cudaEventCreateWithFlags(&g.ev_join, cudaEventDisableTiming));
        ICPB_CUDA(cudaEventCreateWithFlags(&g.ev_hit, cudaEventDisableTiming));
    }
    unsigned* d_small = g.small.as<unsigned>();
    unsigned long long* d_stats = reinterpret_cast<unsigned long long*>(g.small.as<unsigned char>() + 64);
    StageTimer tm;
    tm.init(st);
    tm.mark("begin");
    ICPB_CUDA(cudaMemsetAsync(d_small, 0, 64, st));
    ICPB_CUDA(cudaMemsetAsync(g.tile_count.p, 0, sizeof(unsigned) * kLenClasses * (size_t)n_tiles, st));
    ICPB_CUDA(cudaMemsetAsync(g.tile_flag.p, 0, sizeof(unsigned) * (size_t)n_tiles, st));

    FastArgs a;
    a.hits = reinterpret_cast<const double2*>(d_hits);
    a.hit_off = d_hit_off; a.origins = d_origins;
    a.ray_begin = rb; a.ray_end = re;
    a.scan_begin = s0; a.chunk_scans = cs; a.n_scans = n_scans;
    a.min_x = g.min_x; a.min_y = g.min_y; a.res = g.res;
    a.nx = g.nx; a.ny = g.ny; a.tiles_x = tiles_x; a.n_tiles = n_tiles;
    a.rank = g.rank; a.world = g.world;
    a.ray_cell = g.ray_cell.as<int2>();
    a.ray_scan = g.ray_scan.as<int>();
    a.origin_cell = g.origin_cell.as<int2>();
    a.slotmap = g.slotmap.as<unsigned>();
    a.slot_cell = g.slot_cell.as<unsigned>();
    a.ord = nullptr; a.ord_stride = 0;
    a.tile_count = g.tile_count.as<unsigned>();
    a.tile_flag = g.tile_flag.as<unsigned>();
    a.tile_off = nullptr; a.runs = nullptr;
    a.runs_cap = 0; a.ord_cap = 0;
    a.small = d_small; a.stats = d_stats;
    const unsigned nblk = (unsigned)((nr + 255) / 256);
    const bool sharded = g.world > 1;
    if (sharded) occ_fast_rays<false, false, true><<<nblk, 256, 0, st>>>(a);
    else occ_fast_rays<false, false, false><<<nblk, 256, 0, st>>>(a);
    ICPB_LAUNCH_CHECK();
    tm.mark("count");
    occ_tile_scan<<<1, 1024, 0, st>>>(g.tile_count.as<unsigned>(), n_tiles, g.offsets.as<unsigned>(), g.class_off.as<unsigned>(), g.tile_flag.as<unsigned>(), g.items.as<uint2>(), g.multi.as<int>(), d_small, g.dirty.as<unsigned char>(), occ_item_runs(g.world));
    ICPB_LAUNCH_CHECK();
    unsigned h_small[40];
    tm.mark("scan");
    const int stride = (cs + 3) & ~3;
    bool big_scan = big_scan_in > 0;
    auto launch_fill = [&]() -> int {
        ICPB_CUDA(cudaMemsetAsync(g.tile_count.p, 0, sizeof(unsigned) * kLenClasses * (size_t)n_tiles, st));
        a.ord = g.ord.as<unsigned>(); a.ord_stride = stride;
        a.tile_off = g.class_off.as<unsigned>();
        a.runs = g.runs.as<uint4>();
        a.runs_cap = g.runs.cap / sizeof(uint4);
        a.ord_cap = g.ord.cap / sizeof(unsigned);
        if (sharded) {
            if (big_scan) occ_fast_rays<true, true, true><<<nblk, 256, 0, st>>>(a);
            else occ_fast_rays<true, false, true><<<nblk, 256, 0, st>>>(a);
        } else {
            if (big_scan) occ_fast_rays<true, true, false><<<nblk, 256, 0, st>>>(a);
            else occ_fast_rays<true, false, false><<<nblk, 256, 0, st>>>(a);
        }
        ICPB_LAUNCH_CHECK();
        return ICPB200_OK;
    };
    // The fill pass needs buffers sized by totals the host does not know yet.  From the second call on the buffers of
    // the previous call are almost always large enough: launch the fill first (it checks the totals on the device and
    // leaves without a write if they do not fit), read the totals back underneath it, repeat it only if it had to leave.
    const bool speculative = g.runs.cap > 0 && g.ord.cap > 0;
    if (speculative) {
        int rc = launch_fill();
        if (rc) return rc;
    }
    ICPB_CUDA(cudaMemcpyAsync(h_small, d_small, sizeof(h_small), cudaMemcpyDeviceToHost, st));
    ICPB_CUDA(cudaStreamSynchronize(st));
    if (h_small[kFlagBadOffsets]) { *bad_offsets = true; return ICPB200_OK; }
    const bool wrong_variant = big_scan_in < 0 && h_small[kFlagBigScan] != 0;      // the speculative fill left: it needs CHECK
    if (wrong_variant) big_scan = true;
    const unsigned total_runs = h_small[0], n_slots = h_small[4];
    const size_t ord_words = (size_t)n_slots * stride;
    const bool filled = speculative && !wrong_variant && (size_t)total_runs + 64 <= g.runs.cap / sizeof(uint4) &&
                        ord_words <= g.ord.cap / sizeof(unsigned);
    if (ord_words > kOccOrdBudget) {
        // too many hit cells for the dense table: undo the claims, the ordered path takes the chunk
        if (n_slots) {
            if (g.ev_count.reserve(sizeof(unsigned) * (size_t)n_slots)) return ICPB200_ERR_CUDA;
            ICPB_CUDA(cudaMemsetAsync(g.ev_count.p, 0, sizeof(unsigned) * (size_t)n_slots, st));
            launch_chain(n_slots, st, g.grid.as<float>(), g.slotmap.as<unsigned>(),
                                                                    g.slot_cell.as<unsigned>(), nullptr, g.ev_count.as<unsigned>(), d_small,
                                                                    0, g.l_hit, g.l_miss, (float)g.lo_min, (float)g.lo_max);
            ICPB_LAUNCH_CHECK();
        }
        return 1;
    }
    if (!filled) {
        void* before = g.ord.p;
        const size_t cap_before = g.ord.cap;
        if (g.ord.reserve(sizeof(unsigned) * std::max<size_t>(ord_words, 1))) return ICPB200_ERR_CUDA;
        if (g.ord.p != before || g.ord.cap != cap_before) ICPB_CUDA(cudaMemsetAsync(g.ord.p, 0, g.ord.cap, st));
        if (g.runs.reserve(sizeof(uint4) * ((size_t)total_runs + 64))) return ICPB200_ERR_CUDA;
    }
    if (g.ev.reserve(std::max<size_t>(g.ord.cap, sizeof(unsigned))) || g.ev_count.reserve(sizeof(unsigned) * std::max<size_t>(n_slots, 1)))
        return ICPB200_ERR_CUDA;
    tm.mark("host gap");
    if (!filled) {
        int rc = launch_fill();
        if (rc) return rc;
    }
    tm.mark("fill");
    tm.ref = tm.n - 1;
    const float lo = (float)g.lo_min, hi = (float)g.lo_max;
    bool replayed = false;
    if (total_runs) {
        TileArgs t;
        t.grid = g.grid.as<float>();
        t.nx = g.nx; t.ny = g.ny; t.tiles_x = tiles_x;
        t.tile_off = g.offsets.as<unsigned>();
        t.runs = g.runs.as<uint4>();
        t.items = g.items.as<uint2>();
        t.multi = g.multi.as<int>();
        t.small = d_small;
        t.ncount = g.ncount.as<unsigned>();
        t.slotmap = g.slotmap.as<unsigned>();
        t.ord = g.ord.as<unsigned>();
        t.ord_stride = stride;
        t.l_hit = g.l_hit; t.l_miss = g.l_miss; t.lo = lo; t.hi = hi;
        const unsigned n_items = h_small[1], n_multi = h_small[5];
        t.n_hit_items = h_small[6];
        t.item_runs = occ_item_runs(g.world);
        // Two launches side by side.  The tiles that hold hit cells go to the priority stream (persistent CTAs pulling
        // from a queue, heaviest first) followed by the hit cells' replay; the other tiles go to the main stream, one
        // item per CTA, and fill every SM slot the first launch leaves free -- its tail and the replay run underneath.
        bool forked = false;
        if (t.n_hit_items || n_slots) {
            ICPB_CUDA(cudaEventRecord(g.ev_fork, st));
            ICPB_CUDA(cudaStreamWaitEvent(g.aux_stream, g.ev_fork, 0));
            forked = true;
        }
        if (t.n_hit_items) {
            t.item_first = 0u; t.item_end = t.n_hit_items; t.persistent = 1;
            occ_fast_tiles<<<std::min<unsigned>(t.n_hit_items, (unsigned)g.fast_ctas), kTileNT, 0, g.aux_stream>>>(t);
            ICPB_LAUNCH_CHECK();
            ICPB_CUDA(cudaEventRecord(g.ev_hit, g.aux_stream));
            tm.mark_aux("hit tiles", g.aux_stream);
        }
        if (n_slots) {
            // the hit cells' tables are complete once the hit tiles are done
            occ_fast_compact<<<(n_slots * 32u + 255u) / 256u, 256, 0, g.aux_stream>>>(g.ord.as<unsigned>(), g.ev.as<unsigned>(),
                                                                                      g.ev_count.as<unsigned>(), d_small, cs, stride);
            ICPB_LAUNCH_CHECK();
            tm.mark_aux("compact", g.aux_stream);
            launch_chain(n_slots, g.aux_stream, g.grid.as<float>(), g.slotmap.as<unsigned>(),
                                                                             g.slot_cell.as<unsigned>(), g.ev.as<unsigned>(),
                                                                             g.ev_count.as<unsigned>(), d_small, stride, g.l_hit, g.l_miss, lo, hi);
            ICPB_LAUNCH_CHECK();
            tm.mark_aux("chain", g.aux_stream);
            replayed = true;
        }
        if (forked) ICPB_CUDA(cudaEventRecord(g.ev_join, g.aux_stream));
        if (n_items > t.n_hit_items) {
            t.item_first = t.n_hit_items; t.item_end = n_items; t.persistent = 0;
            occ_fast_tiles<<<n_items - t.n_hit_items, kTileNT, 0, st>>>(t);
            ICPB_LAUNCH_CHECK();
        }
        tm.mark("tiles rest");
        if (n_multi) {
            if (t.n_hit_items) ICPB_CUDA(cudaStreamWaitEvent(st, g.ev_hit, 0));       // pieces of a tile may sit in both launches
            occ_fast_apply_multi<<<n_multi * 4u, 256, 0, st>>>(t);
            ICPB_LAUNCH_CHECK();
        }
        tm.mark("apply multi");
        if (forked) ICPB_CUDA(cudaStreamWaitEvent(st, g.ev_join, 0));
        tm.mark("join replay");
    }
    if (n_slots && !replayed) {                       // no tile runs at all: hits only
        occ_fast_compact<<<(n_slots * 32u + 255u) / 256u, 256, 0, st>>>(g.ord.as<unsigned>(), g.ev.as<unsigned>(),
                                                                        g.ev_count.as<unsigned>(), d_small, cs, stride);
        ICPB_LAUNCH_CHECK();
        launch_chain(n_slots, st, g.grid.as<float>(), g.slotmap.as<unsigned>(),
                                                                g.slot_cell.as<unsigned>(), g.ev.as<unsigned>(),
                                                                g.ev_count.as<unsigned>(), d_small, stride, g.l_hit, g.l_miss, lo, hi);
        ICPB_LAUNCH_CHECK();
    }
    tm.mark("end");
    tm.report();
    return ICPB200_OK;
}

// ---- read-out of the touched tiles only ---------------------------------------------------------------------------
// A map is mostly unexplored (C4: 510 of the 4096 tiles carry anything): the tiles touched since the last reset are packed
// into one block and only that block crosses PCIe.  view: 0 log-odds, 1 probability, 2 display value (mapping.py:150-160),
// evaluated in float32 throughout as numpy does; expf differs from numpy's float32 exp by at most 2 ulp.
__device__ __forceinline__ float occ_view(float lo, int view) {
    if (view == 0) return lo;
    const float p = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-lo)));              // mapping.py:153
    if (view == 1) return p;
    return lo == 0.0f ? 1.0f : (lo < 0.0f ? 0.85f : __fsub_rn(1.0f, p));       // mapping.py:157-159
}

__global__ void __launch_bounds__(256) occ_pack_tiles(const float* __restrict__ grid, int nx, int ny, int tiles_x,
                                                      const int* __restrict__ ids, float* __restrict__ pack, int view) {
    const int t = ids[blockIdx.x];
    const int tx0 = (t % tiles_x) * TS, ty0 = (t / tiles_x) * TS;
    float* out = pack + (size_t)blockIdx.x * TCELLS;
    for (int c = threadIdx.x; c < TCELLS; c += 256) {
        const int x = tx0 + (c & (TS - 1)), y = ty0 + c / TS;
        out[c] = (x < nx && y < ny) ? occ_view(grid[(size_t)y * nx + x], view) : 0.f;
    }
}

// The caller's array is page-locked (icpb200_pin_host): every touched tile is written straight to its place in host memory
// over PCIe, 256-byte row segments at a time -- no staging block, no list of tiles on the host, no scatter by the CPU.
__global__ void __launch_bounds__(256) occ_write_tiles_mapped(const float* __restrict__ grid, int nx, int ny, int tiles_x,
                                                              const unsigned char* __restrict__ dirty, float* __restrict__ host_out,
                                                              unsigned* __restrict__ count, int view) {
    const int t = blockIdx.x;
    if (!(dirty[t] & 1)) return;
    if (threadIdx.x == 0) atomicAdd(count, 1u);
    const int tx0 = (t % tiles_x) * TS, ty0 = (t / tiles_x) * TS;
    for (int c4 = threadIdx.x; c4 < TCELLS / 4; c4 += 256) {
        const int x = tx0 + (c4 % (TS / 4)) * 4, y = ty0 + c4 / (TS / 4);
        if (y >= ny || x >= nx) continue;
        const size_t at = (size_t)y * nx + x;
        if (x + 3 < nx && (at & 3) == 0) {
            const float4 v = *reinterpret_cast<const float4*>(grid + at);
            *reinterpret_cast<float4*>(host_out + at) = make_float4(occ_view(v.x, view), occ_view(v.y, view), occ_view(v.z, view), occ_view(v.w, view));
        } else {
            for (int u = 0; u < 4 && x + u < nx; ++u) host_out[at + u] = occ_view(grid[at + u], view);
        }
    }
}

// ---- multi-GPU: push the touched tiles this rank owns straight into every peer's grid (NVLink peer stores) ------------
// The map is sharded by bands of 64 rows (occ_owner); a tile is written by exactly one rank.  Instead of gathering whole
// bands (64 MiB through NCCL plus pack / unpack copies) every rank stores the tiles it has touched since its last push
// into the same place of every peer's grid, whose device pointers were exchanged once (CUDA IPC handles): the exchange
// moves the explored part of the map only, in one kernel, with no staging.  The peers' touched-tile flags are set as well,
// so their read-outs see the tiles.  A cross-rank barrier after the push is the caller's (icp_b200.dist.grid_push_device).
struct PeerTable {
    float* grid[16];
    unsigned char* dirty[16];
};

__global__ void __launch_bounds__(256) occ_push_tiles(const float* __restrict__ grid, int nx, int ny, int tiles_x, int rank, int world,
                                                      unsigned char* __restrict__ dirty, const PeerTable peers, int push_all,
                                                      unsigned* __restrict__ count) {
    const int t = blockIdx.x;
    const int ty0 = (t / tiles_x) * TS, tx0 = (t % tiles_x) * TS;
    if (occ_owner(tx0, ty0, nx, ny, world) != rank) return;
    const unsigned char flag = dirty[t];
    if (!push_all && !(flag & 2)) return;
    if (threadIdx.x == 0) atomicAdd(count, 1u);
    for (int c4 = threadIdx.x; c4 < TCELLS / 4; c4 += 256) {
        const int x = tx0 + (c4 % (TS / 4)) * 4, y = ty0 + c4 / (TS / 4);
        if (y >= ny || x >= nx) continue;
        const size_t at = (size_t)y * nx + x;
        if (x + 3 < nx && (at & 3) == 0) {
            const float4 v = *reinterpret_cast<const float4*>(grid + at);
            for (int p = 0; p < world; ++p)
                if (p != rank) *reinterpret_cast<float4*>(peers.grid[p] + at) = v;
        } else {
            for (int u = 0; u < 4 && x + u < nx; ++u)
                for (int p = 0; p < world; ++p)
                    if (p != rank) peers.grid[p][at + u] = grid[at + u];
        }
    }
    if (threadIdx.x == 0) {
        for (int p = 0; p < world; ++p)
            if (p != rank) peers.dirty[p][t] = (unsigned char)(peers.dirty[p][t] | 1);    // (only this rank ever writes tile t's flag on a peer)
        dirty[t] = (unsigned char)(flag & ~2);
    }
}

int occ_push_to_peers(OccGrid& g, cudaStream_t st) {
    if (g.world <= 1) return ICPB200_OK;
    if (!g.peers_attached) { set_error("icpb200_grid_push_tiles: icpb200_grid_ipc_attach has not been called"); return ICPB200_ERR_ARG; }
    const int tiles_x = (g.nx + TS - 1) / TS, tiles_y = (g.ny + TS - 1) / TS, n_tiles = tiles_x * tiles_y;
    if (g.small.reserve(256)) return ICPB200_ERR_CUDA;
    PeerTable pt;
    for (int p = 0; p < 16; ++p) { pt.grid[p] = static_cast<float*>(g.peer_grid[p]); pt.dirty[p] = static_cast<unsigned char*>(g.peer_dirty[p]); }
    unsigned* d_count = g.small.as<unsigned>() + 49;
    ICPB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned), st));
    occ_push_tiles<<<(unsigned)n_tiles, 256, 0, st>>>(g.grid.as<float>(), g.nx, g.ny, tiles_x, g.rank, g.world, g.dirty.as<unsigned char>(), pt,
                                                       g.all_dirty ? 1 : 0, d_count);
    ICPB_LAUNCH_CHECK();
    return ICPB200_OK;
}

int occ_ensure_dirty(OccGrid& g, cudaStream_t st) {
    const int n_tiles = ((g.nx + TS - 1) / TS) * ((g.ny + TS - 1) / TS);
    if (!g.dirty.p) {
        if (g.dirty.reserve((size_t)n_tiles)) return ICPB200_ERR_CUDA;
        ICPB_CUDA(cudaMemsetAsync(g.dirty.p, 0, g.dirty.cap, st));
        ICPB_CUDA(cudaStreamSynchronize(st));
    }
    return ICPB200_OK;
}

__global__ void __launch_bounds__(256) occ_view_all(const float* __restrict__ grid, float* __restrict__ out, size_t n, int view) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = occ_view(grid[i], view);
}

static int ensure_host_stage(OccGrid& g, size_t bytes) {
    if (bytes <= g.h_pack_cap) return ICPB200_OK;
    if (g.h_pack) { cudaFreeHost(g.h_pack); g.h_pack = nullptr; g.h_pack_cap = 0; }
    const size_t cap = bytes + bytes / 4 + 4096;
    ICPB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g.h_pack), cap, cudaHostAllocDefault));
    g.h_pack_cap = cap;
    return ICPB200_OK;
}

// dirty_only: write only the tiles touched since the last reset into `out` (the caller's array already holds the view of
// an untouched cell everywhere else); otherwise the whole map.  *tiles_out = tiles copied.
int occ_read_view(OccGrid& g, float* out, int view, bool dirty_only, int* tiles_out, cudaStream_t st) {
    const int tiles_x = (g.nx + TS - 1) / TS, tiles_y = (g.ny + TS - 1) / TS, n_tiles = tiles_x * tiles_y;
    const size_t n_cells = (size_t)g.nx * g.ny;
    if (tiles_out) *tiles_out = n_tiles;
    if (!dirty_only || g.all_dirty || !g.dirty.p) {
        if (dirty_only && !g.all_dirty && !g.dirty.p) { if (tiles_out) *tiles_out = 0; return ICPB200_OK; }   // never updated
        if (view == 0) {
            ICPB_CUDA(cudaMemcpyAsync(out, g.grid.p, sizeof(float) * n_cells, cudaMemcpyDeviceToHost, st));
        } else {
            if (g.pack.reserve(sizeof(float) * n_cells)) return ICPB200_ERR_CUDA;
            occ_view_all<<<2048, 256, 0, st>>>(g.grid.as<float>(), g.pack.as<float>(), n_cells, view);
            ICPB_LAUNCH_CHECK();
            ICPB_CUDA(cudaMemcpyAsync(out, g.pack.p, sizeof(float) * n_cells, cudaMemcpyDeviceToHost, st));
        }
        ICPB_CUDA(cudaStreamSynchronize(st));
        return ICPB200_OK;
    }
    {   // page-locked destination: the device writes the touched tiles in place
        cudaPointerAttributes attr;
        void* d_out = nullptr;
        if (cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
            cudaHostGetDevicePointer(&d_out, out, 0) == cudaSuccess && d_out) {
            if (g.small.reserve(256)) return ICPB200_ERR_CUDA;
            unsigned* d_count = g.small.as<unsigned>() + 48;
            ICPB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned), st));
            occ_write_tiles_mapped<<<(unsigned)n_tiles, 256, 0, st>>>(g.grid.as<float>(), g.nx, g.ny, tiles_x, g.dirty.as<unsigned char>(),
                                                                        static_cast<float*>(d_out), d_count, view);
            ICPB_LAUNCH_CHECK();
            unsigned n = 0;
            ICPB_CUDA(cudaMemcpyAsync(&n, d_count, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            ICPB_CUDA(cudaStreamSynchronize(st));
            if (tiles_out) *tiles_out = (int)n;
            return ICPB200_OK;
        }
        cudaGetLastError();                                      // a pageable destination is not an error: staging path below
    }
    int rc = ensure_host_stage(g, (size_t)n_tiles * (1 + sizeof(int)));
    if (rc) return rc;
    unsigned char* h_dirty = g.h_pack;
    ICPB_CUDA(cudaMemcpyAsync(h_dirty, g.dirty.p, (size_t)n_tiles, cudaMemcpyDeviceToHost, st));
    ICPB_CUDA(cudaStreamSynchronize(st));
    std::vector<int> ids;
    ids.reserve(1024);
    for (int t = 0; t < n_tiles; ++t) if (h_dirty[t] & 1) ids.push_back(t);
    const size_t n = ids.size();
    if (tiles_out) *tiles_out = (int)n;
    if (n == 0) return ICPB200_OK;
    const size_t b_pack = sizeof(float) * TCELLS * n;
    if (g.pack.reserve(b_pack) || g.pack_ids.reserve(sizeof(int) * (size_t)n_tiles)) return ICPB200_ERR_CUDA;
    if ((rc = ensure_host_stage(g, b_pack + sizeof(int) * n))) return rc;
    int* h_ids = reinterpret_cast<int*>(g.h_pack + b_pack);
    memcpy(h_ids, ids.data(), sizeof(int) * n);
    ICPB_CUDA(cudaMemcpyAsync(g.pack_ids.p, h_ids, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    occ_pack_tiles<<<(unsigned)n, 256, 0, st>>>(g.grid.as<float>(), g.nx, g.ny, tiles_x, g.pack_ids.as<int>(), g.pack.as<float>(), view);
    ICPB_LAUNCH_CHECK();
    ICPB_CUDA(cudaMemcpyAsync(g.h_pack, g.pack.p, b_pack, cudaMemcpyDeviceToHost, st));
    ICPB_CUDA(cudaStreamSynchronize(st));
    const float* src = reinterpret_cast<const float*>(g.h_pack);
    for (size_t k = 0; k < n; ++k) {
        const int t = ids[k], tx0 = (t % tiles_x) * TS, ty0 = (t / tiles_x) * TS;
        const int w = std::min(TS, g.nx - tx0), hgt = std::min(TS, g.ny - ty0);
        for (int r = 0; r < hgt; ++r)
            memcpy(out + (size_t)(ty0 + r) * g.nx + tx0, src + k * TCELLS + (size_t)r * TS, sizeof(float) * (size_t)w);
    }
    return ICPB200_OK;
}

int occ_fast_ctas(int sm_count) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, occ_fast_tiles, kTileNT, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    return sm_count * per_sm;
}

int occ_collect(OccGrid& g) {
    if (!g.stats_pending) return ICPB200_OK;
    g.stats_pending = false;
    ICPB_CUDA(cudaEventSynchronize(g.ev_stats));
    const unsigned long long* hs = reinterpret_cast<const unsigned long long*>(g.pending_host + 64);
    for (int k = 1; k < 4; ++k) g.stats[k] = (long long)hs[k];
    if (reinterpret_cast<const unsigned*>(g.pending_host)[3]) {
        set_error("grid_update: more than 4095 hits landed in one cell within one scan of the previous update (unsupported)");
        return ICPB200_ERR_LIMIT;
    }
    return ICPB200_OK;
}

int occ_update_fast(OccGrid& g, int n_scans, const double* d_origins, const double* d_hits,
                    const long long* d_hit_off, const long long* h_hit_off, long long total_hits, bool defer,
                    cudaStream_t st) {
    const long long n_rays = total_hits;
    g.stats[0] = n_rays; g.stats[1] = g.stats[2] = g.stats[3] = 0;
    if (n_rays <= 0) return ICPB200_OK;                                   // mapping.py:113-114
    if (!h_hit_off && n_scans > kOccMaxChunkScans) { set_error("grid_update: host offsets needed for more than one chunk"); return ICPB200_ERR_ARG; }
    const int n_tiles = ((g.nx + TS - 1) / TS) * ((g.ny + TS - 1) / TS);
    const size_t n_cells = (size_t)g.nx * g.ny;
    const long long max_chunk_rays = [&] {
        if (!h_hit_off) return n_rays;
        long long m = 0;
        for (int s0 = 0; s0 < n_scans; s0 += kOccMaxChunkScans)
            m = std::max(m, h_hit_off[std::min(n_scans, s0 + kOccMaxChunkScans)] - h_hit_off[s0]);
        return m;
    }();
    if (g.origin_cell.reserve(sizeof(int2) * (size_t)n_scans) || g.ray_cell.reserve(sizeof(int2) * (size_t)max_chunk_rays) ||
        g.ray_scan.reserve(sizeof(int) * (size_t)max_chunk_rays) || g.order.reserve(sizeof(int) * (size_t)n_tiles) ||
        g.small.reserve(256))
        return ICPB200_ERR_CUDA;
    if (!g.slotmap.p) {
        if (g.slotmap.reserve(sizeof(unsigned) * n_cells)) return ICPB200_ERR_CUDA;
        ICPB_CUDA(cudaMemsetAsync(g.slotmap.p, 0xff, sizeof(unsigned) * n_cells, st));
    }
    if (!g.ncount.p) {
        if (g.ncount.reserve(sizeof(unsigned) * n_cells)) return ICPB200_ERR_CUDA;
        ICPB_CUDA(cudaMemsetAsync(g.ncount.p, 0, sizeof(unsigned) * n_cells, st));
    }
    if (!g.dirty.p) {
        if (g.dirty.reserve((size_t)n_tiles)) return ICPB200_ERR_CUDA;
        ICPB_CUDA(cudaMemsetAsync(g.dirty.p, 0, (size_t)n_tiles, st));
    }
    if (defer && !g.pending_host) {
        ICPB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g.pending_host), 256, cudaHostAllocDefault));
        ICPB_CUDA(cudaEventCreateWithFlags(&g.ev_stats, cudaEventDisableTiming));
    }
    ICPB_CUDA(cudaMemsetAsync(g.small.p, 0, 256, st));
    occ_fast_origins<<<(n_scans + 255) / 256, 256, 0, st>>>(d_origins, n_scans, g.min_x, g.min_y, g.res, g.origin_cell.as<int2>(),
                                                            d_hit_off, total_hits,
                                                            h_hit_off ? nullptr : g.small.as<unsigned>() + kFlagBigScan);
    ICPB_LAUNCH_CHECK();
    std::vector<long long> fetched;                                        // host copy of the offsets, if it is needed after all
    long long tot[4] = {0, 0, 0, 0};
    bool all_fast = true;
    for (int s0 = 0; s0 < n_scans; s0 += kOccMaxChunkScans) {
        const int cs = std::min(kOccMaxChunkScans, n_scans - s0);
        long long rb = 0, re = n_rays;
        int big = -1;
        if (h_hit_off) {
            rb = h_hit_off[s0]; re = h_hit_off[s0 + cs];
            big = 0;
            bool huge = false;
            for (int s = s0; s < s0 + cs; ++s) {
                const long long len = h_hit_off[s + 1] - h_hit_off[s];
                if (len > 4095) big = 1;
                if (len >= (long long)kHitUnit) huge = true;      // a cell could collect 2^20 misses of one scan: beyond the event word
            }
            if (huge) {
                all_fast = false;
                g.all_dirty = true;
                std::vector<long long> off((size_t)cs + 1);
                for (int k = 0; k <= cs; ++k) off[k] = h_hit_off[s0 + k] - h_hit_off[s0];
                if (g.hit_off_shift.reserve(sizeof(long long) * off.size())) return ICPB200_ERR_CUDA;
                ICPB_CUDA(cudaMemcpyAsync(g.hit_off_shift.p, off.data(), sizeof(long long) * off.size(), cudaMemcpyHostToDevice, st));
                ICPB_CUDA(cudaStreamSynchronize(st));
                long long keep[4] = {g.stats[0], g.stats[1], g.stats[2], g.stats[3]};
                int rc2 = occ_update_ordered(g, cs, d_origins + 2 * (size_t)s0, d_hits + 2 * (size_t)h_hit_off[s0],
                                             g.hit_off_shift.as<long long>(), off.data(), st);
                if (rc2) return rc2;
                for (int k = 1; k < 4; ++k) tot[k] += g.stats[k];
                for (int k = 0; k < 4; ++k) g.stats[k] = keep[k];
                continue;
            }
        }
        bool bad_offsets = false;
        int rc = fast_chunk(g, n_scans, s0, cs, d_origins, d_hits, d_hit_off, rb, re, big, &bad_offsets, st);
        if (bad_offsets) {
            set_error("grid_update: hit_off[0] must be 0, hit_off must not decrease and hit_off[n_scans] must equal total_hits");
            return ICPB200_ERR_ARG;
        }
        if (rc == 1) {
            all_fast = false;
            g.all_dirty = true;
            if (!h_hit_off) {                                              // single chunk, offsets checked on the device
                fetched.resize((size_t)n_scans + 1);
                ICPB_CUDA(cudaMemcpyAsync(fetched.data(), d_hit_off, sizeof(long long) * fetched.size(), cudaMemcpyDeviceToHost, st));
                ICPB_CUDA(cudaStreamSynchronize(st));
                h_hit_off = fetched.data();
            }
            long long keep[4] = {g.stats[0], g.stats[1], g.stats[2], g.stats[3]};
            std::vector<long long> off((size_t)cs + 1);
            for (int k = 0; k <= cs; ++k) off[k] = h_hit_off[s0 + k] - h_hit_off[s0];
            // the ordered path wants offsets that start at 0: shift the device copies as well
            if (g.hit_off_shift.reserve(sizeof(long long) * off.size())) return ICPB200_ERR_CUDA;
            ICPB_CUDA(cudaMemcpyAsync(g.hit_off_shift.p, off.data(), sizeof(long long) * off.size(), cudaMemcpyHostToDevice, st));
            ICPB_CUDA(cudaStreamSynchronize(st));
            rc = occ_update_ordered(g, cs, d_origins + 2 * (size_t)s0, d_hits + 2 * (size_t)h_hit_off[s0],
                                    g.hit_off_shift.as<long long>(), off.data(), st);
            if (rc) return rc;
            for (int k = 1; k < 4; ++k) tot[k] += g.stats[k];
            for (int k = 0; k < 4; ++k) g.stats[k] = keep[k];
            continue;
        }
        if (rc) return rc;
        if (defer && all_fast && n_scans <= kOccMaxChunkScans) {
            // the only chunk: statistics and the overflow flag are collected later (occ_collect)
            ICPB_CUDA(cudaMemcpyAsync(g.pending_host, g.small.p, 256, cudaMemcpyDeviceToHost, st));
            ICPB_CUDA(cudaEventRecord(g.ev_stats, st));
            g.stats_pending = true;
            g.seen_nonempty_scan = true;
            return ICPB200_OK;
        }
        unsigned char host_small[256];
        ICPB_CUDA(cudaMemcpyAsync(host_small, g.small.p, 256, cudaMemcpyDeviceToHost, st));
        ICPB_CUDA(cudaStreamSynchronize(st));
        const unsigned long long* hs = reinterpret_cast<const unsigned long long*>(host_small + 64);
        for (int k = 1; k < 4; ++k) tot[k] += (long long)hs[k];
        ICPB_CUDA(cudaMemsetAsync(g.small.as<unsigned char>() + 64, 0, 64, st));
        if (reinterpret_cast<const unsigned*>(host_small)[3]) {
            set_error("grid_update: more than 4095 hits landed in one cell within one scan (unsupported)");
            return ICPB200_ERR_LIMIT;
        }
    }
    g.stats[0] = n_rays;
    for (int k = 1; k < 4; ++k) g.stats[k] = tot[k];
    g.seen_nonempty_scan = true;
    return ICPB200_OK;
}

}  // namespace icpb
