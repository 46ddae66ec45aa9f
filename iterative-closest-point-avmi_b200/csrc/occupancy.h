// Host-side state of one occupancy grid handle and the device entry point.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace icpb {

constexpr int kOccTile = 32;                       // cells per tile edge (power of two)
constexpr int kOccOwnTile = 64;                    // multi-GPU ownership granularity (cells); see occ_owner
constexpr int kOccMaxChunkScans = 2048;            // scans replayed per tile pass
constexpr long long kOccMaxMatrix = 64LL << 20;    // (tile, scan) counters per pass
constexpr size_t kOccOrdBudget = 1ull << 30;       // order-free path: (hit cell, scan) counters per chunk (4 GiB)

// Spatial sharding: the grid is cut into bands of kOccOwnTile rows (one row of 64 x 64-cell tiles) that are dealt
// round-robin: rank r owns the bands b with b % world == r.  A map's activity is spatially concentrated (on C4 one of
// eight contiguous strips carried 44 % of the traversed cells); neighbouring bands see nearly the same load, so dealing
// them out balances the ranks, and a ray is still clipped to a band in closed form: a rank walks only the parts of every
// ray that cross its own bands.  Both device paths use this one rule, so a cell never changes owner between calls.
__host__ __device__ inline int occ_mod_world(int b, int world) {       // b >= 0; the usual rank counts are powers of two
    return (world & (world - 1)) == 0 ? (b & (world - 1)) : b % world;
}
__host__ __device__ inline int occ_owner(int x, int y, int nx, int ny, int world) {
    (void)x; (void)nx; (void)ny;
    return occ_mod_world(y / kOccOwnTile, world);
}

struct OccGrid {
    int nx = 0, ny = 0, tiles_x = 0, tiles_y = 0;
    double min_x = 0, min_y = 0, res = 1;
    double l_hit = 0, l_miss = 0, lo_min = 0, lo_max = 0;
    int rank = 0, world = 1;
    int apply_ctas = 0;
    // A clamp interval that excludes 0 makes the reference's whole-grid clip
    // (mapping.py:141) move every untouched cell on the first non-empty scan.
    bool zero_outside_clamp = false;
    bool seen_nonempty_scan = false;
    bool virgin_finalised = false;
    long long stats[4] = {0, 0, 0, 0};
    DevBuf grid;                                   // ny * nx float32, row major
    DevBuf origins, hits, hit_off;                 // staging for the host-buffer entry point
    DevBuf local_pts, poses;                       // icpb200_grid_rebuild: scans in their local frames + 3x3 poses
    DevBuf in_pack;                                // small updates: origins | hit_off | hits in one block ...
    unsigned char* h_in_pack = nullptr;            // ... uploaded from one page-locked block with one copy
    size_t h_in_cap = 0;
    DevBuf origin_cell, ray_cell, ray_scan;
    DevBuf counts, offsets, sums, runs, order, small, tile_prof;
    // order-free path (occupancy_fast.cu)
    DevBuf slotmap, slot_cell, ord, tile_count, hit_off_shift, items, multi, ncount, ev, ev_count, class_off, tile_flag;
    // read-out: tiles (64 x 64 cells) touched since the last reset; the ordered path does not track them (all_dirty)
    DevBuf dirty, pack, pack_ids;
    bool all_dirty = false;
    // multi-GPU: the peers' grids and touched-tile flags, mapped through CUDA IPC (icpb200_grid_ipc_attach)
    void* peer_grid[16] = {};
    void* peer_dirty[16] = {};
    bool peers_attached = false;
    unsigned char* h_pack = nullptr;               // page-locked staging of the packed tiles
    size_t h_pack_cap = 0;
    cudaStream_t aux_stream = nullptr;             // the hit cells' replay runs here, under the remaining tiles
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_hit = nullptr;
    cudaEvent_t ev_done = nullptr;                 // end of the last stream-ordered update (icpb200_grid_update_dev)
    // Deferred read-back of the device-resident entry point (occ_update_fast with defer = true): the statistics and
    // the hit-overflow flag of the last update land in `pending_host` (page-locked) behind `ev_stats`; occ_collect()
    // folds them into `stats` and reports the flag.  Every entry point that touches the grid collects first.
    unsigned char* pending_host = nullptr;
    cudaEvent_t ev_stats = nullptr;
    bool stats_pending = false;
    int fast_ctas = 0;
    bool use_fast = true;                          // ICPB200_OCC_PATH=ordered forces the ordered tile replay
    bool profile_tiles = false;                    // icpb200_grid_tile_profile() requested per-tile timings
    void release_all();
};

// d_* are device pointers; h_hit_off is the same offsets array on the host
// (n_scans + 1 entries, h_hit_off[0] == 0).  Stream-ordered except for one
// small readback per scan chunk.
int occ_update_device(OccGrid& g, int n_scans, const double* d_origins, const double* d_hits,
                      const long long* d_hit_off, const long long* h_hit_off, cudaStream_t st);
int occ_update_ordered(OccGrid& g, int n_scans, const double* d_origins, const double* d_hits,
                       const long long* d_hit_off, const long long* h_hit_off, cudaStream_t st);
// h_hit_off may be null when n_scans <= kOccMaxChunkScans: the offsets are then validated on the device and the host
// never waits for them.  defer = true leaves the final statistics / overflow read-back to occ_collect().
int occ_update_fast(OccGrid& g, int n_scans, const double* d_origins, const double* d_hits,
                    const long long* d_hit_off, const long long* h_hit_off, long long total_hits, bool defer,
                    cudaStream_t st);
// Waits for a deferred read-back (if any); returns ICPB200_ERR_LIMIT when the update it belongs to overflowed.
int occ_collect(OccGrid& g);
// Read-out as log-odds (view 0), probability (1) or display value (2), mapping.py:150-160; dirty_only copies just the tiles
// touched since the last reset (occupancy_fast.cu).
int occ_read_view(OccGrid& g, float* out, int view, bool dirty_only, int* tiles_out, cudaStream_t st);
// multi-GPU exchange of touched tiles by peer stores (occupancy_fast.cu)
int occ_push_to_peers(OccGrid& g, cudaStream_t st);
int occ_ensure_dirty(OccGrid& g, cudaStream_t st);
// slam.py:46-50 for every scan of a history: world = local @ R.T + t, origin = t (device pointers)
int occ_transform_history(int n_scans, long long n_points, const double* d_poses, const double* d_local,
                          const long long* d_off, double* d_world, double* d_origins, cudaStream_t st);
int occ_apply_ctas(int sm_count);
int occ_fast_ctas(int sm_count);

}  // namespace icpb
