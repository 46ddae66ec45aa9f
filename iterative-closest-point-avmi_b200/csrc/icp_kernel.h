// Host <-> kernel argument blocks and launchers for the ICP kernels.
//
// Three kernels (SURVEY.md section 2 kernel inventory K1/K2/K3):
//   K1 voxel_clouds_kernel   one CTA per referenced cloud: voxel-grid mean
//   K2 normals_kernel        one CTA per cloud used as a point-to-line target
//   K3 icp_pairs_kernel      persistent, one CTA per (source, target) pair:
//                            the whole iteration loop on-chip
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace icpb {

// One set of clouds on the device plus its preprocessed (downsampled) form.
// Cloud c owns raw rows off[c] .. off[c+1]) of `raw`; its downsampled rows live
// at the same offset in `ds` (count ds_n[c] <= raw count), its normals at the
// same offset in `nrm` (2 doubles per row), its raw bounding box in box[c*6..].
struct CloudSet {
    const double* raw;
    const long long* off;
    int n_clouds;
    double* ds;
    int* ds_n;              // -1: voxel index range does not fit (cloud unusable)
    double* box;            // lo[3], hi[3] per cloud
    double* nrm;            // may be nullptr when no cloud of the set is a p2l target
    unsigned char* used;    // per cloud: 1 if referenced by any pair (nullptr: all are)
    unsigned char* is_tgt;  // per cloud: 1 if it is the target of a point-to-line pair
};

// Hash grid over one (downsampled) target cloud in global memory (grid mode).
// Bucket b holds items[(b ? start[b-1] : 0) .. start[b]); cell[e] is the grid cell
// of items[e] (several cells may share a bucket).
struct BigGrid {
    const int* start;
    const int* items;
    const int2* cell;
    double h, lox, loy;
    int nx, ny;
    unsigned mask;             // buckets - 1 (power of two)
};

__host__ __device__ inline unsigned big_cell_hash(int cx, int cy) {
    return ((unsigned)cx * 73856093u) ^ ((unsigned)cy * 19349663u);
}

struct IcpArgs {
    int n_pairs;
    CloudSet s, t;             // may describe the same set
    const int* src_idx;        // per pair, or nullptr (pair p -> cloud p)
    const int* tgt_idx;
    const double* R_init;      // n_pairs * dim * dim or nullptr
    const double* t_init;      // n_pairs * dim or nullptr
    double err_thr;
    int max_iter;
    double voxel;
    int brute_slab;            // brute mode: sweep the voxel-ordered target outwards from the query (nn_slab)
    int tma_normals;           // 512-thread variant: stage the target's normals in shared memory with a TMA bulk copy
    int method;
    int normal_k;
    double max_corr;           // < 0: no gate
    double* R_out;
    double* t_out;
    double* err_out;
    double* prev_out;          // may be nullptr
    int* iters_out;
    int* status_out;
    int cap_s, cap_t;          // multiples of 32, >= largest raw cloud of each set
    int sort_pad;              // power of two >= largest raw cloud, >= 256
    unsigned int* queue;       // zeroed before launch
    // a launch may cover only the pairs order[pair_first .. pair_first + n_pairs) (order == nullptr: the identity):
    // the host-buffer entry point registers the pairs of an upload chunk while the next chunk is still on its way
    int pair_first;
    const int* pair_order;
    unsigned long long* stats; // [0] fp32 sweep pair evaluations executed, [1] points re-decided by the
                               // full fp64 scan, [2] iterations, [3] source points swept, [4] source points
                               // whose correspondence was carried over by the movement bound; may be nullptr
    const BigGrid* grids;      // grid mode: one per target cloud (device array), else nullptr
    // two-phase schedule: the bulk launch hands pairs that need more than phase_cap iterations to a
    // second launch (resume = 1) that runs one CTA per SM, so the long tail does not share its SM
    int phase_cap;             // <= 0: single launch
    int resume;
    unsigned int* cont_count;  // [0] pairs handed over, [1..3] per cost class
    int* cont_list;            // [slot] -> pair
    int* cont_bucket;          // [class][cont_cap] -> slot: the second launch takes the expensive classes first
    int cont_cap;              // pairs of the whole call (row length of cont_bucket)
    int class_lo, class_hi;    // resume launches: the cost classes [class_lo, class_hi) this launch takes
    int coop_ctas;             // resume launches: CTAs of the cluster-variant launch running beside this one (0: none)
    int coop_factor;           // ... which takes the work when at most this many pairs per cluster CTA were handed over
    double* cont_cur;          // [slot][dim][cap_s]
    int* cont_match;           // [slot][cap_s]
    float* cont_d2lb;
    float* cont_moved;
    double* cont_scalar;       // [slot][16]: r_tot(9) t_tot(3) prev err iters
    unsigned long long* pair_prof; // optional [pair][8]: SM cycles, points swept, fp64 rescans, iterations, cycles of classify / NN / rest (launches add up)
    int* trace_match;          // optional: correspondences of the first trace_iters iterations (pair 0)
    int trace_iters;
    int trace_stride;
};

// K8: rotation-search scoring.  Problems are independent (source, target, angle list, shift).
struct RotArgs {
    const double* src; const long long* src_off;       // (sum n_s, 2), n_problems + 1
    const double* tgt; const long long* tgt_off;
    const double* angles; const long long* ang_off;    // radians
    const double* shift;                               // 2 per problem: added after the rotation
    double* scores;                                    // one per angle
    double* nn_dist; int* nn_idx;                      // optional, per source point (single-angle queries)
    int cap_t;                                         // multiple of 32 >= the largest target (slice)
    int angles_per_cta;                                // set by the launcher
    // targets beyond shared memory: slices of slice_len points, one launch per slice (slice_len 0: whole targets)
    int slice_len, slice, n_slices;
    double* part_d2; int* part_j;                      // running minimum per (angle, source point), rows of part_stride
    long long part_stride;
};
size_t rot_smem_bytes(int cap_t);
int launch_rot_scores(const RotArgs& a, int n_problems, int max_angles, int sm_count, cudaStream_t stream);

size_t icp_pair_smem_bytes(int dim, int cap_s, int cap_t, int nt = 256);      // nt: threads per CTA of the variant (256 bulk, 512 roomy)
size_t icp_voxel_smem_bytes(int sort_pad);
size_t icp_normals_smem_bytes(int cap_t);
int icp_max_ctas_per_sm(int dim, bool grid, size_t smem);

// used[c] = 1 for every referenced cloud, is_tgt[c] = 1 for p2l targets (device-side, from the idx arrays)
int launch_mark_used(const IcpArgs& a, bool p2l, cudaStream_t stream);
int launch_voxel_clouds(const CloudSet& cs, int dim, double voxel, int sort_pad, cudaStream_t stream, int first = 0, int count = -1);
int launch_normals(const CloudSet& cs, int cap_t, int normal_k, double voxel, cudaStream_t stream, int first = 0, int count = -1);
int launch_icp_pairs(const IcpArgs& a, int dim, bool grid, int n_ctas, size_t smem_min, cudaStream_t stream);
// the most expensive class of handed-over pairs on clusters of 4 CTAs (2-D brute mode only); false: not available
bool icp_cluster_variant(int dim, bool grid);
int launch_icp_pairs_cluster(const IcpArgs& a, int max_ctas, size_t smem_min, cudaStream_t stream, int* n_ctas_out);

// big-cloud kernels (icp_big.cu)
int launch_big_voxel(const CloudSet& cs, int dim, double voxel, unsigned long long* key_buf, unsigned* idx_buf,
                     long long total_points, cudaStream_t stream);
int launch_big_grid(const CloudSet& cs, double cell_size, const long long* d_grid_off, const int* d_buckets,
                    int* start_buf, int* items_buf, int2* cell_buf, BigGrid* d_grids, cudaStream_t stream);
int launch_big_normals(const CloudSet& cs, const BigGrid* d_grids, int normal_k, long long max_points,
                       cudaStream_t stream);

}  // namespace icpb
