/*
 * libicp_b200.so -- C ABI of the B200-native ICP registration + occupancy
 * raycast hot path (sm_100a).
 *
 * The reference (DUBSON0/iterative-closest-point-avmi) is pure Python and has
 * no FFI of its own; its boundary for this path is the Python call surface
 *     utilities/icp.py:132-134      ICP(source, target, error_threshold, ...)
 *     utilities/icp.py:117          voxel_downsample(points, voxel_size)
 *     utilities/mapping.py:28-37    OccupancyGrid2D(...)
 *     utilities/mapping.py:103      OccupancyGrid2D.update_scan(origin, hits)
 *     utilities/mapping.py:143      OccupancyGrid2D.reset()
 *     utilities/mapping.py:47       OccupancyGrid2D.log_odds  (ny, nx) float32
 * Each entry point below names the reference lines it replaces.  A ctypes
 * binding (the one the drop-in `utilities` shim uses) is shown in
 * INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; every host buffer is owned by the caller and is not
 *     retained after the call returns; all device memory, streams and handles
 *     are owned by the library;
 *   - calls are blocking: results are valid on return (the *_dev variants are
 *     stream-ordered instead and return after enqueueing);
 *   - return value 0 = success, negative = failure (see ICPB200_ERR_*), text in
 *     icpb200_last_error().  Numerical outcomes the reference treats as normal
 *     (singular normal equations -> identity step; too few inliers -> early
 *     exit) are NOT errors: they are reported per pair in status_out;
 *   - there is no CPU fallback: without a usable CUDA device every compute
 *     call fails with ICPB200_ERR_CUDA;
 *   - one process drives one GPU (torch.distributed / torchrun style);
 *     icpb200_init(device) selects it.  Calls are serialised by an internal
 *     mutex.
 */
#ifndef ICP_B200_H
#define ICP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICPB200_OK            0
#define ICPB200_ERR_CUDA     -1   /* CUDA runtime / driver failure, or no device */
#define ICPB200_ERR_ARG      -2   /* invalid argument */
#define ICPB200_ERR_LIMIT    -3   /* input exceeds a documented limit of this build */

/* per-pair status codes (status_out) */
#define ICPB200_CONVERGED     0   /* |prev_error - error| < error_threshold   (icp.py:216-219) */
#define ICPB200_MAX_ITER      1   /* loop exhausted                           (icp.py:222-223) */
#define ICPB200_FEW_INLIERS   2   /* inliers < max(3, N/10): loop left early  (icp.py:186-187) */
#define ICPB200_BAD_VOXELS    3   /* pair not processed: voxel index range beyond 62 bits, or (device-resident
                                     entry point) a cloud index outside the set */

/* method (icp.py:133, 162) */
#define ICPB200_SINGULAR      4   /* pose graph only: the normal matrix is not positive definite (pose_graph.py:115-119) */

#define ICPB200_POINT_TO_POINT 0
#define ICPB200_POINT_TO_LINE  1  /* 2-D only; silently point-to-point for dim 3, as the reference */

/* nn_mode */
#define ICPB200_NN_AUTO  0        /* brute force up to ICPB200_BRUTE_MAX_POINTS raw points, else grid */
#define ICPB200_NN_BRUTE 1        /* shared-memory tiled brute force (fp32 sweep + fp64 decision) */
#define ICPB200_NN_GRID  2        /* uniform-grid exact nearest neighbour (fp64) */

#define ICPB200_BRUTE_MAX_POINTS 4096

/* ---- library / device ------------------------------------------------- */

/* Bind this process to CUDA device `device` (-1: keep the current device) and
 * create the library's stream and workspaces.  Idempotent for the same
 * device. */
int icpb200_init(int device);
void icpb200_shutdown(void);
const char *icpb200_last_error(void);
/* Number of CUDA kernels this library has launched since load (bench.py's
 * gpu_launches). */
int64_t icpb200_launch_count(void);
/* Compile-time facts for tests: returns the sm architecture the kernels were
 * built for (100 for sm_100a). */
int icpb200_built_arch(void);

/* ---- ICP registration  (replaces utilities/icp.py:132-223) ------------- */

/*
 * Register n_pairs independent (source, target) pairs.  Pair p's source is
 * src[src_off[p] .. src_off[p+1]) (rows of `dim` float64, C order), likewise
 * the target.  R_init / t_init: n_pairs*dim*dim / n_pairs*dim values, or NULL
 * for the identity start (the reference uses the initial guess only when
 * BOTH are given, icp.py:153).  max_corr_dist < 0 means None (icp.py:169).
 * Outputs: R_out n_pairs*dim*dim (row major), t_out n_pairs*dim, err_out,
 * prev_err_out (error of the iteration before the last one, +inf if none; may
 * be NULL -- it lets a caller print the reference's `delta`, icp.py:216-218),
 * iters_out (completed solve steps), status_out (ICPB200_* above); the
 * forward transform is p' = R p + t as in the reference.
 * A single ICP() call is n_pairs = 1.
 */
int icpb200_icp_batch(int n_pairs, int dim,
                      const double *src, const int64_t *src_off,
                      const double *tgt, const int64_t *tgt_off,
                      const double *R_init, const double *t_init,
                      double error_threshold, int max_iterations, double voxel_size,
                      int method, int normal_k, double max_corr_dist, int nn_mode,
                      double *R_out, double *t_out, double *err_out, double *prev_err_out,
                      int32_t *iters_out, int32_t *status_out);

/*
 * Same registration loop over pairs drawn from one set of clouds (scan
 * history): cloud c is pts[cloud_off[c] .. cloud_off[c+1]); pair p registers
 * cloud src_idx[p] onto cloud tgt_idx[p].  This is the batch seam of the
 * reference's loop-closure candidate loop (slam.py:575-579) and of offline
 * scan-pair sweeps; the clouds cross PCIe once.
 */
int icpb200_icp_pairs(int n_clouds, int dim,
                      const double *pts, const int64_t *cloud_off,
                      int n_pairs, const int32_t *src_idx, const int32_t *tgt_idx,
                      const double *R_init, const double *t_init,
                      double error_threshold, int max_iterations, double voxel_size,
                      int method, int normal_k, double max_corr_dist, int nn_mode,
                      double *R_out, double *t_out, double *err_out, double *prev_err_out,
                      int32_t *iters_out, int32_t *status_out);

/*
 * Device-resident variant of icpb200_icp_pairs: every pointer is a device
 * pointer, work is enqueued on `stream` (a cudaStream_t; NULL = the library
 * stream) and the call returns without synchronising.  max_cloud_points is an
 * upper bound on the raw size of any cloud (sizes workspaces).
 */
int icpb200_icp_pairs_dev(int n_clouds, int dim,
                          const double *d_pts, const int64_t *d_cloud_off,
                          int64_t max_cloud_points,
                          int n_pairs, const int32_t *d_src_idx, const int32_t *d_tgt_idx,
                          const double *d_R_init, const double *d_t_init,
                          double error_threshold, int max_iterations, double voxel_size,
                          int method, int normal_k, double max_corr_dist, int nn_mode,
                          double *d_R_out, double *d_t_out, double *d_err_out, double *d_prev_err_out,
                          int32_t *d_iters_out, int32_t *d_status_out, void *stream);

/*
 * One registration with its intermediate state exposed, for parity tests:
 * the voxel-downsampled clouds (icp.py:150-151), the target normals
 * (icp.py:165-167; 2-D point-to-line only) and the correspondence indices of
 * the first trace_iters iterations (icp.py:179).  Buffers may be NULL.
 * src_ds / tgt_ds need n_src*dim / n_tgt*dim doubles, normals n_tgt*2,
 * matches trace_iters*n_src int32 (row i = iteration i, -1 where unused).
 */
int icpb200_icp_trace(int dim, const double *src, int64_t n_src,
                      const double *tgt, int64_t n_tgt,
                      const double *R_init, const double *t_init,
                      double error_threshold, int max_iterations, double voxel_size,
                      int method, int normal_k, double max_corr_dist, int nn_mode,
                      double *R_out, double *t_out, double *err_out, double *prev_err_out,
                      int32_t *iters_out, int32_t *status_out,
                      double *src_ds, int64_t *n_src_ds,
                      double *tgt_ds, int64_t *n_tgt_ds,
                      double *normals, int32_t *matches, int trace_iters);

/* Work counters of the most recent registration call (synchronises with it):
 * stats[0] fp32 sweep pair evaluations executed, stats[1] source points
 * re-decided by the full fp64 scan, stats[2] iterations over all pairs,
 * stats[3] source points swept, stats[4] source points whose correspondence
 * was carried over by the exact movement bound; stats[5..7] device time in ns
 * of the voxel (K1), normals (K2) and pair (K3) kernels (CUDA events).
 * bench.py uses them for the FP32-FMA roofline. */
int icpb200_icp_last_stats(int64_t *stats8);

/* Profiling aid: SM cycles spent by thread 0 of the pair kernel in the five
 * phases of an iteration (classify, nearest neighbour, accumulate + reduce,
 * solve, apply + error), summed over iterations >= 8 of all pairs of the last
 * call; out8[5] is the number of such iterations. */
int icpb200_icp_phase_profile(int64_t *out8);

/* More counters of the most recent registration call: out8[0] iterations
 * decided against the far-field front set (a diverged source), out8[1..3] grid
 * mode: nearest-neighbour queries, target points evaluated, grid cells
 * visited (bench.py: bytes per query of the hash-grid roofline); two-phase
 * batches: out8[4..6] pairs handed over to the second launch by cost class
 * (256+ points to decide per iteration, 96+, fewer), out8[7] how many times a
 * CTA of the cluster variant joined a cluster mate's pair as a helper. */
int icpb200_icp_extra_stats(int64_t *out8);

/* Profiling aid: the first call switches per-pair counters on; after the next
 * registration call, a call with a buffer of 8*cap_pairs int64 receives, per
 * pair of that call, {SM cycles spent on the pair (both launches), source
 * points swept, points re-decided by the fp64 fallback, iterations, cycles of
 * the classify phase, of the nearest-neighbour phase, of the rest (iterations
 * >= 8 only), 0} and returns the number of pairs written. */
int icpb200_icp_pair_profile(int64_t *out, int64_t cap_pairs);

/* Voxel-grid mean downsample (replaces utilities/icp.py:117-129).  `out`
 * needs n*dim doubles; *n_out receives the number of occupied voxels; rows
 * are in lexicographic voxel-index order like np.unique(axis=0). */
int icpb200_voxel_downsample(const double *pts, int64_t n, int dim, double voxel_size,
                             double *out, int64_t *n_out);

/* ---- occupancy grid  (replaces utilities/mapping.py:28-145) ------------ */

/* mapping.py:28-52.  The caller computes nx, ny, l_hit, l_miss exactly as the
 * reference does (ceil of extent/resolution; log(p/(1-p)) in float64). */
void *icpb200_grid_create(int nx, int ny, double min_x, double min_y, double resolution,
                          double l_hit, double l_miss, double lo_min, double lo_max);
void icpb200_grid_destroy(void *grid);

/* Multi-GPU spatial sharding: the grid is cut into bands of 64 rows (one row
 * of 64 x 64-cell tiles) dealt round-robin; this process updates only the
 * bands b with b % world == rank and never touches any other row, so a gather
 * of every rank's bands reassembles the map (icp_b200.dist.grid_gather_device).
 * Call before the first update of the grid. */
int icpb200_grid_set_shard(void *grid, int rank, int world);

/* mapping.py:103-141 for n_scans scans applied in array order (n_scans = 1 is
 * update_scan; n_scans > 1 is the _rebuild_map replay, slam.py:271-277).
 * origins: n_scans*2; hits: rows of 2 float64; scan s owns rows
 * hit_off[s] .. hit_off[s+1]. */
int icpb200_grid_update(void *grid, int n_scans, const double *origins,
                        const double *hits, const int64_t *hit_off);
/* Device-resident inputs, stream-ordered on `stream` (NULL = library stream).
 * total_hits = hit_off[n_scans] must be supplied by the caller.  For up to
 * 2048 scans per call the offsets are checked on the device and the call
 * returns with the update enqueued (the host waits once, for the binning
 * totals, while the fill pass runs): the grid is valid in stream order, the
 * statistics and a hit-field overflow (more than 4095 endpoints of one scan in
 * one cell) are collected -- and the overflow reported -- by the next call
 * that names this grid (update, read, reset, last_stats), each of which waits
 * for the update to finish first.  Readers of icpb200_grid_device_ptr() must
 * order themselves after `stream`. */
int icpb200_grid_update_dev(void *grid, int n_scans, const double *d_origins,
                            const double *d_hits, const int64_t *d_hit_off,
                            int64_t total_hits, void *stream);
/* slam.py:271-277 `_rebuild_map` with slam.py:46-50 `transform_points_2d` fused in:
 * clear the grid, then replay n_scans scans given in their LOCAL frames with their
 * current 3x3 homogeneous poses (row major, n_scans*9): world = local @ R.T + t
 * evaluated on the device as fma(p1, r_1, p0*r_0) + t -- the roundings of numpy's
 * matmul followed by the broadcast add -- origin = the pose's translation.
 * local_pts: rows of 2 float64; scan s owns rows off[s] .. off[s+1]. */
int icpb200_grid_rebuild(void *grid, int n_scans, const double *poses,
                         const double *local_pts, const int64_t *off);
/* Copy the (ny, nx) float32 log-odds array, row major [iy][ix], to `out`. */
int icpb200_grid_read(void *grid, float *out);
/* Read-out with the display transforms of mapping.py:150-160 evaluated on the
 * device, and -- with dirty_only != 0 -- only of the 64 x 64-cell tiles some
 * update has touched since the last reset (a map is mostly unexplored: the
 * copy shrinks from ny*nx*4 bytes to the explored part).
 *   view 0  log-odds                       (mapping.py:47)
 *   view 1  1 / (1 + exp(-log_odds))       (to_probability, mapping.py:150-153)
 *   view 2  1 - p, unexplored 1, free 0.85 (to_display, mapping.py:155-160)
 * float32 arithmetic throughout, as numpy evaluates it; the device exp differs
 * from numpy's float32 exp by at most 2 ulp (views 1 and 2 agree to 1e-6).
 * With dirty_only the cells of untouched tiles are NOT written: `out` must
 * already hold the view of an untouched cell there (0, 0.5 or 1 -- a mirror the
 * caller keeps between calls; in a grid shared out with
 * icpb200_grid_set_shard only this rank's tiles count as touched).
 * tiles_copied (may be NULL) receives the number of tiles that crossed PCIe. */
int icpb200_grid_read_view(void *grid, int view, int dirty_only, float *out, int32_t *tiles_copied);
/* mapping.py:143-145 */
int icpb200_grid_reset(void *grid);
/* Multi-GPU exchange of touched tiles by peer stores over NVLink (one process per
 * GPU on one node).  Every rank exports CUDA IPC handles of its grid and of its
 * touched-tile flags (2 x 64 bytes), the host framework gathers the handles of
 * all ranks (icp_b200.dist.grid_share_setup does it with one all_gather) and
 * every rank attaches them (rank-major, world x 128 bytes; call after
 * icpb200_grid_set_shard).  icpb200_grid_push_tiles then stores every tile this
 * rank owns and has touched since its last push into the same place of every
 * peer's grid and marks it touched there -- one kernel, stream-ordered, no
 * staging, only the explored part of the map moves.  A cross-rank barrier after
 * the push (any stream-ordered collective) makes the tiles visible to the
 * peers' readers; icpb200_grid_reset on all ranks must not overlap a peer's push. */
int icpb200_grid_ipc_export(void *grid, unsigned char *handles128);
int icpb200_grid_ipc_attach(void *grid, int world, int rank, const unsigned char *handles);
int icpb200_grid_push_tiles(void *grid, void *stream);
/* Raw device pointer of the ny*nx float32 grid (for NCCL collectives issued by
 * the host framework). */
void *icpb200_grid_device_ptr(void *grid);
/* Counters of the most recent update call: stats[0] = rays, stats[1] =
 * in-bounds free-cell updates (traversed cells), stats[2] = in-bounds hit
 * updates, stats[3] = tile runs.  Used by bench.py for the roofline bytes. */
int icpb200_grid_last_stats(void *grid, int64_t *stats4);

/* Profiling aid: the first call switches per-tile timing on; after the next
 * update, a call with a buffer of 4*cap_tiles int64 receives, per tile index,
 * {tile, scans replayed, runs, SM cycles spent on the tile} of the last scan
 * chunk (zeros for idle tiles) and returns the number of tiles written. */
int icpb200_grid_tile_profile(void *grid, int64_t *out, int64_t cap_tiles);

/* ---- device-resident submap -------------------------------------------------
 * The rolling window of global-frame scans the reference keeps as a Python list
 * (slam.py:559-562 `submap_buffer.append / pop(0)`, rebuilt after a loop closure,
 * slam.py:611-615) and everything it recomputes from the whole window in every
 * scan: np.vstack + voxel_downsample (slam.py:103-108 `_build_submap`) and, inside
 * ICP(scan, submap, ...) (slam.py:217-225), the second voxel_downsample of the
 * target plus its KD-tree (icp.py:151, 173).  Here the window lives on the device:
 * a push uploads ONE scan, the two downsamples and the hash grid are computed on
 * the device from the resident window and cached until the window changes, so
 * registrations against an unchanged window run the pair kernel only.
 * Results are those of ICP(source, voxel_downsample(vstack(window), submap_voxel), ...). */
void *icpb200_submap_create(int dim, int capacity_scans);
void icpb200_submap_destroy(void *submap);
/* append a scan (rows of dim float64, global frame); the oldest scan leaves once
 * capacity_scans are held (slam.py:561-562) */
int icpb200_submap_push(void *submap, const double *pts, int64_t n);
int icpb200_submap_clear(void *submap);
int icpb200_submap_size(void *submap, int64_t *n_scans, int64_t *n_points);
/* slam.py:103-108: *n_out = rows of voxel_downsample(vstack(window), submap_voxel);
 * `out` (may be NULL) receives them, bit-identical to the reference's array. */
int icpb200_submap_build(void *submap, double submap_voxel, double *out, int64_t out_capacity_rows, int64_t *n_out);
/* n_sources registrations against the window's submap (outputs as
 * icpb200_icp_batch; every source is registered onto the same target). */
int icpb200_submap_icp(void *submap, double submap_voxel, int n_sources, const double *src, const int64_t *src_off,
                       const double *R_init, const double *t_init, double error_threshold, int max_iterations,
                       double voxel_size, int method, int normal_k, double max_corr_dist, int nn_mode,
                       double *R_out, double *t_out, double *err_out, double *prev_err_out,
                       int32_t *iters_out, int32_t *status_out);

/* ---- rotation-search scoring -------------------------------------------------
 * The pre-alignment sweeps in front of every ICP call:
 *   utilities/features.py:165-242  rotation_search  (score every angle of a coarse, then a fine sweep)
 *   slam.py:111-183                _submap_rotation_search (the same around a predicted pose, plus one
 *                                  nearest-neighbour translation step, slam.py:166-181)
 * Both evaluate  score(angle) = mean_i min_j |R(angle) s_i + shift - t_j|^2  (features.py:205-211,
 * slam.py:138-143: KDTree distance, squared again, np.mean).  Problems are independent and packed like
 * the clouds of icpb200_icp_pairs: src (sum n_s, 2), tgt (sum n_t, 2), angles (radians) and one shift
 * (2 doubles) per problem; scores_out receives one value per angle in the order given.  With
 * nn_dist_out / nn_idx_out (both or neither; exactly one angle per problem) the exact nearest target
 * index and distance of every source point are returned as tgt_tree.query would (slam.py:168).
 * 2-D.  Targets of more than 8192 points are swept in slices of 8192 (one launch
 * each; the running nearest neighbours of every angle live in HBM). */
int icpb200_rotation_scores(int n_problems, const double *src, const int64_t *src_off,
                            const double *tgt, const int64_t *tgt_off, const double *angles,
                            const int64_t *ang_off, const double *shift, double *scores_out,
                            double *nn_dist_out, int32_t *nn_idx_out);

/* ---- pose graph --------------------------------
 * Replaces PoseGraph2D.optimize (utilities/pose_graph.py:83-134): Gauss-Newton
 * on SE(2) over `n_nodes` poses [x, y, theta] (row-major, updated in place) and
 * `n_edges` relative-pose constraints edge_i[e] -> edge_j[e] with measurement
 * meas[3e..] = [dx, dy, dtheta] and information matrix info[9e..] (row-major
 * 3 x 3).  The pose `fix_node` is anchored as the reference does it (1e10 on its
 * diagonal block, pose_graph.py:107-112).  Up to n_iterations steps; stops when
 * |step| < convergence_eps.  iters_out = the iteration index the loop ended on
 * (n_iterations at the limit: what the reference prints), step_norm_out = |step|
 * of the last iteration, status_out = ICPB200_CONVERGED, ICPB200_MAX_ITER or
 * ICPB200_SINGULAR.  The reference assembles a dense 3n x 3n matrix and calls
 * np.linalg.solve (O(n^3)); this is a block-skyline Cholesky on the HOST (cost
 * proportional to nodes + loop lengths) -- the one entry point that needs no CUDA
 * device (SURVEY: "a sparse Cholesky on host is the pragmatic fix").  Poses agree
 * with the reference to rounding (the linear solve differs), not bit for bit. */
int icpb200_pose_graph_optimize(int64_t n_nodes, double *poses, int64_t n_edges, const int32_t *edge_i,
                                const int32_t *edge_j, const double *meas, const double *info,
                                int n_iterations, int fix_node, double convergence_eps,
                                int32_t *iters_out, double *step_norm_out, int32_t *status_out);

/* ---- host buffers ----------------------------------------------------------
 * The host-buffer entry points above accept any host pointer.  Buffers that
 * are reused from call to call (a scan history, the caller's copy of the map)
 * transfer at full PCIe rate once they are page-locked; these two calls do
 * that for memory the CALLER owns (cudaHostRegister / cudaHostUnregister).
 * The caller keeps ownership and must unpin before freeing.  The reference
 * has no counterpart (numpy arrays in one address space, mapping.py:47). */
int icpb200_pin_host(void *ptr, size_t bytes);
int icpb200_unpin_host(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* ICP_B200_H */
