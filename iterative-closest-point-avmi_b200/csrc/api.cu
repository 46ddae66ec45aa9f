// C ABI of libicp_b200.so (declared in include/icp_b200.h): argument checking,
// host <-> device staging, workspace management and kernel launches.  No
// compute happens on the host and there is no CPU fallback: every compute
// entry point fails with ICPB200_ERR_CUDA when no CUDA device is usable.
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "icp_b200.h"
#include "common.cuh"
#include "icp_kernel.h"
#include "occupancy.h"

namespace icpb {

std::mutex g_api_mutex;
long long g_launches = 0;
static char g_error[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 4 + 256;            // grow geometrically
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        e = cudaMalloc(&p, bytes);                    // exact size as a second try
        want = bytes;
    }
    if (e != cudaSuccess) {
        p = nullptr;
        set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return ICPB200_ERR_CUDA;
    }
    cap = want;
    return 0;
}

void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

static Context g_ctx;
Context& ctx() { return g_ctx; }

static int init_locked(int device) {
    Context& c = g_ctx;
    if (c.ready && (device < 0 || device == c.device)) {
        ICPB_CUDA(cudaSetDevice(c.device));
        return ICPB200_OK;
    }
    if (c.ready) {
        set_error("icpb200_init: already bound to device %d (one process drives one GPU)", c.device);
        return ICPB200_ERR_ARG;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no usable CUDA device (%s); libicp_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return ICPB200_ERR_CUDA;
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= count) {
        set_error("icpb200_init: device %d requested but only %d visible", device, count);
        return ICPB200_ERR_ARG;
    }
    ICPB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    ICPB_CUDA(cudaGetDeviceProperties(&prop, device));
    c.device = device;
    c.sm_count = prop.multiProcessorCount;
    c.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    ICPB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) ICPB_CUDA(cudaEventCreate(&c.ev[i]));
    ICPB_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < kUploadChunks; ++i) {
        ICPB_CUDA(cudaEventCreateWithFlags(&c.chunk_ev[i], cudaEventDisableTiming));
        ICPB_CUDA(cudaEventCreateWithFlags(&c.chunk_done[i], cudaEventDisableTiming));
        ICPB_CUDA(cudaEventCreateWithFlags(&c.group_done[i], cudaEventDisableTiming));
        ICPB_CUDA(cudaStreamCreateWithFlags(&c.chunk_stream[i], cudaStreamNonBlocking));
    }
    ICPB_CUDA(cudaEventCreateWithFlags(&c.fork_ev, cudaEventDisableTiming));
    ICPB_CUDA(cudaEventCreateWithFlags(&c.icp_done, cudaEventDisableTiming));
    c.ready = true;
    return ICPB200_OK;
}


static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

struct IcpCommon {
    int dim;
    double error_threshold;
    int max_iterations;
    double voxel_size;
    int method, normal_k;
    double max_corr_dist;
    int nn_mode;
};

static int check_common(const IcpCommon& k, const char* who) {
    if (k.dim != 2 && k.dim != 3) { set_error("%s: dim must be 2 or 3 (got %d)", who, k.dim); return ICPB200_ERR_ARG; }
    if (!(k.voxel_size > 0.0)) { set_error("%s: voxel_size must be > 0", who); return ICPB200_ERR_ARG; }
    if (k.max_iterations < 0) { set_error("%s: max_iterations must be >= 0", who); return ICPB200_ERR_ARG; }
    if (k.method != ICPB200_POINT_TO_POINT && k.method != ICPB200_POINT_TO_LINE) {
        set_error("%s: unknown method %d", who, k.method); return ICPB200_ERR_ARG;
    }
    if (k.method == ICPB200_POINT_TO_LINE && k.dim == 2) {
        if (k.normal_k < 1) { set_error("%s: normal_k must be >= 1 for point_to_line", who); return ICPB200_ERR_ARG; }
        if (k.normal_k > 63) { set_error("%s: normal_k > 63 is not supported by this build", who); return ICPB200_ERR_LIMIT; }
    }
    if (k.nn_mode < 0 || k.nn_mode > 2) { set_error("%s: unknown nn_mode %d", who, k.nn_mode); return ICPB200_ERR_ARG; }
    return ICPB200_OK;
}

struct DevClouds {                 // one set of raw clouds resident on the device
    const double* pts;
    const long long* off;          // device
    const int64_t* h_off;          // the same offsets on the host (nullptr: not available)
    int n_clouds;
    long long role_max;            // largest raw cloud referenced in this role (source / target)
    long long set_max;             // largest raw cloud of the whole set
    long long total_points;        // upper bound on off[n_clouds]
};

struct IcpTrace {                  // optional debug outputs of a single-pair call
    int* match = nullptr;
    int iters = 0;
    int stride = 0;
};

// Preprocessed form of ONE target cloud that outlives a call: its voxel means, box, normals and hash grid live in buffers
// of their own (a device-resident submap owns one, icpb200_submap_*), so registrations against the same target skip the
// target's preprocessing altogether.
struct PrepTarget {
    DevBuf ds, n, box, nrm, flags, off;
    DevBuf grid_start, grid_items, grid_cell, grid_desc, grid_off, grid_buckets;
    bool ready = false;            // ds / box / (nrm) / (grid) hold the preprocessing for the parameters below
    double voxel = 0.0;
    int want_normals = 0, normal_k = 0, has_grid = 0;
    void release() {
        DevBuf* b[] = {&ds, &n, &box, &nrm, &flags, &off, &grid_start, &grid_items, &grid_cell, &grid_desc, &grid_off, &grid_buckets};
        for (DevBuf* p : b) p->release();
        ready = false;
    }
};

static int make_prep_cloud_set(PrepTarget& pt, const struct DevClouds& d, int dim, bool want_normals, CloudSet* out);

// Device buffers for the preprocessed form of cloud set `slot` (0 or 1).
static int make_cloud_set(Context& c, int slot, const DevClouds& d, int dim, bool want_normals, CloudSet* out) {
    const size_t np = (size_t)d.total_points, nc = (size_t)d.n_clouds;
    if (c.aux_ds[slot].reserve(sizeof(double) * dim * np) || c.aux_n[slot].reserve(sizeof(int) * nc) ||
        c.aux_box[slot].reserve(sizeof(double) * 6 * nc) || c.aux_flags[slot].reserve(2 * nc) ||
        (want_normals && c.aux_nrm[slot].reserve(sizeof(double) * 2 * np)))
        return ICPB200_ERR_CUDA;
    out->raw = d.pts; out->off = d.off; out->n_clouds = d.n_clouds;
    out->ds = c.aux_ds[slot].as<double>();
    out->ds_n = c.aux_n[slot].as<int>();
    out->box = c.aux_box[slot].as<double>();
    out->nrm = want_normals ? c.aux_nrm[slot].as<double>() : nullptr;
    out->used = c.aux_flags[slot].as<unsigned char>();
    out->is_tgt = out->used + nc;
    return ICPB200_OK;
}

static int make_prep_cloud_set(PrepTarget& pt, const DevClouds& d, int dim, bool want_normals, CloudSet* out) {
    const size_t np = (size_t)d.total_points;
    if (pt.ds.reserve(sizeof(double) * dim * np) || pt.n.reserve(sizeof(int)) || pt.box.reserve(sizeof(double) * 6) ||
        pt.flags.reserve(2) || (want_normals && pt.nrm.reserve(sizeof(double) * 2 * np)))
        return ICPB200_ERR_CUDA;
    out->raw = d.pts; out->off = d.off; out->n_clouds = 1;
    out->ds = pt.ds.as<double>();
    out->ds_n = pt.n.as<int>();
    out->box = pt.box.as<double>();
    out->nrm = want_normals ? pt.nrm.as<double>() : nullptr;
    out->used = pt.flags.as<unsigned char>();
    out->is_tgt = out->used + 1;
    return ICPB200_OK;
}

// voxel-grid means of a whole set: shared-memory kernel for scan-sized clouds,
// global-memory radix-sort kernel when any cloud of the set exceeds one CTA
// The host-buffer entry point uploads the clouds in a few chunks on a second stream; K1 (and K2) of a chunk start as
// soon as that chunk has arrived, under the copy of the next one.
struct UploadPlan {
    int n_chunks;
    int first[kUploadChunks + 1];              // cloud ranges
    // pairs grouped by the last chunk they need (group g: both clouds have arrived once chunk g has): the bulk launch of
    // group g runs on chunk g's stream right after its K1 / K2, under the uploads and per-cloud kernels of later chunks
    int group_first[kUploadChunks + 1];        // ranges of the ordered pair list
    const int* d_order;                        // device: ordered pair list, nullptr = the pairs are grouped as they come
};

static int voxel_set(Context& c, const CloudSet& cs, const DevClouds& d, int dim, double voxel, cudaStream_t st) {
    if (d.set_max <= ICPB200_BRUTE_MAX_POINTS)
        return launch_voxel_clouds(cs, dim, voxel, next_pow2((int)std::max<long long>(d.set_max, 256)), st);
    const size_t np = (size_t)d.total_points;
    if (c.big_keys.reserve(sizeof(unsigned long long) * 2 * np) || c.big_idx.reserve(sizeof(unsigned) * 2 * np))
        return ICPB200_ERR_CUDA;
    return launch_big_voxel(cs, dim, voxel, c.big_keys.as<unsigned long long>(), c.big_idx.as<unsigned>(),
                            d.total_points, st);
}

// Enqueue the registration of n_pairs pairs on `st` (device pointers everywhere):
// K0 mark -> K1 voxel means per referenced cloud -> [K4 hash grids] -> K2 normals per p2l target -> K3 pairs.
static int icp_enqueue(const IcpCommon& k, int n_pairs, const DevClouds& s, const DevClouds& t, bool same_set,
                       const int* d_src_idx, const int* d_tgt_idx,
                       const double* d_R_init, const double* d_t_init, double* d_R, double* d_t,
                       double* d_err, double* d_prev, int* d_iters, int* d_status, cudaStream_t st,
                       const IcpTrace& tr, IcpArgs* args_out, const UploadPlan* plan = nullptr, PrepTarget* prep = nullptr) {
    Context& c = g_ctx;
    if (n_pairs == 0) return ICPB200_OK;
    // The workspaces below (preprocessed clouds, queues, hand-over state) are per process: a call enqueued on another
    // stream while the previous one is still running must not touch them, so every enqueue waits for the previous one.
    if (c.icp_done_valid) ICPB_CUDA(cudaStreamWaitEvent(st, c.icp_done, 0));
    const bool grid = k.nn_mode == ICPB200_NN_GRID ||
                      (k.nn_mode == ICPB200_NN_AUTO && t.role_max > ICPB200_BRUTE_MAX_POINTS);
    if (!grid && t.role_max > ICPB200_BRUTE_MAX_POINTS) {
        set_error("icp: nn_mode brute supports targets of at most %d raw points (got %lld)", ICPB200_BRUTE_MAX_POINTS, t.role_max);
        return ICPB200_ERR_LIMIT;
    }
    if (s.role_max > ICPB200_BRUTE_MAX_POINTS) {
        set_error("icp: source clouds of more than %d raw points are not supported (got %lld)", ICPB200_BRUTE_MAX_POINTS, s.role_max);
        return ICPB200_ERR_LIMIT;
    }
    if (grid && k.dim != 2) {
        set_error("icp: the grid nearest-neighbour path is 2-D only in this build");
        return ICPB200_ERR_LIMIT;
    }
    const bool p2l = k.method == ICPB200_POINT_TO_LINE && k.dim == 2;
    IcpArgs a;
    memset(&a, 0, sizeof(a));
    int rc;
    if ((rc = make_cloud_set(c, 0, s, k.dim, p2l && same_set, &a.s))) return rc;
    // a target that keeps its preprocessing between calls (prep): reuse it when it was made with the same parameters
    const bool prep_hit = prep && prep->ready && prep->voxel == k.voxel_size && prep->want_normals == (p2l ? 1 : 0) &&
                          (!p2l || prep->normal_k == k.normal_k) && prep->has_grid == (grid ? 1 : 0);
    if (same_set) a.t = a.s;
    else if (prep) { if ((rc = make_prep_cloud_set(*prep, t, k.dim, p2l, &a.t))) return rc; }
    else if ((rc = make_cloud_set(c, 1, t, k.dim, p2l, &a.t))) return rc;
    a.n_pairs = n_pairs;
    a.src_idx = d_src_idx; a.tgt_idx = d_tgt_idx;
    a.R_init = d_R_init; a.t_init = d_t_init;
    a.err_thr = k.error_threshold; a.max_iter = k.max_iterations; a.voxel = k.voxel_size;
    a.method = k.method; a.normal_k = k.normal_k; a.max_corr = k.max_corr_dist;
    a.R_out = d_R; a.t_out = d_t; a.err_out = d_err; a.prev_out = d_prev; a.iters_out = d_iters; a.status_out = d_status;
    a.cap_s = round_up((int)std::max<long long>(s.role_max, 32), 32);
    a.cap_t = grid ? 32 : round_up((int)std::max<long long>(t.role_max, 32), 32);
    a.sort_pad = next_pow2((int)std::max<long long>(std::min<long long>(std::max(s.set_max, t.set_max), ICPB200_BRUTE_MAX_POINTS), 256));
    const size_t smem = icp_pair_smem_bytes(k.dim, a.cap_s, a.cap_t);               // bulk variant; the launcher adds what the roomy one needs
    if (icp_pair_smem_bytes(k.dim, a.cap_s, a.cap_t, 512) > (size_t)c.max_smem_optin || (!grid && p2l && icp_normals_smem_bytes(a.cap_t) > (size_t)c.max_smem_optin) ||
        icp_voxel_smem_bytes(a.sort_pad) > (size_t)c.max_smem_optin) {
        set_error("icp: %zu bytes of shared memory needed, device allows %d", smem, c.max_smem_optin);
        return ICPB200_ERR_LIMIT;
    }
    if (c.queue.reserve(16 * sizeof(unsigned)) || c.stats.reserve(24 * sizeof(unsigned long long))) return ICPB200_ERR_CUDA;
    a.queue = c.queue.as<unsigned>();
    // two-phase schedule (see icp_kernel.h): worthwhile once the batch fills the machine
    const int kPhaseCap = 12;                                    // flat between 8 and 16 on C2 (profiles/README.md)
    const bool two_phase = n_pairs >= 2 * c.sm_count && k.max_iterations > 2 * kPhaseCap;
    a.phase_cap = two_phase ? kPhaseCap : 0;
    static const bool no_slab = getenv("ICPB200_NO_SLAB") != nullptr;      // A/B switch: tile sweep for every decision
    a.brute_slab = (!grid && k.voxel_size > 0.0 && !no_slab) ? 1 : 0;
    static const bool no_tma = getenv("ICPB200_NO_TMA") != nullptr;       // A/B switch (profiles/README.md)
    a.tma_normals = no_tma ? 0 : 1;
    a.resume = 0;
    if (two_phase) {
        const size_t np_ = (size_t)n_pairs, cs_ = (size_t)a.cap_s;
        if (c.cont_cur.reserve(sizeof(double) * k.dim * cs_ * np_) || c.cont_match.reserve(sizeof(int) * cs_ * np_) ||
            c.cont_d2lb.reserve(sizeof(float) * cs_ * np_) || c.cont_moved.reserve(sizeof(float) * k.dim * cs_ * np_) ||
            c.cont_scalar.reserve(sizeof(double) * 16 * np_) || c.cont_list.reserve(sizeof(int) * 5 * np_))
            return ICPB200_ERR_CUDA;
        a.cont_count = a.queue + 8;                                   // [8] total, [9..12] per cost class
        a.cont_list = c.cont_list.as<int>();
        a.cont_bucket = a.cont_list + np_;
        a.cont_cap = n_pairs;
        a.cont_cur = c.cont_cur.as<double>();
        a.cont_match = c.cont_match.as<int>();
        a.cont_d2lb = c.cont_d2lb.as<float>();
        a.cont_moved = c.cont_moved.as<float>();
        a.cont_scalar = c.cont_scalar.as<double>();
    }
    a.stats = c.stats.as<unsigned long long>();
    a.trace_match = tr.match; a.trace_iters = tr.iters; a.trace_stride = tr.stride;
    a.pair_prof = nullptr;
    if (c.pair_prof_on) {                                          // icpb200_icp_pair_profile() asked for per-pair counters
        if (c.pair_prof.reserve(sizeof(unsigned long long) * 8 * (size_t)n_pairs)) return ICPB200_ERR_CUDA;
        ICPB_CUDA(cudaMemsetAsync(c.pair_prof.p, 0, sizeof(unsigned long long) * 8 * (size_t)n_pairs, st));
        a.pair_prof = c.pair_prof.as<unsigned long long>();
        c.pair_prof_n = n_pairs;
    }
    ICPB_CUDA(cudaMemsetAsync(a.queue, 0, 16 * sizeof(unsigned), st));
    ICPB_CUDA(cudaMemsetAsync(a.stats, 0, 24 * sizeof(unsigned long long), st));
    ICPB_CUDA(cudaMemsetAsync(a.s.used, 0, 2 * (size_t)s.n_clouds, st));
    if (!same_set && !prep_hit) ICPB_CUDA(cudaMemsetAsync(a.t.used, 0, 2 * (size_t)t.n_clouds, st));
    a.grids = nullptr;
    if ((rc = launch_mark_used(a, p2l || grid, st))) return rc;
    ICPB_CUDA(cudaEventRecord(c.ev[0], st));
    const bool chunked = plan && same_set && !grid && s.set_max <= ICPB200_BRUTE_MAX_POINTS;
    bool bulk_done = false;                        // chunked: the bulk launches went out per upload chunk
    if (plan && !chunked)
        for (int ch = 0; ch < plan->n_chunks; ++ch) ICPB_CUDA(cudaStreamWaitEvent(st, c.chunk_ev[ch], 0));
    if (chunked) {
        // K1 (and K2) of an upload chunk run under the copy of the next chunk
        const int sort_pad = next_pow2((int)std::max<long long>(s.set_max, 256));
        // each chunk's kernels go to their own stream: a quarter of the clouds does not fill the GPU, so the
        // chunks' kernels must be able to run side by side (and under the copies still in flight)
        ICPB_CUDA(cudaEventRecord(c.fork_ev, st));                 // flags from mark_used
        const int per_sm_g = icp_max_ctas_per_sm(k.dim, grid, smem);
        for (int ch = 0; ch < plan->n_chunks; ++ch) {
            const int first = plan->first[ch], count = plan->first[ch + 1] - first;
            cudaStream_t cs_ = c.chunk_stream[ch];
            ICPB_CUDA(cudaStreamWaitEvent(cs_, c.fork_ev, 0));
            ICPB_CUDA(cudaStreamWaitEvent(cs_, c.chunk_ev[ch], 0));
            if ((rc = launch_voxel_clouds(a.s, k.dim, k.voxel_size, sort_pad, cs_, first, count))) return rc;
            if (p2l && (rc = launch_normals(a.t, a.cap_t, k.normal_k, k.voxel_size, cs_, first, count))) return rc;
            ICPB_CUDA(cudaEventRecord(c.chunk_done[ch], cs_));
            // the pairs whose clouds are complete with this chunk: bulk launch on the chunk's stream
            const int g_pairs = plan->group_first[ch + 1] - plan->group_first[ch];
            if (g_pairs > 0) {
                for (int prev = 0; prev < ch; ++prev) ICPB_CUDA(cudaStreamWaitEvent(cs_, c.chunk_done[prev], 0));
                IcpArgs ga = a;
                ga.n_pairs = g_pairs;
                ga.pair_first = plan->group_first[ch];
                ga.pair_order = plan->d_order;
                ga.queue = a.queue + 4 + ch;
                if ((rc = launch_icp_pairs(ga, k.dim, grid, std::min(g_pairs, c.sm_count * per_sm_g), smem, cs_))) return rc;
            }
            ICPB_CUDA(cudaEventRecord(c.group_done[ch], cs_));
            ICPB_CUDA(cudaStreamWaitEvent(st, c.group_done[ch], 0));
        }
        bulk_done = true;
        ICPB_CUDA(cudaEventRecord(c.ev[1], st));
    } else {
        if ((rc = voxel_set(c, a.s, s, k.dim, k.voxel_size, st))) return rc;
        if (!same_set && !prep_hit && (rc = voxel_set(c, a.t, t, k.dim, k.voxel_size, st))) return rc;
        ICPB_CUDA(cudaEventRecord(c.ev[1], st));
        // the hash grids live in the context's buffers, or in the prepared target's own
        DevBuf& g_start = prep ? prep->grid_start : c.grid_start;
        DevBuf& g_items = prep ? prep->grid_items : c.grid_items;
        DevBuf& g_cell = prep ? prep->grid_cell : c.grid_cell;
        DevBuf& g_desc = prep ? prep->grid_desc : c.grid_desc;
        DevBuf& g_off = prep ? prep->grid_off : c.grid_off;
        DevBuf& g_buckets = prep ? prep->grid_buckets : c.grid_buckets;
        if (prep_hit) {
            if (grid) a.grids = g_desc.as<BigGrid>();
        } else if (grid) {
            // one hash grid per target cloud; bucket counts from the raw sizes (host side)
            std::vector<int64_t> fetched;
            const int64_t* h_off = t.h_off;
            if (!h_off) {
                fetched.resize((size_t)t.n_clouds + 1);
                ICPB_CUDA(cudaMemcpyAsync(fetched.data(), t.off, sizeof(int64_t) * fetched.size(), cudaMemcpyDeviceToHost, st));
                ICPB_CUDA(cudaStreamSynchronize(st));
                h_off = fetched.data();
            }
            std::vector<long long> goff((size_t)t.n_clouds);
            std::vector<int> gbuckets((size_t)t.n_clouds);
            long long total_start = 0;
            for (int i = 0; i < t.n_clouds; ++i) {
                const long long n = h_off[i + 1] - h_off[i];
                int b = 1024;
                while (b < 2 * n && b < (1 << 24)) b <<= 1;
                gbuckets[i] = b;
                goff[i] = total_start;
                total_start += b + 1;
            }
            const size_t np = (size_t)t.total_points, nc = (size_t)t.n_clouds;
            if (g_start.reserve(sizeof(int) * (size_t)total_start) || g_items.reserve(sizeof(int) * np) ||
                g_cell.reserve(sizeof(int2) * np) || g_desc.reserve(sizeof(BigGrid) * nc) ||
                g_off.reserve(sizeof(long long) * nc) || g_buckets.reserve(sizeof(int) * nc))
                return ICPB200_ERR_CUDA;
            ICPB_CUDA(cudaMemcpyAsync(g_off.p, goff.data(), sizeof(long long) * nc, cudaMemcpyHostToDevice, st));
            ICPB_CUDA(cudaMemcpyAsync(g_buckets.p, gbuckets.data(), sizeof(int) * nc, cudaMemcpyHostToDevice, st));
            ICPB_CUDA(cudaStreamSynchronize(st));          // goff / gbuckets are stack-owned
            ICPB_CUDA(cudaMemsetAsync(g_desc.p, 0, sizeof(BigGrid) * nc, st));
            // cell edge: 8 voxels (a 3x3 block of cells holds a few dozen wall points)
            if ((rc = launch_big_grid(a.t, 8.0 * k.voxel_size, g_off.as<long long>(), g_buckets.as<int>(),
                                      g_start.as<int>(), g_items.as<int>(), g_cell.as<int2>(),
                                      g_desc.as<BigGrid>(), st)))
                return rc;
            a.grids = g_desc.as<BigGrid>();
            if (p2l && (rc = launch_big_normals(a.t, a.grids, k.normal_k, t.role_max, st))) return rc;
        } else if (p2l) {
            if ((rc = launch_normals(a.t, a.cap_t, k.normal_k, k.voxel_size, st))) return rc;
        }
        if (prep && !prep_hit) {
            prep->ready = true; prep->voxel = k.voxel_size; prep->want_normals = p2l ? 1 : 0; prep->normal_k = k.normal_k;
            prep->has_grid = grid ? 1 : 0;
        }
    }
    ICPB_CUDA(cudaEventRecord(c.ev[2], st));
    const int per_sm = icp_max_ctas_per_sm(k.dim, grid, smem);
    const int n_ctas = std::min(n_pairs, c.sm_count * per_sm);
    c.last_icp_stream = st;
    if (args_out) *args_out = a;
    if (!bulk_done && (rc = launch_icp_pairs(a, k.dim, grid, n_ctas, smem, st))) return rc;
    if (two_phase) {
        // the handed-over pairs: one CTA per SM (shared memory request above half an SM forces it)
        IcpArgs b = a;
        b.resume = 1;
        b.queue = a.queue + 1;
        // (two or three CTAs per SM were measured slower: the hand-over pairs are latency-bound and want the SM alone)
        const size_t smem2 = std::min(std::max(smem, (size_t)c.max_smem_optin / 2 + 1024), (size_t)c.max_smem_optin);
        static const bool no_cluster = getenv("ICPB200_NO_CLUSTER") != nullptr;      // A/B switch (profiles/README.md)
        static const int coop_factor = getenv("ICPB200_COOP_FACTOR") ? atoi(getenv("ICPB200_COOP_FACTOR")) : 10;     // tuning switch (profiles/README.md)
        b.coop_factor = coop_factor;
        b.class_lo = 0; b.class_hi = 4;                                  // every cost class (IcpArgs::cont_bucket), expensive first
        if (icp_cluster_variant(k.dim, grid) && !no_cluster) {
            // two launches side by side, clusters of CTAs (a CTA that runs out of pairs helps a cluster mate with its
            // sweeps) and plain CTAs; the kernels decide from the handed-over lists which of them does the work
            // (coop_plan in icp_kernel.cu: clusters when a few long chains bound the launch, plain CTAs otherwise)
            int n_cl = 0;
            cudaStream_t side = c.chunk_stream[0];
            ICPB_CUDA(cudaEventRecord(c.fork_ev, st));                      // the bulk launch and its hand-over lists
            if ((rc = launch_icp_pairs_cluster(b, c.sm_count, smem2, st, &n_cl))) return rc;
            b.coop_ctas = n_cl;
            ICPB_CUDA(cudaStreamWaitEvent(side, c.fork_ev, 0));
            if ((rc = launch_icp_pairs(b, k.dim, grid, c.sm_count, smem2, side))) return rc;
            ICPB_CUDA(cudaEventRecord(c.group_done[0], side));
            ICPB_CUDA(cudaStreamWaitEvent(st, c.group_done[0], 0));
        } else {
            if ((rc = launch_icp_pairs(b, k.dim, grid, c.sm_count, smem2, st))) return rc;
        }
    }
    ICPB_CUDA(cudaEventRecord(c.ev[3], st));
    ICPB_CUDA(cudaEventRecord(c.icp_done, st));
    c.icp_done_valid = true;
    return ICPB200_OK;
}

static int check_offsets(const int64_t* off, int n, const char* what, long long* max_points) {
    long long mx = 0;
    if (off[0] != 0) { set_error("%s[0] must be 0", what); return ICPB200_ERR_ARG; }
    for (int i = 0; i < n; ++i) {
        const long long len = off[i + 1] - off[i];
        if (len <= 0) { set_error("%s: cloud %d is empty or offsets decrease", what, i); return ICPB200_ERR_ARG; }
        mx = std::max(mx, len);
    }
    *max_points = mx;
    return ICPB200_OK;
}

struct IcpOutputs { double* R; double* t; double* err; double* prev; int32_t* iters; int32_t* status; };

static int fetch_outputs(Context& c, int n_pairs, int dim, const IcpOutputs& o) {
    // The caller's arrays are pageable: six device-to-pageable copies would each go through the driver's bounce buffer
    // and block.  The results land in one page-locked staging block instead (asynchronous, one wait) and are copied
    // out by the host.
    const size_t np = (size_t)n_pairs;
    const size_t bytes[6] = {sizeof(double) * dim * dim * np, sizeof(double) * dim * np, sizeof(double) * np,
                             o.prev ? sizeof(double) * np : 0, sizeof(int) * np, sizeof(int) * np};
    const void* src[6] = {c.out_r.p, c.out_t.p, c.out_err.p, c.out_prev.p, c.out_iters.p, c.out_status.p};
    void* dst[6] = {o.R, o.t, o.err, o.prev, o.iters, o.status};
    size_t at[6], total = 0;
    for (int k = 0; k < 6; ++k) { at[k] = total; total += (bytes[k] + 63) & ~(size_t)63; }
    if (total > c.h_stage_cap) {
        if (c.h_stage) { cudaFreeHost(c.h_stage); c.h_stage = nullptr; c.h_stage_cap = 0; }
        const size_t cap = std::max<size_t>(total + total / 2, 1u << 16);
        ICPB_CUDA(cudaHostAlloc(&c.h_stage, cap, cudaHostAllocDefault));
        c.h_stage_cap = cap;
    }
    unsigned char* stage = static_cast<unsigned char*>(c.h_stage);
    for (int k = 0; k < 6; ++k)
        if (bytes[k]) ICPB_CUDA(cudaMemcpyAsync(stage + at[k], src[k], bytes[k], cudaMemcpyDeviceToHost, c.stream));
    ICPB_CUDA(cudaStreamSynchronize(c.stream));
    for (int k = 0; k < 6; ++k)
        if (bytes[k]) memcpy(dst[k], stage + at[k], bytes[k]);
    return ICPB200_OK;
}

static int reserve_outputs(Context& c, int n_pairs, int dim) {
    if (c.out_r.reserve(sizeof(double) * dim * dim * (size_t)n_pairs) || c.out_t.reserve(sizeof(double) * dim * (size_t)n_pairs) ||
        c.out_err.reserve(sizeof(double) * (size_t)n_pairs) || c.out_prev.reserve(sizeof(double) * (size_t)n_pairs) || c.out_iters.reserve(sizeof(int) * (size_t)n_pairs) ||
        c.out_status.reserve(sizeof(int) * (size_t)n_pairs))
        return ICPB200_ERR_CUDA;
    return ICPB200_OK;
}

static int upload_init(Context& c, int n_pairs, int dim, const double* R_init, const double* t_init,
                       const double** d_R, const double** d_t) {
    *d_R = nullptr; *d_t = nullptr;
    if (R_init && t_init) {                         // icp.py:153: both or neither
        if (c.rinit.reserve(sizeof(double) * dim * dim * (size_t)n_pairs) || c.tinit.reserve(sizeof(double) * dim * (size_t)n_pairs))
            return ICPB200_ERR_CUDA;
        ICPB_CUDA(cudaMemcpyAsync(c.rinit.p, R_init, sizeof(double) * dim * dim * (size_t)n_pairs, cudaMemcpyHostToDevice, c.stream));
        ICPB_CUDA(cudaMemcpyAsync(c.tinit.p, t_init, sizeof(double) * dim * (size_t)n_pairs, cudaMemcpyHostToDevice, c.stream));
        *d_R = c.rinit.as<double>();
        *d_t = c.tinit.as<double>();
    }
    return ICPB200_OK;
}

// ---- device-resident submap (slam.py:103-108, 559-562, 611-615) ---------------------------------------------------------
// The rolling window of global-frame scans the reference keeps as a Python list (`submap_buffer`): here the scans live in
// fixed slots of one device arena, a push uploads ONE scan (17 KB), and everything the reference recomputes from the
// whole window in every scan -- np.vstack, the first voxel_downsample (`_build_submap`), and inside ICP() the second
// voxel_downsample of the target plus its KD-tree (here: hash grid) -- is computed on the device from the resident window
// and cached until the window changes.
struct SubmapSlot { int slot; long long n; };

struct Submap {
    int dim = 2, capacity = 0;
    long long slot_points = 4096;                  // points per slot (grows when a larger scan arrives)
    DevBuf arena;                                  // capacity x slot_points x dim doubles
    std::vector<SubmapSlot> scans;                 // window, oldest first
    std::vector<int> free_slots;
    long long version = 0;                         // bumped by every push / clear
    // np.vstack of the window (device), then voxel_downsample(.., submap_voxel): `built`
    DevBuf cat, cat_desc, cat_off, built, built_n, built_box, built_off;
    long long built_version = -1, built_points = 0;
    double built_voxel = 0.0;
    PrepTarget prep;                               // ICP()'s own preprocessing of that cloud (second downsample + grid)
    long long prep_version = -1;
    void release() {
        DevBuf* b[] = {&arena, &cat, &cat_desc, &cat_off, &built, &built_n, &built_box, &built_off};
        for (DevBuf* p : b) p->release();
        prep.release();
    }
};

__global__ void submap_concat_kernel(const double* __restrict__ arena, const long long* __restrict__ desc /* [scan] = {src row, dst row, rows} */,
                                     int n_scans, int dim, double* __restrict__ out) {
    const int s = blockIdx.y;
    if (s >= n_scans) return;
    const long long src = desc[3 * s] * dim, dst = desc[3 * s + 1] * dim, n = desc[3 * s + 2] * dim;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[dst + i] = arena[src + i];
}

// vstack + voxel_downsample(window, submap_voxel) on the device (slam.py:103-108); *n_out = rows of the result
static int submap_build(Context& c, Submap& sm, double voxel, long long* n_out, cudaStream_t st) {
    if (sm.built_version == sm.version && sm.built_voxel == voxel) { *n_out = sm.built_points; return ICPB200_OK; }
    long long total = 0;
    for (const SubmapSlot& s : sm.scans) total += s.n;
    *n_out = 0;
    sm.built_points = 0; sm.built_version = sm.version; sm.built_voxel = voxel;
    sm.prep.ready = false;
    if (total == 0) return ICPB200_OK;
    const int ns = (int)sm.scans.size();
    std::vector<long long> desc((size_t)3 * ns);
    long long at = 0;
    for (int i = 0; i < ns; ++i) {
        desc[3 * i] = (long long)sm.scans[i].slot * sm.slot_points; desc[3 * i + 1] = at; desc[3 * i + 2] = sm.scans[i].n;
        at += sm.scans[i].n;
    }
    const long long off[2] = {0, total};
    if (sm.cat.reserve(sizeof(double) * sm.dim * (size_t)total) || sm.cat_desc.reserve(sizeof(long long) * desc.size()) ||
        sm.cat_off.reserve(sizeof(off)) || sm.built.reserve(sizeof(double) * sm.dim * (size_t)total) ||
        sm.built_n.reserve(sizeof(int)) || sm.built_box.reserve(sizeof(double) * 6) || sm.built_off.reserve(sizeof(off)) ||
        c.big_keys.reserve(sizeof(unsigned long long) * 2 * (size_t)total) || c.big_idx.reserve(sizeof(unsigned) * 2 * (size_t)total))
        return ICPB200_ERR_CUDA;
    ICPB_CUDA(cudaMemcpyAsync(sm.cat_desc.p, desc.data(), sizeof(long long) * desc.size(), cudaMemcpyHostToDevice, st));
    ICPB_CUDA(cudaMemcpyAsync(sm.cat_off.p, off, sizeof(off), cudaMemcpyHostToDevice, st));
    ICPB_CUDA(cudaStreamSynchronize(st));              // desc / off are stack-owned
    submap_concat_kernel<<<dim3(8, (unsigned)ns), 256, 0, st>>>(sm.arena.as<double>(), sm.cat_desc.as<long long>(), ns, sm.dim, sm.cat.as<double>());
    ICPB_LAUNCH_CHECK();
    CloudSet cs;
    memset(&cs, 0, sizeof(cs));
    cs.raw = sm.cat.as<double>(); cs.off = sm.cat_off.as<long long>(); cs.n_clouds = 1;
    cs.ds = sm.built.as<double>(); cs.ds_n = sm.built_n.as<int>(); cs.box = sm.built_box.as<double>();
    int rc = launch_big_voxel(cs, sm.dim, voxel, c.big_keys.as<unsigned long long>(), c.big_idx.as<unsigned>(), total, st);
    if (rc) return rc;
    int m = 0;
    ICPB_CUDA(cudaMemcpyAsync(&m, cs.ds_n, sizeof(int), cudaMemcpyDeviceToHost, st));
    ICPB_CUDA(cudaStreamSynchronize(st));
    if (m < 0) { set_error("icpb200_submap: voxel index range does not fit 62 bits"); return ICPB200_ERR_LIMIT; }
    const long long boff[2] = {0, m};
    ICPB_CUDA(cudaMemcpyAsync(sm.built_off.p, boff, sizeof(boff), cudaMemcpyHostToDevice, st));
    ICPB_CUDA(cudaStreamSynchronize(st));
    sm.built_points = m;
    *n_out = m;
    return ICPB200_OK;
}

}  // namespace icpb

using namespace icpb;

extern "C" {

int icpb200_init(int device) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    return init_locked(device);
}

void icpb200_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    Context& c = g_ctx;
    if (!c.ready) return;
    cudaSetDevice(c.device);
    cudaStreamSynchronize(c.stream);
    DevBuf* bufs[] = {&c.pts_a, &c.pts_b, &c.off_a, &c.off_b, &c.idx_a, &c.idx_b, &c.idx_order, &c.rinit, &c.tinit, &c.out_r,
                      &c.out_t, &c.out_err, &c.out_prev, &c.out_iters, &c.out_status, &c.queue, &c.trace, &c.stats,
                      &c.aux_ds[0], &c.aux_ds[1], &c.aux_n[0], &c.aux_n[1], &c.aux_box[0], &c.aux_box[1],
                      &c.aux_nrm[0], &c.aux_nrm[1], &c.aux_flags[0], &c.aux_flags[1], &c.vox_in, &c.vox_out,
                      &c.big_keys, &c.big_idx, &c.grid_start, &c.grid_items, &c.grid_cell, &c.grid_desc, &c.grid_off,
                      &c.grid_buckets, &c.rot_src, &c.rot_tgt, &c.rot_ang, &c.rot_off, &c.rot_out, &c.cont_cur, &c.cont_match, &c.cont_d2lb, &c.cont_moved, &c.cont_scalar, &c.cont_list, &c.pair_prof, &c.rot_part};
    for (DevBuf* b : bufs) b->release();
    for (int i = 0; i < 4; ++i) if (c.ev[i]) { cudaEventDestroy(c.ev[i]); c.ev[i] = nullptr; }
    for (int i = 0; i < kUploadChunks; ++i) {
        if (c.chunk_ev[i]) { cudaEventDestroy(c.chunk_ev[i]); c.chunk_ev[i] = nullptr; }
        if (c.chunk_done[i]) { cudaEventDestroy(c.chunk_done[i]); c.chunk_done[i] = nullptr; }
        if (c.group_done[i]) { cudaEventDestroy(c.group_done[i]); c.group_done[i] = nullptr; }
        if (c.chunk_stream[i]) { cudaStreamDestroy(c.chunk_stream[i]); c.chunk_stream[i] = nullptr; }
    }
    if (c.fork_ev) { cudaEventDestroy(c.fork_ev); c.fork_ev = nullptr; }
    if (c.icp_done) { cudaEventDestroy(c.icp_done); c.icp_done = nullptr; c.icp_done_valid = false; }
    if (c.h_stage) { cudaFreeHost(c.h_stage); c.h_stage = nullptr; c.h_stage_cap = 0; }
    if (c.copy_stream) { cudaStreamDestroy(c.copy_stream); c.copy_stream = nullptr; }
    cudaStreamDestroy(c.stream);
    c.stream = nullptr;
    c.ready = false;
    c.device = -1;
}

const char* icpb200_last_error(void) { return g_error; }
int64_t icpb200_launch_count(void) { return g_launches; }
int icpb200_built_arch(void) { return 100; }

int icpb200_icp_batch(int n_pairs, int dim, const double* src, const int64_t* src_off, const double* tgt,
                      const int64_t* tgt_off, const double* R_init, const double* t_init, double error_threshold,
                      int max_iterations, double voxel_size, int method, int normal_k, double max_corr_dist,
                      int nn_mode, double* R_out, double* t_out, double* err_out, double* prev_err_out,
                      int32_t* iters_out, int32_t* status_out) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    const IcpCommon k{dim, error_threshold, max_iterations, voxel_size, method, normal_k, max_corr_dist, nn_mode};
    int rc = check_common(k, "icpb200_icp_batch");
    if (rc) return rc;
    if (n_pairs < 0 || (n_pairs > 0 && (!src || !src_off || !tgt || !tgt_off || !R_out || !t_out || !err_out || !iters_out || !status_out))) {
        set_error("icpb200_icp_batch: null pointer or negative n_pairs");
        return ICPB200_ERR_ARG;
    }
    if (n_pairs == 0) return ICPB200_OK;
    long long max_s, max_t;
    if ((rc = check_offsets(src_off, n_pairs, "src_off", &max_s))) return rc;
    if ((rc = check_offsets(tgt_off, n_pairs, "tgt_off", &max_t))) return rc;
    if ((rc = init_locked(-1))) return rc;
    Context& c = g_ctx;
    const size_t ns = (size_t)src_off[n_pairs], nt = (size_t)tgt_off[n_pairs];
    if (c.pts_a.reserve(sizeof(double) * dim * ns) || c.pts_b.reserve(sizeof(double) * dim * nt) ||
        c.off_a.reserve(sizeof(int64_t) * (n_pairs + 1)) || c.off_b.reserve(sizeof(int64_t) * (n_pairs + 1)))
        return ICPB200_ERR_CUDA;
    if ((rc = reserve_outputs(c, n_pairs, dim))) return rc;
    ICPB_CUDA(cudaMemcpyAsync(c.pts_a.p, src, sizeof(double) * dim * ns, cudaMemcpyHostToDevice, c.stream));
    ICPB_CUDA(cudaMemcpyAsync(c.pts_b.p, tgt, sizeof(double) * dim * nt, cudaMemcpyHostToDevice, c.stream));
    ICPB_CUDA(cudaMemcpyAsync(c.off_a.p, src_off, sizeof(int64_t) * (n_pairs + 1), cudaMemcpyHostToDevice, c.stream));
    ICPB_CUDA(cudaMemcpyAsync(c.off_b.p, tgt_off, sizeof(int64_t) * (n_pairs + 1), cudaMemcpyHostToDevice, c.stream));
    const double *d_Ri, *d_ti;
    if ((rc = upload_init(c, n_pairs, dim, R_init, t_init, &d_Ri, &d_ti))) return rc;
    const DevClouds s{c.pts_a.as<double>(), c.off_a.as<long long>(), src_off, n_pairs, max_s, max_s, (long long)ns};
    const DevClouds t{c.pts_b.as<double>(), c.off_b.as<long long>(), tgt_off, n_pairs, max_t, max_t, (long long)nt};
    rc = icp_enqueue(k, n_pairs, s, t, false, nullptr, nullptr, d_Ri, d_ti, c.out_r.as<double>(), c.out_t.as<double>(),
                     c.out_err.as<double>(), c.out_prev.as<double>(), c.out_iters.as<int>(), c.out_status.as<int>(),
                     c.stream, IcpTrace{}, nullptr);
    if (rc) return rc;
    return fetch_outputs(c, n_pairs, dim, IcpOutputs{R_out, t_out, err_out, prev_err_out, iters_out, status_out});
}

int icpb200_icp_pairs(int n_clouds, int dim, const double* pts, const int64_t* cloud_off, int n_pairs,
                      const int32_t* src_idx, const int32_t* tgt_idx, const double* R_init, const double* t_init,
                      double error_threshold, int max_iterations, double voxel_size, int method, int normal_k,
                      double max_corr_dist, int nn_mode, double* R_out, double* t_out, double* err_out,
                      double* prev_err_out, int32_t* iters_out, int32_t* status_out) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    const IcpCommon k{dim, error_threshold, max_iterations, voxel_size, method, normal_k, max_corr_dist, nn_mode};
    int rc = check_common(k, "icpb200_icp_pairs");
    if (rc) return rc;
    if (n_pairs < 0 || n_clouds < 0 || (n_pairs > 0 && (!pts || !cloud_off || !src_idx || !tgt_idx || !R_out || !t_out || !err_out || !iters_out || !status_out))) {
        set_error("icpb200_icp_pairs: null pointer or negative count");
        return ICPB200_ERR_ARG;
    }
    if (n_pairs == 0) return ICPB200_OK;
    long long max_pts;
    if ((rc = check_offsets(cloud_off, n_clouds, "cloud_off", &max_pts))) return rc;
    for (int p = 0; p < n_pairs; ++p)
        if (src_idx[p] < 0 || src_idx[p] >= n_clouds || tgt_idx[p] < 0 || tgt_idx[p] >= n_clouds) {
            set_error("icpb200_icp_pairs: pair %d references a cloud outside [0, %d)", p, n_clouds);
            return ICPB200_ERR_ARG;
        }
    if ((rc = init_locked(-1))) return rc;
    Context& c = g_ctx;
    const size_t np = (size_t)cloud_off[n_clouds];
    if (c.pts_a.reserve(sizeof(double) * dim * np) || c.off_a.reserve(sizeof(int64_t) * (n_clouds + 1)) ||
        c.idx_a.reserve(sizeof(int32_t) * (size_t)n_pairs) || c.idx_b.reserve(sizeof(int32_t) * (size_t)n_pairs))
        return ICPB200_ERR_CUDA;
    if ((rc = reserve_outputs(c, n_pairs, dim))) return rc;
    ICPB_CUDA(cudaMemcpyAsync(c.off_a.p, cloud_off, sizeof(int64_t) * (n_clouds + 1), cudaMemcpyHostToDevice, c.stream));
    ICPB_CUDA(cudaMemcpyAsync(c.idx_a.p, src_idx, sizeof(int32_t) * (size_t)n_pairs, cudaMemcpyHostToDevice, c.stream));
    ICPB_CUDA(cudaMemcpyAsync(c.idx_b.p, tgt_idx, sizeof(int32_t) * (size_t)n_pairs, cudaMemcpyHostToDevice, c.stream));
    // the clouds go up in a few chunks on the copy stream; the per-cloud kernels of a chunk start when it has landed
    static const bool e2e_timing = getenv("ICPB200_E2E_TIMING") != nullptr;
    static const bool no_chunk = getenv("ICPB200_NO_CHUNK") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    // Only the clouds some pair references cross PCIe (a loop-closure batch names a few scans of a long history,
    // slam.py:575-579; a rank of a sharded batch names its own block's): runs of referenced clouds are copied, gaps of
    // up to 64 KB are bridged rather than paid for with another copy call.  Chunk boundaries balance the bytes copied.
    std::vector<unsigned char> used((size_t)n_clouds, 0);
    for (int p = 0; p < n_pairs; ++p) used[(size_t)src_idx[p]] = used[(size_t)tgt_idx[p]] = 1;
    size_t used_rows = 0;
    for (int i = 0; i < n_clouds; ++i) if (used[(size_t)i]) used_rows += (size_t)(cloud_off[i + 1] - cloud_off[i]);
    UploadPlan plan;
    plan.n_chunks = (int)std::max<size_t>(1, std::min<size_t>(kUploadChunks, used_rows * dim * sizeof(double) / (4u << 20)));
    if (no_chunk) plan.n_chunks = 1;
    plan.n_chunks = std::min(plan.n_chunks, n_clouds);
    {
        plan.first[0] = 0;
        size_t seen = 0;
        int ch = 1;
        for (int i = 0; i < n_clouds && ch < plan.n_chunks; ++i) {
            if (used[(size_t)i]) seen += (size_t)(cloud_off[i + 1] - cloud_off[i]);
            if (seen * (size_t)plan.n_chunks >= used_rows * (size_t)ch) plan.first[ch++] = i + 1;
        }
        for (; ch <= plan.n_chunks; ++ch) plan.first[ch] = n_clouds;
    }
    const long long bridge_rows = (64 << 10) / (long long)(sizeof(double) * dim);
    for (int ch = 0; ch < plan.n_chunks; ++ch) {
        int i = plan.first[ch];
        const int end = plan.first[ch + 1];
        while (i < end) {
            while (i < end && !used[(size_t)i]) ++i;
            if (i >= end) break;
            int j = i + 1, last_used = i;                     // extend the run over small gaps
            while (j < end) {
                if (used[(size_t)j]) last_used = j;
                else if (cloud_off[j + 1] - cloud_off[last_used + 1] > bridge_rows) break;
                ++j;
            }
            const size_t b0 = (size_t)cloud_off[i] * dim, b1 = (size_t)cloud_off[last_used + 1] * dim;
            ICPB_CUDA(cudaMemcpyAsync(c.pts_a.as<double>() + b0, pts + b0, sizeof(double) * (b1 - b0), cudaMemcpyHostToDevice, c.copy_stream));
            i = last_used + 1;
        }
        ICPB_CUDA(cudaEventRecord(c.chunk_ev[ch], c.copy_stream));
    }
    const double *d_Ri, *d_ti;
    // from here on the caller's (possibly page-locked) buffer is being read by the copy stream: every error return waits for it
    if ((rc = upload_init(c, n_pairs, dim, R_init, t_init, &d_Ri, &d_ti))) { cudaStreamSynchronize(c.copy_stream); cudaStreamSynchronize(c.stream); return rc; }
    long long max_src = 0, max_tgt = 0;
    // pairs by the last upload chunk they need (see UploadPlan)
    std::vector<int> group((size_t)n_pairs);
    int count[kUploadChunks + 1] = {};
    bool grouped = true;
    for (int p = 0; p < n_pairs; ++p) {
        max_src = std::max<long long>(max_src, cloud_off[src_idx[p] + 1] - cloud_off[src_idx[p]]);
        max_tgt = std::max<long long>(max_tgt, cloud_off[tgt_idx[p] + 1] - cloud_off[tgt_idx[p]]);
        const int last = std::max(src_idx[p], tgt_idx[p]);
        int g = 0;
        while (g + 1 < plan.n_chunks && last >= plan.first[g + 1]) ++g;
        group[(size_t)p] = g;
        ++count[g + 1];
        if (p > 0 && g < group[(size_t)p - 1]) grouped = false;
    }
    plan.group_first[0] = 0;
    for (int g = 0; g < kUploadChunks; ++g) plan.group_first[g + 1] = plan.group_first[g] + count[g + 1];
    plan.d_order = nullptr;
    std::vector<int> order;
    if (!grouped) {
        order.resize((size_t)n_pairs);
        int cursor[kUploadChunks];
        for (int g = 0; g < kUploadChunks; ++g) cursor[g] = plan.group_first[g];
        for (int p = 0; p < n_pairs; ++p) order[(size_t)cursor[group[(size_t)p]]++] = p;
        if (c.idx_order.reserve(sizeof(int) * (size_t)n_pairs)) { cudaStreamSynchronize(c.copy_stream); cudaStreamSynchronize(c.stream); return ICPB200_ERR_CUDA; }
        // pageable source: the copy is staged before the call returns, `order` may go out of scope afterwards
        ICPB_CUDA(cudaMemcpyAsync(c.idx_order.p, order.data(), sizeof(int) * (size_t)n_pairs, cudaMemcpyHostToDevice, c.stream));
        plan.d_order = c.idx_order.as<int>();
    }
    const DevClouds s{c.pts_a.as<double>(), c.off_a.as<long long>(), cloud_off, n_clouds, max_src, max_pts, (long long)np};
    const DevClouds t{c.pts_a.as<double>(), c.off_a.as<long long>(), cloud_off, n_clouds, max_tgt, max_pts, (long long)np};
    rc = icp_enqueue(k, n_pairs, s, t, true, c.idx_a.as<int>(), c.idx_b.as<int>(), d_Ri, d_ti, c.out_r.as<double>(),
                     c.out_t.as<double>(), c.out_err.as<double>(), c.out_prev.as<double>(), c.out_iters.as<int>(),
                     c.out_status.as<int>(), c.stream, IcpTrace{}, nullptr, &plan);
    if (rc) { cudaStreamSynchronize(c.copy_stream); cudaStreamSynchronize(c.stream); return rc; }
    if (e2e_timing) {
        const auto t_enq = std::chrono::steady_clock::now();
        cudaStreamSynchronize(c.copy_stream);
        const auto t_copy = std::chrono::steady_clock::now();
        cudaStreamSynchronize(c.stream);
        const auto t_done = std::chrono::steady_clock::now();
        auto us = [](auto a_, auto b_) { return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(b_ - a_).count() / 1e3; };
        fprintf(stderr, "[icp e2e] enqueue %.0f us, copies done at %.0f us, kernels done at %.0f us (%d chunks)\n",
                us(t_begin, t_enq), us(t_begin, t_copy), us(t_begin, t_done), plan.n_chunks);
    }
    // paths that did not consume the chunk events (grid mode, big clouds) still need the data before their kernels:
    // icp_enqueue waited for every chunk in that case (see below), so nothing is pending here
    return fetch_outputs(c, n_pairs, dim, IcpOutputs{R_out, t_out, err_out, prev_err_out, iters_out, status_out});
}

int icpb200_icp_pairs_dev(int n_clouds, int dim, const double* d_pts, const int64_t* d_cloud_off,
                          int64_t max_cloud_points, int n_pairs, const int32_t* d_src_idx, const int32_t* d_tgt_idx,
                          const double* d_R_init, const double* d_t_init, double error_threshold, int max_iterations,
                          double voxel_size, int method, int normal_k, double max_corr_dist, int nn_mode,
                          double* d_R_out, double* d_t_out, double* d_err_out, double* d_prev_err_out,
                          int32_t* d_iters_out, int32_t* d_status_out, void* stream) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    const IcpCommon k{dim, error_threshold, max_iterations, voxel_size, method, normal_k, max_corr_dist, nn_mode};
    int rc = check_common(k, "icpb200_icp_pairs_dev");
    if (rc) return rc;
    if (n_pairs < 0 || n_clouds <= 0 || max_cloud_points <= 0 || !d_pts || !d_cloud_off || !d_src_idx || !d_tgt_idx ||
        !d_R_out || !d_t_out || !d_err_out || !d_iters_out || !d_status_out) {
        set_error("icpb200_icp_pairs_dev: null pointer or bad count");
        return ICPB200_ERR_ARG;
    }
    if ((rc = init_locked(-1))) return rc;
    Context& c = g_ctx;
    cudaStream_t st = stream ? (cudaStream_t)stream : c.stream;
    DevClouds s{d_pts, (const long long*)d_cloud_off, nullptr, n_clouds, max_cloud_points, max_cloud_points,
                (long long)n_clouds * max_cloud_points};
    DevClouds t = s;
    std::vector<int64_t> h_off;
    if (max_cloud_points > ICPB200_BRUTE_MAX_POINTS && n_pairs > 0) {
        // big clouds: the source / target roles need their own size bounds -> one small readback
        h_off.resize((size_t)n_clouds + 1);
        std::vector<int32_t> h_si((size_t)n_pairs), h_ti((size_t)n_pairs);
        ICPB_CUDA(cudaMemcpyAsync(h_off.data(), d_cloud_off, sizeof(int64_t) * h_off.size(), cudaMemcpyDeviceToHost, st));
        ICPB_CUDA(cudaMemcpyAsync(h_si.data(), d_src_idx, sizeof(int32_t) * h_si.size(), cudaMemcpyDeviceToHost, st));
        ICPB_CUDA(cudaMemcpyAsync(h_ti.data(), d_tgt_idx, sizeof(int32_t) * h_ti.size(), cudaMemcpyDeviceToHost, st));
        ICPB_CUDA(cudaStreamSynchronize(st));
        long long ms = 0, mt = 0;
        for (int p = 0; p < n_pairs; ++p) {
            if (h_si[p] < 0 || h_si[p] >= n_clouds || h_ti[p] < 0 || h_ti[p] >= n_clouds) {
                set_error("icpb200_icp_pairs_dev: pair %d references a cloud outside [0, %d)", p, n_clouds);
                return ICPB200_ERR_ARG;
            }
            ms = std::max<long long>(ms, h_off[h_si[p] + 1] - h_off[h_si[p]]);
            mt = std::max<long long>(mt, h_off[h_ti[p] + 1] - h_off[h_ti[p]]);
        }
        s.h_off = t.h_off = h_off.data();
        s.role_max = ms; t.role_max = mt;
        s.total_points = t.total_points = h_off[n_clouds];
    }
    const bool init = d_R_init && d_t_init;
    return icp_enqueue(k, n_pairs, s, t, true, d_src_idx, d_tgt_idx, init ? d_R_init : nullptr,
                       init ? d_t_init : nullptr, d_R_out, d_t_out, d_err_out, d_prev_err_out, d_iters_out,
                       d_status_out, st, IcpTrace{}, nullptr);
}

int icpb200_icp_trace(int dim, const double* src, int64_t n_src, const double* tgt, int64_t n_tgt,
                      const double* R_init, const double* t_init, double error_threshold, int max_iterations,
                      double voxel_size, int method, int normal_k, double max_corr_dist, int nn_mode, double* R_out,
                      double* t_out, double* err_out, double* prev_err_out, int32_t* iters_out, int32_t* status_out, double* src_ds,
                      int64_t* n_src_ds, double* tgt_ds, int64_t* n_tgt_ds, double* normals, int32_t* matches,
                      int trace_iters) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    const IcpCommon k{dim, error_threshold, max_iterations, voxel_size, method, normal_k, max_corr_dist, nn_mode};
    int rc = check_common(k, "icpb200_icp_trace");
    if (rc) return rc;
    if (!src || !tgt || n_src <= 0 || n_tgt <= 0 || !R_out || !t_out || !err_out || !iters_out || !status_out || trace_iters < 0) {
        set_error("icpb200_icp_trace: null pointer or empty cloud");
        return ICPB200_ERR_ARG;
    }
    if ((rc = init_locked(-1))) return rc;
    Context& c = g_ctx;
    const int64_t off_s[2] = {0, n_src}, off_t[2] = {0, n_tgt};
    if (c.pts_a.reserve(sizeof(double) * dim * (size_t)n_src) || c.pts_b.reserve(sizeof(double) * dim * (size_t)n_tgt) ||
        c.off_a.reserve(sizeof(int64_t) * 2) || c.off_b.reserve(sizeof(int64_t) * 2))
        return ICPB200_ERR_CUDA;
    if ((rc = reserve_outputs(c, 1, dim))) return rc;
    const size_t b_src = sizeof(double) * dim * (size_t)n_src, b_tgt = sizeof(double) * dim * (size_t)n_tgt;
    const size_t b_nrm = sizeof(double) * 2 * (size_t)n_tgt;
    const size_t b_mat = sizeof(int) * (size_t)n_src * (size_t)trace_iters;
    int* d_mat = nullptr;
    if (b_mat && matches) {
        if (c.trace.reserve(b_mat)) return ICPB200_ERR_CUDA;
        d_mat = c.trace.as<int>();
        ICPB_CUDA(cudaMemsetAsync(d_mat, 0xff, b_mat, c.stream));
    }
    ICPB_CUDA(cudaMemcpyAsync(c.pts_a.p, src, b_src, cudaMemcpyHostToDevice, c.stream));
    ICPB_CUDA(cudaMemcpyAsync(c.pts_b.p, tgt, b_tgt, cudaMemcpyHostToDevice, c.stream));
    ICPB_CUDA(cudaMemcpyAsync(c.off_a.p, off_s, sizeof(off_s), cudaMemcpyHostToDevice, c.stream));
    ICPB_CUDA(cudaMemcpyAsync(c.off_b.p, off_t, sizeof(off_t), cudaMemcpyHostToDevice, c.stream));
    const double *d_Ri, *d_ti;
    if ((rc = upload_init(c, 1, dim, R_init, t_init, &d_Ri, &d_ti))) return rc;
    const DevClouds s{c.pts_a.as<double>(), c.off_a.as<long long>(), off_s, 1, n_src, n_src, n_src};
    const DevClouds t{c.pts_b.as<double>(), c.off_b.as<long long>(), off_t, 1, n_tgt, n_tgt, n_tgt};
    IcpTrace tr;
    tr.match = d_mat; tr.iters = d_mat ? trace_iters : 0; tr.stride = (int)n_src;
    IcpArgs a;
    memset(&a, 0, sizeof(a));
    rc = icp_enqueue(k, 1, s, t, false, nullptr, nullptr, d_Ri, d_ti, c.out_r.as<double>(), c.out_t.as<double>(),
                     c.out_err.as<double>(), c.out_prev.as<double>(), c.out_iters.as<int>(), c.out_status.as<int>(),
                     c.stream, tr, &a);
    if (rc) return rc;
    // the preprocessed clouds and normals are the kernels' own intermediate buffers
    int counts[2] = {0, 0};
    ICPB_CUDA(cudaMemcpyAsync(&counts[0], a.s.ds_n, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    ICPB_CUDA(cudaMemcpyAsync(&counts[1], a.t.ds_n, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    if (src_ds) ICPB_CUDA(cudaMemcpyAsync(src_ds, a.s.ds, b_src, cudaMemcpyDeviceToHost, c.stream));
    if (tgt_ds) ICPB_CUDA(cudaMemcpyAsync(tgt_ds, a.t.ds, b_tgt, cudaMemcpyDeviceToHost, c.stream));
    if (normals && a.t.nrm) ICPB_CUDA(cudaMemcpyAsync(normals, a.t.nrm, b_nrm, cudaMemcpyDeviceToHost, c.stream));
    if (d_mat) ICPB_CUDA(cudaMemcpyAsync(matches, d_mat, b_mat, cudaMemcpyDeviceToHost, c.stream));
    rc = fetch_outputs(c, 1, dim, IcpOutputs{R_out, t_out, err_out, prev_err_out, iters_out, status_out});
    if (rc) return rc;
    if (n_src_ds) *n_src_ds = counts[0] > 0 ? counts[0] : 0;
    if (n_tgt_ds) *n_tgt_ds = counts[1] > 0 ? counts[1] : 0;
    return ICPB200_OK;
}

int icpb200_icp_last_stats(int64_t* stats8) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!stats8) { set_error("icpb200_icp_last_stats: null pointer"); return ICPB200_ERR_ARG; }
    int rc = init_locked(-1);
    if (rc) return rc;
    Context& c = g_ctx;
    for (int i = 0; i < 8; ++i) stats8[i] = 0;
    if (!c.stats.p) return ICPB200_OK;
    cudaStream_t st = c.last_icp_stream ? c.last_icp_stream : c.stream;
    ICPB_CUDA(cudaMemcpyAsync(stats8, c.stats.p, 8 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    ICPB_CUDA(cudaStreamSynchronize(st));
    // device time of the three kernels of the last call, in nanoseconds (CUDA events on its stream)
    for (int i = 0; i < 3; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c.ev[i], c.ev[i + 1]) == cudaSuccess) stats8[5 + i] = (int64_t)(ms * 1e6);
        else cudaGetLastError();
    }
    return ICPB200_OK;
}

int icpb200_icp_phase_profile(int64_t* out8) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!out8) { set_error("icpb200_icp_phase_profile: null pointer"); return ICPB200_ERR_ARG; }
    int rc = init_locked(-1);
    if (rc) return rc;
    Context& c = g_ctx;
    for (int i = 0; i < 8; ++i) out8[i] = 0;
    if (!c.stats.p) return ICPB200_OK;
    cudaStream_t st = c.last_icp_stream ? c.last_icp_stream : c.stream;
    ICPB_CUDA(cudaMemcpyAsync(out8, c.stats.as<int64_t>() + 8, 8 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    ICPB_CUDA(cudaStreamSynchronize(st));
    return ICPB200_OK;
}

int icpb200_icp_extra_stats(int64_t* out8) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!out8) { set_error("icpb200_icp_extra_stats: null pointer"); return ICPB200_ERR_ARG; }
    int rc = init_locked(-1);
    if (rc) return rc;
    Context& c = g_ctx;
    for (int i = 0; i < 8; ++i) out8[i] = 0;
    if (!c.stats.p || c.stats.cap < 24 * sizeof(unsigned long long)) return ICPB200_OK;
    cudaStream_t st = c.last_icp_stream ? c.last_icp_stream : c.stream;
    int64_t h[10];
    ICPB_CUDA(cudaMemcpyAsync(h, c.stats.as<int64_t>() + 14, sizeof(h), cudaMemcpyDeviceToHost, st));
    ICPB_CUDA(cudaStreamSynchronize(st));
    out8[0] = h[0];                                // far-field iterations
    out8[1] = h[2]; out8[2] = h[3]; out8[3] = h[4];   // grid mode: queries, candidates evaluated, cells visited
    out8[7] = h[1];                                // times a CTA joined a cluster mate's pair as a helper
    if (c.queue.p) {                               // handed-over pairs per cost class (two-phase batches)
        unsigned q[5] = {0, 0, 0, 0, 0};
        ICPB_CUDA(cudaMemcpyAsync(q, c.queue.as<unsigned>() + 8, sizeof(q), cudaMemcpyDeviceToHost, st));
        ICPB_CUDA(cudaStreamSynchronize(st));
        out8[4] = q[1]; out8[5] = q[2]; out8[6] = q[3] + q[4];      // chains, heavy, the rest
    }
    return ICPB200_OK;
}

int icpb200_icp_pair_profile(int64_t* out, int64_t cap_pairs) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    int rc = init_locked(-1);
    if (rc) return rc;
    Context& c = g_ctx;
    c.pair_prof_on = true;                         // takes effect from the next registration call
    if (!out || cap_pairs <= 0 || !c.pair_prof.p || c.pair_prof_n <= 0) return 0;
    cudaStream_t st = c.last_icp_stream ? c.last_icp_stream : c.stream;
    const int64_t n = std::min<int64_t>(cap_pairs, c.pair_prof_n);
    ICPB_CUDA(cudaMemcpyAsync(out, c.pair_prof.p, sizeof(int64_t) * 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
    ICPB_CUDA(cudaStreamSynchronize(st));
    return (int)n;
}

int icpb200_voxel_downsample(const double* pts, int64_t n, int dim, double voxel_size, double* out, int64_t* n_out) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!pts || !out || !n_out || n <= 0 || (dim != 2 && dim != 3) || !(voxel_size > 0.0)) {
        set_error("icpb200_voxel_downsample: bad argument (n=%lld, dim=%d, voxel=%g)", (long long)n, dim, voxel_size);
        return ICPB200_ERR_ARG;
    }
    int rc = init_locked(-1);
    if (rc) return rc;
    Context& c = g_ctx;
    if (n > 0x7fffffffLL / 4) {
        set_error("icpb200_voxel_downsample: %lld points exceed the 2^29 limit of this build", (long long)n);
        return ICPB200_ERR_LIMIT;
    }
    const bool big = n > 16384;                    // beyond one CTA's shared memory: global-memory radix sort
    const int sort_pad = big ? 256 : next_pow2((int)std::max<int64_t>(n, 256));
    // vox_in: points | offsets[2] ; vox_out: ds points | box[6] | count
    const size_t b_pts = sizeof(double) * dim * (size_t)n;
    if (c.vox_in.reserve(b_pts + 16) || c.vox_out.reserve(b_pts + 6 * sizeof(double) + 16)) return ICPB200_ERR_CUDA;
    const int64_t off[2] = {0, n};
    ICPB_CUDA(cudaMemcpyAsync(c.vox_in.p, pts, b_pts, cudaMemcpyHostToDevice, c.stream));
    ICPB_CUDA(cudaMemcpyAsync(c.vox_in.as<unsigned char>() + b_pts, off, sizeof(off), cudaMemcpyHostToDevice, c.stream));
    CloudSet cs;
    memset(&cs, 0, sizeof(cs));
    cs.raw = c.vox_in.as<double>();
    cs.off = reinterpret_cast<const long long*>(c.vox_in.as<unsigned char>() + b_pts);
    cs.n_clouds = 1;
    cs.ds = c.vox_out.as<double>();
    cs.box = reinterpret_cast<double*>(c.vox_out.as<unsigned char>() + b_pts);
    cs.ds_n = reinterpret_cast<int*>(c.vox_out.as<unsigned char>() + b_pts + 6 * sizeof(double));
    if (big) {
        if (c.big_keys.reserve(sizeof(unsigned long long) * 2 * (size_t)n) || c.big_idx.reserve(sizeof(unsigned) * 2 * (size_t)n))
            return ICPB200_ERR_CUDA;
        rc = launch_big_voxel(cs, dim, voxel_size, c.big_keys.as<unsigned long long>(), c.big_idx.as<unsigned>(), n, c.stream);
    } else {
        rc = launch_voxel_clouds(cs, dim, voxel_size, sort_pad, c.stream);
    }
    if (rc) return rc;
    int m = 0;
    ICPB_CUDA(cudaMemcpyAsync(&m, cs.ds_n, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    ICPB_CUDA(cudaStreamSynchronize(c.stream));
    if (m < 0) {
        set_error("icpb200_voxel_downsample: voxel index range does not fit 62 bits");
        return ICPB200_ERR_LIMIT;
    }
    ICPB_CUDA(cudaMemcpy(out, c.vox_out.p, sizeof(double) * dim * (size_t)m, cudaMemcpyDeviceToHost));
    *n_out = m;
    return ICPB200_OK;
}

// ---- rotation-search scoring (features.py:165-242, slam.py:111-183) ---------------------------------
int icpb200_rotation_scores(int n_problems, const double* src, const int64_t* src_off, const double* tgt,
                            const int64_t* tgt_off, const double* angles, const int64_t* ang_off, const double* shift,
                            double* scores_out, double* nn_dist_out, int32_t* nn_idx_out) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (n_problems <= 0 || !src || !src_off || !tgt || !tgt_off || !angles || !ang_off || !shift || !scores_out ||
        ((nn_dist_out == nullptr) != (nn_idx_out == nullptr))) {
        set_error("icpb200_rotation_scores: null pointer or n_problems <= 0");
        return ICPB200_ERR_ARG;
    }
    if (src_off[0] != 0 || tgt_off[0] != 0 || ang_off[0] != 0) { set_error("icpb200_rotation_scores: offsets must start at 0"); return ICPB200_ERR_ARG; }
    long long max_t = 0, max_a = 0;
    for (int p = 0; p < n_problems; ++p) {
        if (src_off[p + 1] < src_off[p] || tgt_off[p + 1] < tgt_off[p] || ang_off[p + 1] < ang_off[p]) {
            set_error("icpb200_rotation_scores: offsets decrease at problem %d", p);
            return ICPB200_ERR_ARG;
        }
        max_t = std::max<long long>(max_t, tgt_off[p + 1] - tgt_off[p]);
        max_a = std::max<long long>(max_a, ang_off[p + 1] - ang_off[p]);
        if (nn_idx_out && ang_off[p + 1] - ang_off[p] != 1) {
            set_error("icpb200_rotation_scores: nearest-neighbour output needs exactly one angle per problem");
            return ICPB200_ERR_ARG;
        }
    }
    const long long n_src = src_off[n_problems], n_tgt = tgt_off[n_problems], n_ang = ang_off[n_problems];
    if (n_ang == 0) return ICPB200_OK;
    if (n_problems > 65535) {
        set_error("icpb200_rotation_scores: at most 65535 problems per call (got %d)", n_problems);
        return ICPB200_ERR_LIMIT;
    }
    // a target of more than kRotSlice points (the downsampled submap of slam.py:125-126 can be) does not fit shared memory:
    // it is swept in slices, one launch each, with the running nearest neighbour of every (angle, source point) in HBM
    constexpr long long kRotSlice = 8192;
    const int n_slices = (int)std::max<long long>(1, (max_t + kRotSlice - 1) / kRotSlice);
    long long max_s = 0;
    for (int p = 0; p < n_problems; ++p) max_s = std::max<long long>(max_s, src_off[p + 1] - src_off[p]);
    int rc = init_locked(-1);
    if (rc) return rc;
    Context& c = g_ctx;
    cudaStream_t st = c.stream;
    const size_t np = (size_t)n_problems;
    const size_t b_src = sizeof(double) * 2 * (size_t)n_src, b_tgt = sizeof(double) * 2 * (size_t)n_tgt;
    const size_t b_ang = sizeof(double) * (size_t)n_ang, b_off = sizeof(int64_t) * (np + 1);
    if (c.rot_src.reserve(b_src + 16) || c.rot_tgt.reserve(b_tgt + 16) || c.rot_ang.reserve(b_ang + sizeof(double) * 2 * np) ||
        c.rot_off.reserve(3 * b_off) || c.rot_out.reserve(b_ang + (sizeof(double) + sizeof(int)) * (size_t)std::max<long long>(n_src, 1)))
        return ICPB200_ERR_CUDA;
    unsigned char* d_off = c.rot_off.as<unsigned char>();
    ICPB_CUDA(cudaMemcpyAsync(c.rot_src.p, src, b_src, cudaMemcpyHostToDevice, st));
    ICPB_CUDA(cudaMemcpyAsync(c.rot_tgt.p, tgt, b_tgt, cudaMemcpyHostToDevice, st));
    ICPB_CUDA(cudaMemcpyAsync(c.rot_ang.p, angles, b_ang, cudaMemcpyHostToDevice, st));
    ICPB_CUDA(cudaMemcpyAsync(c.rot_ang.as<unsigned char>() + b_ang, shift, sizeof(double) * 2 * np, cudaMemcpyHostToDevice, st));
    ICPB_CUDA(cudaMemcpyAsync(d_off, src_off, b_off, cudaMemcpyHostToDevice, st));
    ICPB_CUDA(cudaMemcpyAsync(d_off + b_off, tgt_off, b_off, cudaMemcpyHostToDevice, st));
    ICPB_CUDA(cudaMemcpyAsync(d_off + 2 * b_off, ang_off, b_off, cudaMemcpyHostToDevice, st));
    RotArgs a;
    a.src = c.rot_src.as<double>(); a.src_off = reinterpret_cast<const long long*>(d_off);
    a.tgt = c.rot_tgt.as<double>(); a.tgt_off = reinterpret_cast<const long long*>(d_off + b_off);
    a.angles = c.rot_ang.as<double>(); a.ang_off = reinterpret_cast<const long long*>(d_off + 2 * b_off);
    a.shift = reinterpret_cast<const double*>(c.rot_ang.as<unsigned char>() + b_ang);
    a.scores = c.rot_out.as<double>();
    a.nn_dist = nn_dist_out ? reinterpret_cast<double*>(c.rot_out.as<unsigned char>() + b_ang) : nullptr;
    a.nn_idx = nn_idx_out ? reinterpret_cast<int*>(c.rot_out.as<unsigned char>() + b_ang + sizeof(double) * (size_t)n_src) : nullptr;
    a.cap_t = round_up((int)std::max<long long>(std::min<long long>(max_t, kRotSlice), 32), 32);
    if (rot_smem_bytes(a.cap_t) > (size_t)c.max_smem_optin) { set_error("icpb200_rotation_scores: target too large for shared memory"); return ICPB200_ERR_LIMIT; }
    a.slice_len = 0; a.slice = 0; a.n_slices = 1; a.part_d2 = nullptr; a.part_j = nullptr; a.part_stride = 0;
    if (n_slices > 1) {
        const size_t rows = (size_t)n_ang, stride = (size_t)std::max<long long>(max_s, 1);
        if (rows * stride * 12 > ((size_t)4 << 30)) {
            set_error("icpb200_rotation_scores: %lld angles x %lld source points against a sliced target need more than 4 GiB of running minima", n_ang, max_s);
            return ICPB200_ERR_LIMIT;
        }
        if (c.rot_part.reserve(rows * stride * 12)) return ICPB200_ERR_CUDA;
        a.part_d2 = c.rot_part.as<double>();
        a.part_j = reinterpret_cast<int*>(c.rot_part.as<unsigned char>() + rows * stride * 8);
        a.part_stride = (long long)stride;
        a.slice_len = (int)kRotSlice; a.n_slices = n_slices;
    }
    for (int s = 0; s < n_slices; ++s) {
        a.slice = s;
        if ((rc = launch_rot_scores(a, n_problems, (int)max_a, c.sm_count, st))) return rc;
    }
    ICPB_CUDA(cudaMemcpyAsync(scores_out, a.scores, b_ang, cudaMemcpyDeviceToHost, st));
    if (nn_dist_out) {
        ICPB_CUDA(cudaMemcpyAsync(nn_dist_out, a.nn_dist, sizeof(double) * (size_t)n_src, cudaMemcpyDeviceToHost, st));
        ICPB_CUDA(cudaMemcpyAsync(nn_idx_out, a.nn_idx, sizeof(int) * (size_t)n_src, cudaMemcpyDeviceToHost, st));
    }
    ICPB_CUDA(cudaStreamSynchronize(st));
    return ICPB200_OK;
}

// ---- occupancy grid --------------------------------------------------------------

void OccGrid::release_all() {
    DevBuf* bufs[] = {&grid, &origins, &hits, &hit_off, &local_pts, &poses, &in_pack, &origin_cell, &ray_cell, &ray_scan,
                      &counts, &offsets, &sums, &runs, &order, &small, &tile_prof,
                      &slotmap, &slot_cell, &ord, &tile_count, &hit_off_shift, &items, &multi, &ncount, &ev, &ev_count, &class_off, &tile_flag,
                      &dirty, &pack, &pack_ids};
    for (DevBuf* b : bufs) b->release();
    if (aux_stream) { cudaStreamDestroy(aux_stream); aux_stream = nullptr; }
    if (ev_fork) { cudaEventDestroy(ev_fork); ev_fork = nullptr; }
    if (ev_join) { cudaEventDestroy(ev_join); ev_join = nullptr; }
    if (ev_hit) { cudaEventDestroy(ev_hit); ev_hit = nullptr; }
    if (ev_stats) { cudaEventDestroy(ev_stats); ev_stats = nullptr; }
    if (ev_done) { cudaEventDestroy(ev_done); ev_done = nullptr; }
    if (pending_host) { cudaFreeHost(pending_host); pending_host = nullptr; }
    if (h_in_pack) { cudaFreeHost(h_in_pack); h_in_pack = nullptr; h_in_cap = 0; }
    if (h_pack) { cudaFreeHost(h_pack); h_pack = nullptr; h_pack_cap = 0; }
    if (peers_attached) {
        for (int p = 0; p < world && p < 16; ++p)
            if (p != rank) { if (peer_grid[p]) cudaIpcCloseMemHandle(peer_grid[p]); if (peer_dirty[p]) cudaIpcCloseMemHandle(peer_dirty[p]); }
        peers_attached = false;
    }
    stats_pending = false;
}

void* icpb200_grid_create(int nx, int ny, double min_x, double min_y, double resolution, double l_hit,
                          double l_miss, double lo_min, double lo_max) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (nx <= 0 || ny <= 0 || !(resolution > 0.0) || !(lo_min <= lo_max)) {
        set_error("icpb200_grid_create: bad argument (nx=%d ny=%d resolution=%g clamp=[%g,%g])", nx, ny, resolution, lo_min, lo_max);
        return nullptr;
    }
    if ((long long)nx * ny > (1LL << 31)) { set_error("icpb200_grid_create: grid larger than 2^31 cells"); return nullptr; }
    if (init_locked(-1)) return nullptr;
    OccGrid* g = new OccGrid();
    g->nx = nx; g->ny = ny;
    g->tiles_x = (nx + kOccTile - 1) / kOccTile;
    g->tiles_y = (ny + kOccTile - 1) / kOccTile;
    g->min_x = min_x; g->min_y = min_y; g->res = resolution;
    g->l_hit = l_hit; g->l_miss = l_miss; g->lo_min = lo_min; g->lo_max = lo_max;
    g->zero_outside_clamp = ((float)lo_min > 0.f) || ((float)lo_max < 0.f);
    g->apply_ctas = occ_apply_ctas(g_ctx.sm_count);
    g->fast_ctas = occ_fast_ctas(g_ctx.sm_count);
    if (const char* e = getenv("ICPB200_OCC_PATH")) g->use_fast = strcmp(e, "ordered") != 0;
    if (g->grid.reserve(sizeof(float) * (size_t)nx * ny) ||
        cudaMemsetAsync(g->grid.p, 0, sizeof(float) * (size_t)nx * ny, g_ctx.stream) != cudaSuccess ||
        cudaStreamSynchronize(g_ctx.stream) != cudaSuccess) {
        if (!g_error[0]) set_error("icpb200_grid_create: device allocation failed");
        g->release_all();
        delete g;
        return nullptr;
    }
    return g;
}

void icpb200_grid_destroy(void* grid) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!grid) return;
    OccGrid* g = static_cast<OccGrid*>(grid);
    if (g_ctx.ready) { occ_collect(*g); cudaDeviceSynchronize(); }    // updates may have run on a caller's stream
    g->release_all();
    delete g;
}

int icpb200_grid_set_shard(void* grid, int rank, int world) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!grid || world < 1 || rank < 0 || rank >= world) { set_error("icpb200_grid_set_shard: bad rank/world"); return ICPB200_ERR_ARG; }
    OccGrid* g = static_cast<OccGrid*>(grid);
    g->rank = rank; g->world = world;
    return ICPB200_OK;
}

int icpb200_grid_update(void* grid, int n_scans, const double* origins, const double* hits, const int64_t* hit_off) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!grid || n_scans < 0 || (n_scans > 0 && (!origins || !hit_off))) { set_error("icpb200_grid_update: null pointer"); return ICPB200_ERR_ARG; }
    if (n_scans == 0) return ICPB200_OK;
    if (hit_off[0] != 0) { set_error("icpb200_grid_update: hit_off[0] must be 0"); return ICPB200_ERR_ARG; }
    for (int s = 0; s < n_scans; ++s)
        if (hit_off[s + 1] < hit_off[s]) { set_error("icpb200_grid_update: hit_off decreases at scan %d", s); return ICPB200_ERR_ARG; }
    const long long n_rays = hit_off[n_scans];
    if (n_rays > 0 && !hits) { set_error("icpb200_grid_update: hits is null"); return ICPB200_ERR_ARG; }
    int rc = init_locked(-1);
    if (rc) return rc;
    OccGrid* g = static_cast<OccGrid*>(grid);
    cudaStream_t st = g_ctx.stream;
    if ((rc = occ_collect(*g))) return rc;
    if (n_rays == 0) { g->stats[0] = g->stats[1] = g->stats[2] = g->stats[3] = 0; return ICPB200_OK; }
    const size_t b_org = sizeof(double) * 2 * (size_t)n_scans, b_off = sizeof(int64_t) * ((size_t)n_scans + 1),
                 b_hits = sizeof(double) * 2 * (size_t)n_rays;
    if (b_org + b_off + b_hits <= (256u << 10)) {
        // update_scan and other small updates (the online loop, slam.py:408, 557): three copies from pageable memory
        // would each go through the driver's bounce buffer and block; one block, one copy instead
        const size_t at_off = (b_org + 15) & ~(size_t)15, at_hits = (at_off + b_off + 15) & ~(size_t)15, total = at_hits + b_hits;
        if (total > g->h_in_cap) {
            if (g->h_in_pack) { cudaFreeHost(g->h_in_pack); g->h_in_pack = nullptr; g->h_in_cap = 0; }
            ICPB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g->h_in_pack), 256u << 10, cudaHostAllocDefault));
            g->h_in_cap = 256u << 10;
        }
        if (g->in_pack.reserve(256u << 10)) return ICPB200_ERR_CUDA;
        memcpy(g->h_in_pack, origins, b_org);
        memcpy(g->h_in_pack + at_off, hit_off, b_off);
        memcpy(g->h_in_pack + at_hits, hits, b_hits);
        ICPB_CUDA(cudaMemcpyAsync(g->in_pack.p, g->h_in_pack, total, cudaMemcpyHostToDevice, st));
        unsigned char* d = g->in_pack.as<unsigned char>();
        // the host block is reused by the next call: occ_update_device ends with a wait on `st`, the copy is done by then
        rc = occ_update_device(*g, n_scans, reinterpret_cast<const double*>(d), reinterpret_cast<const double*>(d + at_hits),
                               reinterpret_cast<const long long*>(d + at_off), reinterpret_cast<const long long*>(hit_off), st);
        if (rc) cudaStreamSynchronize(st);         // an early error return may leave the copy in flight
        return rc;
    }
    if (g->origins.reserve(b_org) || g->hits.reserve(b_hits) || g->hit_off.reserve(b_off)) return ICPB200_ERR_CUDA;
    ICPB_CUDA(cudaMemcpyAsync(g->origins.p, origins, b_org, cudaMemcpyHostToDevice, st));
    ICPB_CUDA(cudaMemcpyAsync(g->hits.p, hits, b_hits, cudaMemcpyHostToDevice, st));
    ICPB_CUDA(cudaMemcpyAsync(g->hit_off.p, hit_off, b_off, cudaMemcpyHostToDevice, st));
    return occ_update_device(*g, n_scans, g->origins.as<double>(), g->hits.as<double>(), g->hit_off.as<long long>(),
                             reinterpret_cast<const long long*>(hit_off), st);
}

int icpb200_grid_rebuild(void* grid, int n_scans, const double* poses, const double* local_pts, const int64_t* off) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!grid || n_scans < 0 || (n_scans > 0 && (!poses || !off))) { set_error("icpb200_grid_rebuild: null pointer"); return ICPB200_ERR_ARG; }
    if (n_scans > 0) {
        if (off[0] != 0) { set_error("icpb200_grid_rebuild: off[0] must be 0"); return ICPB200_ERR_ARG; }
        for (int s = 0; s < n_scans; ++s)
            if (off[s + 1] < off[s]) { set_error("icpb200_grid_rebuild: off decreases at scan %d", s); return ICPB200_ERR_ARG; }
        if (off[n_scans] > 0 && !local_pts) { set_error("icpb200_grid_rebuild: local_pts is null"); return ICPB200_ERR_ARG; }
    }
    int rc = init_locked(-1);
    if (rc) return rc;
    OccGrid* g = static_cast<OccGrid*>(grid);
    cudaStream_t st = g_ctx.stream;
    if ((rc = occ_collect(*g))) return rc;
    ICPB_CUDA(cudaMemsetAsync(g->grid.p, 0, sizeof(float) * (size_t)g->nx * g->ny, st));      // slam.py:273, mapping.py:143-145
    if (g->dirty.p) ICPB_CUDA(cudaMemsetAsync(g->dirty.p, 0, g->dirty.cap, st));
    g->all_dirty = false;
    g->seen_nonempty_scan = false;
    g->virgin_finalised = false;
    g->stats[0] = g->stats[1] = g->stats[2] = g->stats[3] = 0;
    const long long n_pts = n_scans > 0 ? (long long)off[n_scans] : 0;
    if (n_pts == 0) { ICPB_CUDA(cudaStreamSynchronize(st)); return ICPB200_OK; }
    if (g->origins.reserve(sizeof(double) * 2 * (size_t)n_scans) || g->hits.reserve(sizeof(double) * 2 * (size_t)n_pts) ||
        g->hit_off.reserve(sizeof(int64_t) * ((size_t)n_scans + 1)) || g->local_pts.reserve(sizeof(double) * 2 * (size_t)n_pts) ||
        g->poses.reserve(sizeof(double) * 9 * (size_t)n_scans))
        return ICPB200_ERR_CUDA;
    ICPB_CUDA(cudaMemcpyAsync(g->poses.p, poses, sizeof(double) * 9 * (size_t)n_scans, cudaMemcpyHostToDevice, st));
    ICPB_CUDA(cudaMemcpyAsync(g->hit_off.p, off, sizeof(int64_t) * ((size_t)n_scans + 1), cudaMemcpyHostToDevice, st));
    ICPB_CUDA(cudaMemcpyAsync(g->local_pts.p, local_pts, sizeof(double) * 2 * (size_t)n_pts, cudaMemcpyHostToDevice, st));
    if ((rc = occ_transform_history(n_scans, n_pts, g->poses.as<double>(), g->local_pts.as<double>(), g->hit_off.as<long long>(),
                                    g->hits.as<double>(), g->origins.as<double>(), st)))
        return rc;
    return occ_update_device(*g, n_scans, g->origins.as<double>(), g->hits.as<double>(), g->hit_off.as<long long>(),
                             reinterpret_cast<const long long*>(off), st);
}

int icpb200_grid_update_dev(void* grid, int n_scans, const double* d_origins, const double* d_hits,
                            const int64_t* d_hit_off, int64_t total_hits, void* stream) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!grid || n_scans < 0 || total_hits < 0 || (n_scans > 0 && (!d_origins || !d_hit_off || (total_hits > 0 && !d_hits)))) {
        set_error("icpb200_grid_update_dev: null pointer or negative count");
        return ICPB200_ERR_ARG;
    }
    if (n_scans == 0 || total_hits == 0) return ICPB200_OK;
    int rc = init_locked(-1);
    if (rc) return rc;
    OccGrid* g = static_cast<OccGrid*>(grid);
    cudaStream_t st = stream ? (cudaStream_t)stream : g_ctx.stream;
    if ((rc = occ_collect(*g))) return rc;
    // the grid's work buffers are shared by its updates: an update on another stream waits for the previous one
    if (!g->ev_done) ICPB_CUDA(cudaEventCreateWithFlags(&g->ev_done, cudaEventDisableTiming));
    else ICPB_CUDA(cudaStreamWaitEvent(st, g->ev_done, 0));
    struct Mark { OccGrid* g; cudaStream_t st; ~Mark() { cudaEventRecord(g->ev_done, st); } } mark{g, st};
    if (g->use_fast && !g->zero_outside_clamp && n_scans <= kOccMaxChunkScans) {
        // One chunk of the order-free path: the offsets are checked on the device, the host waits once (for the
        // binning totals, underneath the fill pass) and returns with the rest of the update enqueued.  The hit-overflow
        // flag and the statistics are collected by the next call on this grid.
        rc = occ_update_fast(*g, n_scans, d_origins, d_hits, reinterpret_cast<const long long*>(d_hit_off), nullptr,
                             (long long)total_hits, true, st);
        if (rc < 0) { g->slotmap.release(); g->ord.release(); g->ncount.release(); }
        return rc;
    }
    std::vector<long long> h_off((size_t)n_scans + 1);
    ICPB_CUDA(cudaMemcpyAsync(h_off.data(), d_hit_off, sizeof(long long) * h_off.size(), cudaMemcpyDeviceToHost, st));
    ICPB_CUDA(cudaStreamSynchronize(st));
    bool monotone = true;
    for (int s = 0; s < n_scans && monotone; ++s) monotone = h_off[s + 1] >= h_off[s];
    if (h_off[0] != 0 || h_off[n_scans] != total_hits || !monotone) {
        set_error("icpb200_grid_update_dev: hit_off[0] must be 0, hit_off must not decrease and hit_off[n_scans] must equal total_hits");
        return ICPB200_ERR_ARG;
    }
    return occ_update_device(*g, n_scans, d_origins, d_hits, reinterpret_cast<const long long*>(d_hit_off), h_off.data(), st);
}

int icpb200_grid_read(void* grid, float* out) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!grid || !out) { set_error("icpb200_grid_read: null pointer"); return ICPB200_ERR_ARG; }
    int rc = init_locked(-1);
    if (rc) return rc;
    OccGrid* g = static_cast<OccGrid*>(grid);
    if ((rc = occ_collect(*g))) return rc;
    ICPB_CUDA(cudaMemcpyAsync(out, g->grid.p, sizeof(float) * (size_t)g->nx * g->ny, cudaMemcpyDeviceToHost, g_ctx.stream));
    ICPB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return ICPB200_OK;
}

int icpb200_grid_reset(void* grid) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!grid) { set_error("icpb200_grid_reset: null pointer"); return ICPB200_ERR_ARG; }
    int rc = init_locked(-1);
    if (rc) return rc;
    OccGrid* g = static_cast<OccGrid*>(grid);
    if ((rc = occ_collect(*g))) return rc;
    ICPB_CUDA(cudaMemsetAsync(g->grid.p, 0, sizeof(float) * (size_t)g->nx * g->ny, g_ctx.stream));   // mapping.py:143-145
    if (g->dirty.p) ICPB_CUDA(cudaMemsetAsync(g->dirty.p, 0, g->dirty.cap, g_ctx.stream));
    g->all_dirty = false;
    ICPB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    g->seen_nonempty_scan = false;
    g->virgin_finalised = false;
    return ICPB200_OK;
}

int icpb200_grid_read_view(void* grid, int view, int dirty_only, float* out, int32_t* tiles_copied) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!grid || !out || view < 0 || view > 2) { set_error("icpb200_grid_read_view: null pointer or unknown view %d", view); return ICPB200_ERR_ARG; }
    int rc = init_locked(-1);
    if (rc) return rc;
    OccGrid* g = static_cast<OccGrid*>(grid);
    if ((rc = occ_collect(*g))) return rc;
    int n = 0;
    rc = occ_read_view(*g, out, view, dirty_only != 0, &n, g_ctx.stream);
    if (tiles_copied) *tiles_copied = n;
    return rc;
}

int icpb200_grid_ipc_export(void* grid, unsigned char* handles128) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!grid || !handles128) { set_error("icpb200_grid_ipc_export: null pointer"); return ICPB200_ERR_ARG; }
    int rc = init_locked(-1);
    if (rc) return rc;
    OccGrid* g = static_cast<OccGrid*>(grid);
    if ((rc = occ_ensure_dirty(*g, g_ctx.stream))) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h0, h1;
    ICPB_CUDA(cudaIpcGetMemHandle(&h0, g->grid.p));
    ICPB_CUDA(cudaIpcGetMemHandle(&h1, g->dirty.p));
    memcpy(handles128, &h0, 64);
    memcpy(handles128 + 64, &h1, 64);
    return ICPB200_OK;
}

int icpb200_grid_ipc_attach(void* grid, int world, int rank, const unsigned char* handles) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!grid || !handles || world < 1 || world > 16 || rank < 0 || rank >= world) {
        set_error("icpb200_grid_ipc_attach: null pointer or bad rank / world (at most 16 ranks)");
        return ICPB200_ERR_ARG;
    }
    int rc = init_locked(-1);
    if (rc) return rc;
    OccGrid* g = static_cast<OccGrid*>(grid);
    if (g->world != world || g->rank != rank) { set_error("icpb200_grid_ipc_attach: call icpb200_grid_set_shard(rank, world) first"); return ICPB200_ERR_ARG; }
    for (int p = 0; p < world; ++p) {
        if (p == rank) { g->peer_grid[p] = g->grid.p; g->peer_dirty[p] = g->dirty.p; continue; }
        cudaIpcMemHandle_t h0, h1;
        memcpy(&h0, handles + (size_t)p * 128, 64);
        memcpy(&h1, handles + (size_t)p * 128 + 64, 64);
        ICPB_CUDA(cudaIpcOpenMemHandle(&g->peer_grid[p], h0, cudaIpcMemLazyEnablePeerAccess));
        ICPB_CUDA(cudaIpcOpenMemHandle(&g->peer_dirty[p], h1, cudaIpcMemLazyEnablePeerAccess));
    }
    g->peers_attached = true;
    return ICPB200_OK;
}

int icpb200_grid_push_tiles(void* grid, void* stream) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!grid) { set_error("icpb200_grid_push_tiles: null handle"); return ICPB200_ERR_ARG; }
    int rc = init_locked(-1);
    if (rc) return rc;
    OccGrid* g = static_cast<OccGrid*>(grid);
    return occ_push_to_peers(*g, stream ? (cudaStream_t)stream : g_ctx.stream);
}

void* icpb200_grid_device_ptr(void* grid) {
    return grid ? static_cast<OccGrid*>(grid)->grid.p : nullptr;
}

int icpb200_grid_tile_profile(void* grid, int64_t* out, int64_t cap_tiles) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!grid) { set_error("icpb200_grid_tile_profile: null pointer"); return ICPB200_ERR_ARG; }
    OccGrid* g = static_cast<OccGrid*>(grid);
    g->profile_tiles = true;                       // takes effect from the next update
    if (!out || cap_tiles <= 0 || !g->tile_prof.p) return 0;
    int rc = init_locked(-1);
    if (rc) return rc;
    const int64_t n = std::min<int64_t>(cap_tiles, (int64_t)g->tiles_x * g->tiles_y);
    ICPB_CUDA(cudaMemcpy(out, g->tile_prof.p, sizeof(int64_t) * 8 * (size_t)n, cudaMemcpyDeviceToHost));
    return (int)n;
}

int icpb200_pin_host(void* ptr, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!ptr || bytes == 0) { set_error("icpb200_pin_host: null pointer or empty range"); return ICPB200_ERR_ARG; }
    int rc = init_locked(-1);
    if (rc) return rc;
    ICPB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));      // mapped: read-outs may write it in place
    return ICPB200_OK;
}

int icpb200_unpin_host(void* ptr) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!ptr) { set_error("icpb200_unpin_host: null pointer"); return ICPB200_ERR_ARG; }
    if (!g_ctx.ready) return ICPB200_OK;
    ICPB_CUDA(cudaHostUnregister(ptr));
    return ICPB200_OK;
}

// ---- device-resident submap ----------------------------------------------------------------------------------------
void* icpb200_submap_create(int dim, int capacity_scans) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if ((dim != 2 && dim != 3) || capacity_scans <= 0) { set_error("icpb200_submap_create: dim must be 2 or 3, capacity > 0"); return nullptr; }
    if (init_locked(-1)) return nullptr;
    Submap* sm = new Submap();
    sm->dim = dim; sm->capacity = capacity_scans;
    for (int i = capacity_scans - 1; i >= 0; --i) sm->free_slots.push_back(i);
    return sm;
}

void icpb200_submap_destroy(void* submap) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!submap) return;
    Submap* sm = static_cast<Submap*>(submap);
    if (g_ctx.ready) cudaStreamSynchronize(g_ctx.stream);
    sm->release();
    delete sm;
}

int icpb200_submap_clear(void* submap) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!submap) { set_error("icpb200_submap_clear: null handle"); return ICPB200_ERR_ARG; }
    Submap* sm = static_cast<Submap*>(submap);
    for (const SubmapSlot& s : sm->scans) sm->free_slots.push_back(s.slot);
    sm->scans.clear();
    ++sm->version;
    return ICPB200_OK;
}

int icpb200_submap_push(void* submap, const double* pts, int64_t n) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!submap || n < 0 || (n > 0 && !pts)) { set_error("icpb200_submap_push: null pointer or negative count"); return ICPB200_ERR_ARG; }
    int rc = init_locked(-1);
    if (rc) return rc;
    Submap* sm = static_cast<Submap*>(submap);
    Context& c = g_ctx;
    if ((int)sm->scans.size() == sm->capacity) {                     // slam.py:561-562: pop(0)
        sm->free_slots.push_back(sm->scans.front().slot);
        sm->scans.erase(sm->scans.begin());
    }
    if (n > sm->slot_points || !sm->arena.p) {                        // (re)size the arena: slots of a power-of-two row count
        long long want = sm->slot_points;
        while (want < n) want *= 2;
        DevBuf fresh;
        if (fresh.reserve(sizeof(double) * sm->dim * (size_t)want * (size_t)sm->capacity)) return ICPB200_ERR_CUDA;
        for (const SubmapSlot& s : sm->scans)
            ICPB_CUDA(cudaMemcpyAsync(fresh.as<double>() + (size_t)s.slot * want * sm->dim,
                                      sm->arena.as<double>() + (size_t)s.slot * sm->slot_points * sm->dim,
                                      sizeof(double) * sm->dim * (size_t)s.n, cudaMemcpyDeviceToDevice, c.stream));
        ICPB_CUDA(cudaStreamSynchronize(c.stream));
        sm->arena.release();
        sm->arena = fresh;
        sm->slot_points = want;
    }
    const int slot = sm->free_slots.back();
    sm->free_slots.pop_back();
    if (n > 0) {
        ICPB_CUDA(cudaMemcpyAsync(sm->arena.as<double>() + (size_t)slot * sm->slot_points * sm->dim, pts,
                                  sizeof(double) * sm->dim * (size_t)n, cudaMemcpyHostToDevice, c.stream));
        ICPB_CUDA(cudaStreamSynchronize(c.stream));                   // the caller's buffer is free on return
    }
    sm->scans.push_back(SubmapSlot{slot, (long long)n});
    ++sm->version;
    return ICPB200_OK;
}

int icpb200_submap_size(void* submap, int64_t* n_scans, int64_t* n_points) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!submap) { set_error("icpb200_submap_size: null handle"); return ICPB200_ERR_ARG; }
    Submap* sm = static_cast<Submap*>(submap);
    long long total = 0;
    for (const SubmapSlot& s : sm->scans) total += s.n;
    if (n_scans) *n_scans = (int64_t)sm->scans.size();
    if (n_points) *n_points = total;
    return ICPB200_OK;
}

int icpb200_submap_build(void* submap, double submap_voxel, double* out, int64_t out_capacity_rows, int64_t* n_out) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!submap || !(submap_voxel > 0.0) || !n_out) { set_error("icpb200_submap_build: null pointer or voxel <= 0"); return ICPB200_ERR_ARG; }
    int rc = init_locked(-1);
    if (rc) return rc;
    Submap* sm = static_cast<Submap*>(submap);
    long long m = 0;
    if ((rc = submap_build(g_ctx, *sm, submap_voxel, &m, g_ctx.stream))) return rc;
    *n_out = m;
    if (out && m > 0) {
        if (out_capacity_rows < m) { set_error("icpb200_submap_build: output holds %lld rows, %lld needed", (long long)out_capacity_rows, m); return ICPB200_ERR_ARG; }
        ICPB_CUDA(cudaMemcpyAsync(out, sm->built.p, sizeof(double) * sm->dim * (size_t)m, cudaMemcpyDeviceToHost, g_ctx.stream));
        ICPB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    }
    return ICPB200_OK;
}

int icpb200_submap_icp(void* submap, double submap_voxel, int n_sources, const double* src, const int64_t* src_off,
                       const double* R_init, const double* t_init, double error_threshold, int max_iterations,
                       double voxel_size, int method, int normal_k, double max_corr_dist, int nn_mode, double* R_out,
                       double* t_out, double* err_out, double* prev_err_out, int32_t* iters_out, int32_t* status_out) {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (!submap || !(submap_voxel > 0.0)) { set_error("icpb200_submap_icp: null handle or submap_voxel <= 0"); return ICPB200_ERR_ARG; }
    Submap* sm = static_cast<Submap*>(submap);
    const IcpCommon k{sm->dim, error_threshold, max_iterations, voxel_size, method, normal_k, max_corr_dist, nn_mode};
    int rc = check_common(k, "icpb200_submap_icp");
    if (rc) return rc;
    if (n_sources < 0 || (n_sources > 0 && (!src || !src_off || !R_out || !t_out || !err_out || !iters_out || !status_out))) {
        set_error("icpb200_submap_icp: null pointer or negative count");
        return ICPB200_ERR_ARG;
    }
    if (n_sources == 0) return ICPB200_OK;
    long long max_s;
    if ((rc = check_offsets(src_off, n_sources, "src_off", &max_s))) return rc;
    if ((rc = init_locked(-1))) return rc;
    Context& c = g_ctx;
    long long m = 0;
    if ((rc = submap_build(c, *sm, submap_voxel, &m, c.stream))) return rc;
    if (m <= 0) { set_error("icpb200_submap_icp: the submap is empty"); return ICPB200_ERR_ARG; }
    const int dim = sm->dim;
    const size_t ns = (size_t)src_off[n_sources];
    if (c.pts_a.reserve(sizeof(double) * dim * ns) || c.off_a.reserve(sizeof(int64_t) * ((size_t)n_sources + 1)) ||
        c.idx_b.reserve(sizeof(int32_t) * (size_t)n_sources))
        return ICPB200_ERR_CUDA;
    if ((rc = reserve_outputs(c, n_sources, dim))) return rc;
    ICPB_CUDA(cudaMemcpyAsync(c.pts_a.p, src, sizeof(double) * dim * ns, cudaMemcpyHostToDevice, c.stream));
    ICPB_CUDA(cudaMemcpyAsync(c.off_a.p, src_off, sizeof(int64_t) * ((size_t)n_sources + 1), cudaMemcpyHostToDevice, c.stream));
    ICPB_CUDA(cudaMemsetAsync(c.idx_b.p, 0, sizeof(int32_t) * (size_t)n_sources, c.stream));       // every pair: target cloud 0
    const double *d_Ri, *d_ti;
    if ((rc = upload_init(c, n_sources, dim, R_init, t_init, &d_Ri, &d_ti))) return rc;
    const int64_t h_toff[2] = {0, m};
    const DevClouds s{c.pts_a.as<double>(), c.off_a.as<long long>(), src_off, n_sources, max_s, max_s, (long long)ns};
    const DevClouds t{sm->built.as<double>(), sm->built_off.as<long long>(), h_toff, 1, m, m, m};
    if (sm->prep_version != sm->version) { sm->prep.ready = false; sm->prep_version = sm->version; }
    rc = icp_enqueue(k, n_sources, s, t, false, nullptr, c.idx_b.as<int>(), d_Ri, d_ti, c.out_r.as<double>(), c.out_t.as<double>(),
                     c.out_err.as<double>(), c.out_prev.as<double>(), c.out_iters.as<int>(), c.out_status.as<int>(),
                     c.stream, IcpTrace{}, nullptr, nullptr, &sm->prep);
    if (rc) { cudaStreamSynchronize(c.stream); return rc; }
    return fetch_outputs(c, n_sources, dim, IcpOutputs{R_out, t_out, err_out, prev_err_out, iters_out, status_out});
}

int icpb200_grid_last_stats(void* grid, int64_t* stats4) {
    if (!grid || !stats4) { set_error("icpb200_grid_last_stats: null pointer"); return ICPB200_ERR_ARG; }
    OccGrid* g = static_cast<OccGrid*>(grid);
    {
        std::lock_guard<std::mutex> lk(g_api_mutex);
        const int rc = occ_collect(*g);
        if (rc) return rc;
    }
    for (int i = 0; i < 4; ++i) stats4[i] = g->stats[i];
    return ICPB200_OK;
}

}  // extern "C"
