// Shared host-side plumbing for libicp_b200.so: error reporting, the
// per-process device context and growable device buffers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>

namespace icpb {

void set_error(const char* fmt, ...);
extern std::mutex g_api_mutex;
extern long long g_launches;          // kernels launched by this library (bench.py: gpu_launches)

#define ICPB_CUDA(call)                                                              \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            icpb::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,     \
                            cudaGetErrorString(e_));                                 \
            return ICPB200_ERR_CUDA;                                                 \
        }                                                                            \
    } while (0)

#define ICPB_LAUNCH_CHECK()                                                          \
    do {                                                                             \
        ++icpb::g_launches;                                                          \
        ICPB_CUDA(cudaGetLastError());                                               \
    } while (0)

// A device allocation that only ever grows; owned by the context.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);        // returns 0 / ICPB200_ERR_CUDA
    void release();
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

constexpr int kUploadChunks = 4;

struct Context {
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;     // library-owned stream for the host-buffer entry points
    cudaStream_t copy_stream = nullptr;                        // chunked cloud uploads of icpb200_icp_pairs
    cudaEvent_t chunk_ev[kUploadChunks] = {}, chunk_done[kUploadChunks] = {}, group_done[kUploadChunks] = {}, fork_ev = nullptr;
    cudaStream_t chunk_stream[kUploadChunks] = {};              // K1/K2 of upload chunk k
    // ICP staging (host-buffer entry points)
    DevBuf pts_a, pts_b, off_a, off_b, idx_a, idx_b, idx_order, rinit, tinit;
    DevBuf out_r, out_t, out_err, out_prev, out_iters, out_status, queue, trace, stats;
    // preprocessed form of cloud sets A and B (see CloudSet in icp_kernel.h)
    DevBuf aux_ds[2], aux_n[2], aux_box[2], aux_nrm[2], aux_flags[2];
    // big-cloud path: radix-sort scratch and hash grids
    DevBuf big_keys, big_idx, grid_start, grid_items, grid_cell, grid_desc, grid_off, grid_buckets;
    DevBuf cont_cur, cont_match, cont_d2lb, cont_moved, cont_scalar, cont_list;
    DevBuf pair_prof;                  // icpb200_icp_pair_profile: [pair][8] counters of the last registration call
    bool pair_prof_on = false;
    int pair_prof_n = 0;
    void* h_stage = nullptr;           // page-locked staging for the result read-back (one wait, then host copies)
    size_t h_stage_cap = 0;
    cudaStream_t last_icp_stream = nullptr;
    cudaEvent_t icp_done = nullptr;    // end of the last registration enqueue: the next one (any stream) waits for it
    bool icp_done_valid = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // K1 start | K2 start | K3 start | K3 end
    // voxel_downsample entry point
    DevBuf vox_in, vox_out;
    // rotation-search scoring
    DevBuf rot_src, rot_tgt, rot_ang, rot_off, rot_out, rot_part;
};

Context& ctx();

}  // namespace icpb
