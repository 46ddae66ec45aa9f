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

struct Context {
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;     // library-owned stream for the host-buffer entry points
    // ICP staging / workspace
    DevBuf pts_a, pts_b, off_a, off_b, idx_a, idx_b, rinit, tinit;
    DevBuf out_r, out_t, out_err, out_prev, out_iters, out_status, icp_ws, queue;
    DevBuf trace;
    // voxel_downsample entry point
    DevBuf vox_in, vox_out;
};

Context& ctx();
int ensure_ready();                    // lazily binds to the current / default device

}  // namespace icpb
