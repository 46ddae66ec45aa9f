// Host <-> kernel argument block and launchers for the per-pair ICP kernel.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace icpb {

struct IcpArgs {
    int n_pairs;
    // clouds: pair p uses cloud (src_idx ? src_idx[p] : p) of the source set,
    // rows src_off[c] .. src_off[c+1]) of `src`; likewise for the target set
    const double* src;
    const long long* src_off;
    const int* src_idx;
    const double* tgt;
    const long long* tgt_off;
    const int* tgt_idx;
    const double* R_init;      // n_pairs * dim * dim or nullptr
    const double* t_init;      // n_pairs * dim or nullptr
    double err_thr;
    int max_iter;
    double voxel;
    int method;
    int normal_k;
    double max_corr;           // < 0: no gate
    double* R_out;
    double* t_out;
    double* err_out;
    double* prev_out;          // may be nullptr
    int* iters_out;
    int* status_out;
    // per-CTA global workspace: src_ds | tgt_ds | normals
    double* ws;
    size_t ws_stride;          // doubles per CTA
    int cap_s, cap_t;          // multiples of 32, >= largest raw cloud
    int sort_pad;              // power of two >= largest raw cloud, >= 256
    unsigned int* queue;       // zeroed before launch
    // optional trace outputs (single-pair debug entry point)
    double* trace_src;
    double* trace_tgt;
    double* trace_nrm;
    int* trace_match;
    int trace_iters;
    int* trace_counts;
};

size_t icp_smem_bytes(int dim, int cap_s, int cap_t, int sort_pad);
inline size_t icp_ws_doubles(int dim, int cap_s, int cap_t) {
    return (size_t)dim * cap_s + (size_t)dim * cap_t + 2 * (size_t)cap_t;
}
int icp_max_ctas_per_sm(int dim, size_t smem);
int launch_icp_pairs(const IcpArgs& a, int dim, int n_ctas, size_t smem, cudaStream_t stream);
int launch_voxel(const double* d_pts, int n, int dim, double voxel, double* d_out, int* d_n_out,
                 int sort_pad, cudaStream_t stream);

}  // namespace icpb
