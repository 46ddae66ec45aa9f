// Host side of libicp_b200.so: 2-D pose-graph Gauss-Newton.
//
// Replaces PoseGraph2D.optimize of the reference (utilities/pose_graph.py:83-134), which assembles a dense 3n x 3n
// normal matrix and calls np.linalg.solve on it in every iteration: O(n^3), ~0.5 s per iteration at 2000 nodes and the
// dominant cost of a loop closure once registration and mapping run on the GPU.  The matrix of a SLAM pose graph is a
// block-tridiagonal chain (odometry) plus one off-diagonal block pair per loop closure, so this solver keeps it as a
// block skyline -- for every block row the blocks from its first nonzero column to the diagonal -- and factors it with
// a block Cholesky (the matrix is symmetric positive definite: sum of J^T Omega J plus the anchor's 1e10 I).  Fill stays
// inside the skyline; a chain row costs one block, a loop-closure row as many blocks as the loop is long.  SURVEY calls
// a sparse Cholesky on the host the pragmatic answer for this row (small, sequential, lowest GPU payoff); it is host code
// on purpose, and it is the one entry point of the library that needs no CUDA device.
//
// Arithmetic follows the reference line by line where that decides the result (error and Jacobians of an edge,
// pose_graph.py:138-178; the anchor, :107-112; the update and the angle wrap, :121-125; the convergence test, :127-130);
// the linear solve differs (Cholesky instead of LU with partial pivoting), so poses agree to rounding, not bit for bit.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "icp_b200.h"

namespace icpb {
void set_error(const char* fmt, ...);          // csrc/api.cu: the text behind icpb200_last_error()
}

namespace {

const double kPi = 3.14159265358979323846;

// pose_graph.py:15-17: (a + pi) % (2 pi) - pi with Python's floored modulo
inline double wrap_angle(double a) {
    const double two_pi = 2.0 * kPi;
    double r = fmod(a + kPi, two_pi);
    if (r < 0.0) r += two_pi;
    return r - kPi;
}

struct Mat3 {
    double m[9];
};

inline void zero(Mat3& a) { memset(a.m, 0, sizeof(a.m)); }

// c += a^T w b
inline void add_atwb(Mat3& c, const double* a, const double* w, const double* b) {
    double wa[9];                                   // w b
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) wa[3 * i + j] = w[3 * i] * b[j] + w[3 * i + 1] * b[3 + j] + w[3 * i + 2] * b[6 + j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c.m[3 * i + j] += a[i] * wa[j] + a[3 + i] * wa[3 + j] + a[6 + i] * wa[6 + j];
}

// lower Cholesky factor of a 3 x 3 block in place (upper part zeroed); false if it is not positive definite
inline bool chol3(Mat3& a) {
    double* m = a.m;
    if (!(m[0] > 0.0)) return false;
    m[0] = sqrt(m[0]);
    m[3] /= m[0]; m[6] /= m[0];
    m[4] -= m[3] * m[3];
    if (!(m[4] > 0.0)) return false;
    m[4] = sqrt(m[4]);
    m[7] = (m[7] - m[6] * m[3]) / m[4];
    m[8] -= m[6] * m[6] + m[7] * m[7];
    if (!(m[8] > 0.0)) return false;
    m[8] = sqrt(m[8]);
    m[1] = m[2] = m[5] = 0.0;
    return true;
}

// x <- x L^-T for a lower-triangular L (row-wise: solves y L^T = x for every row of the block)
inline void right_solve_lt(Mat3& x, const Mat3& l) {
    for (int r = 0; r < 3; ++r) {
        double* v = x.m + 3 * r;
        v[0] = v[0] / l.m[0];
        v[1] = (v[1] - v[0] * l.m[3]) / l.m[4];
        v[2] = (v[2] - v[0] * l.m[6] - v[1] * l.m[7]) / l.m[8];
    }
}

// c -= a b^T
inline void sub_abt(Mat3& c, const Mat3& a, const Mat3& b) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            c.m[3 * i + j] -= a.m[3 * i] * b.m[3 * j] + a.m[3 * i + 1] * b.m[3 * j + 1] + a.m[3 * i + 2] * b.m[3 * j + 2];
}

struct Skyline {
    int n = 0;
    std::vector<int> first;             // first nonzero block column of row k (<= k)
    std::vector<size_t> at;             // offset of row k's blocks: columns first[k] .. k
    std::vector<Mat3> blk;
    Mat3& operator()(int row, int col) { return blk[at[row] + (size_t)(col - first[row])]; }
};

}  // namespace

extern "C" int icpb200_pose_graph_optimize(int64_t n_nodes, double* poses, int64_t n_edges, const int32_t* edge_i,
                                           const int32_t* edge_j, const double* meas, const double* info,
                                           int n_iterations, int fix_node, double convergence_eps, int32_t* iters_out,
                                           double* step_norm_out, int32_t* status_out) {
    if (iters_out) *iters_out = 0;
    if (step_norm_out) *step_norm_out = 0.0;
    if (status_out) *status_out = ICPB200_CONVERGED;
    if (n_nodes < 0 || n_edges < 0 || (n_nodes > 0 && !poses) || (n_edges > 0 && (!edge_i || !edge_j || !meas || !info))) {
        icpb::set_error("icpb200_pose_graph_optimize: bad argument (n_nodes=%lld, n_edges=%lld)", (long long)n_nodes, (long long)n_edges);
        return ICPB200_ERR_ARG;
    }
    if (n_nodes < 2 || n_edges == 0) return ICPB200_OK;                       // pose_graph.py:90-92
    if (fix_node < 0 || fix_node >= n_nodes) {
        icpb::set_error("icpb200_pose_graph_optimize: fix_node %d is not a node (0..%lld)", fix_node, (long long)n_nodes - 1);
        return ICPB200_ERR_ARG;
    }
    const int n = (int)n_nodes;
    for (int64_t e = 0; e < n_edges; ++e)
        if (edge_i[e] < 0 || edge_i[e] >= n || edge_j[e] < 0 || edge_j[e] >= n) {
            icpb::set_error("icpb200_pose_graph_optimize: edge %lld joins %d and %d, nodes are 0..%d", (long long)e, edge_i[e], edge_j[e], n - 1);
            return ICPB200_ERR_ARG;
        }

    Skyline h;
    h.n = n;
    h.first.resize(n);
    for (int k = 0; k < n; ++k) h.first[k] = k;
    for (int64_t e = 0; e < n_edges; ++e) {
        const int lo = std::min(edge_i[e], edge_j[e]), hi = std::max(edge_i[e], edge_j[e]);
        h.first[hi] = std::min(h.first[hi], lo);
    }
    h.at.resize(n + 1);
    h.at[0] = 0;
    for (int k = 0; k < n; ++k) h.at[k + 1] = h.at[k] + (size_t)(k - h.first[k] + 1);
    h.blk.resize(h.at[n]);
    std::vector<double> b(3 * (size_t)n), dx(3 * (size_t)n);

    int status = ICPB200_MAX_ITER;
    double step_norm = 0.0;
    int it = 0;
    for (; it < n_iterations; ++it) {
        for (auto& m : h.blk) zero(m);
        std::fill(b.begin(), b.end(), 0.0);
        for (int64_t e = 0; e < n_edges; ++e) {
            // pose_graph.py:138-178
            const int i = edge_i[e], j = edge_j[e];
            const double* xi = poses + 3 * (size_t)i;
            const double* xj = poses + 3 * (size_t)j;
            const double* z = meas + 3 * e;
            const double* w = info + 9 * e;
            const double ci = cos(xi[2]), si = sin(xi[2]);
            const double dtx = xj[0] - xi[0], dty = xj[1] - xi[1];
            const double dth = wrap_angle(xj[2] - xi[2]);
            const double px = ci * dtx + si * dty, py = -si * dtx + ci * dty;
            const double err[3] = {px - z[0], py - z[1], wrap_angle(dth - z[2])};
            const double A[9] = {-ci, -si, -si * dtx + ci * dty,
                                 si, -ci, -ci * dtx - si * dty,
                                 0.0, 0.0, -1.0};
            const double B[9] = {ci, si, 0.0, -si, ci, 0.0, 0.0, 0.0, 1.0};
            add_atwb(h(i, i), A, w, A);
            add_atwb(h(j, j), B, w, B);
            if (i > j) add_atwb(h(i, j), A, w, B);           // the lower triangle only
            else if (j > i) add_atwb(h(j, i), B, w, A);
            else { add_atwb(h(i, i), A, w, B); add_atwb(h(i, i), B, w, A); }     // (a self edge: both cross terms)
            double we[3];
            for (int r = 0; r < 3; ++r) we[r] = w[3 * r] * err[0] + w[3 * r + 1] * err[1] + w[3 * r + 2] * err[2];
            for (int r = 0; r < 3; ++r) {
                b[3 * (size_t)i + r] += A[r] * we[0] + A[3 + r] * we[1] + A[6 + r] * we[2];
                b[3 * (size_t)j + r] += B[r] * we[0] + B[3 + r] * we[1] + B[6 + r] * we[2];
            }
        }
        // pose_graph.py:107-112: the anchor -- its rows and columns cleared, 1e10 on its diagonal, no right-hand side
        for (int c = h.first[fix_node]; c <= fix_node; ++c) zero(h(fix_node, c));
        for (int k = fix_node + 1; k < n; ++k)
            if (h.first[k] <= fix_node) zero(h(k, fix_node));
        h(fix_node, fix_node).m[0] = h(fix_node, fix_node).m[4] = h(fix_node, fix_node).m[8] = 1e10;
        b[3 * (size_t)fix_node] = b[3 * (size_t)fix_node + 1] = b[3 * (size_t)fix_node + 2] = 0.0;

        // block Cholesky inside the skyline: H = L L^T
        bool ok = true;
        for (int k = 0; k < n && ok; ++k) {
            const int fk = h.first[k];
            for (int c = fk; c < k; ++c) {
                Mat3& s = h(k, c);
                for (int m = std::max(fk, h.first[c]); m < c; ++m) sub_abt(s, h(k, m), h(c, m));
                right_solve_lt(s, h(c, c));
            }
            Mat3& d = h(k, k);
            for (int m = fk; m < k; ++m) sub_abt(d, h(k, m), h(k, m));
            d.m[1] = d.m[3]; d.m[2] = d.m[6]; d.m[5] = d.m[7];       // (symmetric: only the lower part is factored)
            ok = chol3(d);
        }
        if (!ok) { status = ICPB200_SINGULAR; break; }                // pose_graph.py:115-119 (LinAlgError)
        // L y = -b, L^T dx = y
        for (int k = 0; k < n; ++k) {
            double v[3] = {-b[3 * (size_t)k], -b[3 * (size_t)k + 1], -b[3 * (size_t)k + 2]};
            for (int m = h.first[k]; m < k; ++m) {
                const double* l = h(k, m).m;
                const double* y = &dx[3 * (size_t)m];
                for (int r = 0; r < 3; ++r) v[r] -= l[3 * r] * y[0] + l[3 * r + 1] * y[1] + l[3 * r + 2] * y[2];
            }
            const double* l = h(k, k).m;
            double* y = &dx[3 * (size_t)k];
            y[0] = v[0] / l[0];
            y[1] = (v[1] - l[3] * y[0]) / l[4];
            y[2] = (v[2] - l[6] * y[0] - l[7] * y[1]) / l[8];
        }
        for (int k = n - 1; k >= 0; --k) {
            const double* l = h(k, k).m;
            double* x = &dx[3 * (size_t)k];
            x[2] = x[2] / l[8];
            x[1] = (x[1] - l[7] * x[2]) / l[4];
            x[0] = (x[0] - l[3] * x[1] - l[6] * x[2]) / l[0];
            for (int m = h.first[k]; m < k; ++m) {
                const double* lm = h(k, m).m;
                double* y = &dx[3 * (size_t)m];
                for (int c = 0; c < 3; ++c) y[c] -= lm[c] * x[0] + lm[3 + c] * x[1] + lm[6 + c] * x[2];
            }
        }
        // pose_graph.py:121-130
        double ss = 0.0;
        for (int k = 0; k < n; ++k) {
            double* p = poses + 3 * (size_t)k;
            const double* d = &dx[3 * (size_t)k];
            p[0] += d[0];
            p[1] += d[1];
            p[2] = wrap_angle(p[2] + d[2]);
            ss += d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        }
        step_norm = sqrt(ss);
        if (step_norm < convergence_eps) { status = ICPB200_CONVERGED; break; }
    }
    if (iters_out) *iters_out = it;                  // the index of the iteration that ended the loop (n_iterations at the limit)
    if (step_norm_out) *step_norm_out = step_norm;
    if (status_out) *status_out = status;
    return ICPB200_OK;
}
