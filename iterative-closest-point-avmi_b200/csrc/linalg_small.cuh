// Tiny fp64 solvers run by one thread per registration step.
//
// They replace the LAPACK calls the reference reaches through numpy
// (SURVEY.md section 8(a), rows I3-I5):
//   solve3_lu        <- np.linalg.solve on the 3x3 point-to-line normal
//                       equations (/root/reference/utilities/icp.py:103-108)
//   kabsch2 / kabsch3<- np.linalg.svd + det fix (icp.py:201-207)
//   sym2_min_eigvec  <- np.linalg.eigh on the 2x2 neighbourhood covariance
//                       (icp.py:71-73)
// __host__ __device__ so tests/test_linalg_host.py can check them on the CPU
// against numpy.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define ICPB_HDL __host__ __device__ inline
#else
#define ICPB_HDL inline
#endif

namespace icpb {

// Gaussian elimination with partial pivoting (the dgesv algorithm) on a 3x3
// system, row-major a[9].  Returns 0 on success, 1 if a pivot is exactly zero
// (LAPACK info > 0 -> numpy raises LinAlgError -> the reference takes the
// identity step, icp.py:107-108).
ICPB_HDL int solve3_lu(const double* a_in, const double* b_in, double* x) {
    double a[3][3], b[3];
    for (int i = 0; i < 3; ++i) {
        b[i] = b_in[i];
        for (int j = 0; j < 3; ++j) a[i][j] = a_in[3 * i + j];
    }
    for (int k = 0; k < 3; ++k) {
        int p = k;
        double best = fabs(a[k][k]);
        for (int i = k + 1; i < 3; ++i) {
            const double v = fabs(a[i][k]);
            if (v > best) { best = v; p = i; }
        }
        if (!(best > 0.0)) return 1;                      // exactly singular (or NaN)
        if (p != k) {
            for (int j = 0; j < 3; ++j) { const double t = a[k][j]; a[k][j] = a[p][j]; a[p][j] = t; }
            const double t = b[k]; b[k] = b[p]; b[p] = t;
        }
        for (int i = k + 1; i < 3; ++i) {
            const double f = a[i][k] / a[k][k];
            for (int j = k + 1; j < 3; ++j) a[i][j] -= f * a[k][j];
            b[i] -= f * b[k];
        }
    }
    for (int i = 2; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < 3; ++j) s -= a[i][j] * x[j];
        x[i] = s / a[i][i];
    }
    return 0;
}

// The same elimination with one reciprocal per pivot (what LAPACK's dgetf2 does for pivots above the safe minimum)
// instead of six divisions: the solve sits on the critical path of every iteration of the pair kernel.
ICPB_HDL int solve3_lu_rcp(const double* a_in, const double* b_in, double* x) {
    double a[3][3], b[3], inv[3];
    for (int i = 0; i < 3; ++i) {
        b[i] = b_in[i];
        for (int j = 0; j < 3; ++j) a[i][j] = a_in[3 * i + j];
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 3; ++k) {
        int p = k;
        double best = fabs(a[k][k]);
        for (int i = k + 1; i < 3; ++i) {
            const double v = fabs(a[i][k]);
            if (v > best) { best = v; p = i; }
        }
        if (!(best > 0.0)) return 1;                      // exactly singular (or NaN)
        if (p != k) {
            for (int j = 0; j < 3; ++j) { const double t = a[k][j]; a[k][j] = a[p][j]; a[p][j] = t; }
            const double t = b[k]; b[k] = b[p]; b[p] = t;
        }
        inv[k] = 1.0 / a[k][k];
        for (int i = k + 1; i < 3; ++i) {
            const double f = a[i][k] * inv[k];
            for (int j = k + 1; j < 3; ++j) a[i][j] -= f * a[k][j];
            b[i] -= f * b[k];
        }
    }
    x[2] = b[2] * inv[2];
    x[1] = (b[1] - a[1][2] * x[2]) * inv[1];
    x[0] = (b[0] - a[0][1] * x[1] - a[0][2] * x[2]) * inv[0];
    return 0;
}

// 2-D Kabsch: w = S_c^T T_c (row-major 2x2, w[a][b] = sum s_a t_b).  The proper
// rotation maximising trace(R w^T) is R(theta) with
// theta = atan2(w01 - w10, w00 + w11); this equals V U^T with the reference's
// det < 0 fix (icp.py:203-206) -- SURVEY.md section 8(c) fact (3).
ICPB_HDL void kabsch2(const double* w, double* r) {
    const double num = w[1] - w[2], den = w[0] + w[3];
    double c = 1.0, s = 0.0;
    const double h = hypot(num, den);
    if (h > 0.0) { c = den / h; s = num / h; }
    r[0] = c; r[1] = -s; r[2] = s; r[3] = c;
}

// 3-D Kabsch by one-sided Jacobi SVD of w (row-major 3x3): w V = U S, then
// R = V U^T; if det(R) < 0 the column pair of the smallest singular value is
// negated, which is what flipping the last row of vt does when LAPACK returns
// singular values in descending order (icp.py:204-206).
// `vwarm` (optional, 9 doubles, in/out): the right singular basis of the previous
// call.  Successive ICP iterations solve nearly the same problem, and starting from
// w V_prev instead of w leaves 2-3 sweeps to do instead of 6-8 (the rotations
// accumulate in V, which stays orthogonal to a few ulp over hundreds of calls).
// A sweep is three rotations; a rotation costs two divisions, a square root and a
// reciprocal square root (the convergence test compares squares).
ICPB_HDL void kabsch3(const double* w, double* r, double* vwarm = nullptr) {
    double a[3][3], v[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) v[i][j] = vwarm ? vwarm[3 * i + j] : ((i == j) ? 1.0 : 0.0);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            a[i][j] = vwarm ? w[3 * i] * v[0][j] + w[3 * i + 1] * v[1][j] + w[3 * i + 2] * v[2][j] : w[3 * i + j];
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool big = false;                       // some pair of columns is still farther than 1e-16 from orthogonal
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 3; ++i) {
                    alpha += a[i][p] * a[i][p];
                    beta += a[i][q] * a[i][q];
                    gamma += a[i][p] * a[i][q];
                }
                const double g2 = gamma * gamma, ab = alpha * beta;
                if (g2 <= 1e-34 * ab || gamma == 0.0) continue;          // |gamma| <= 1e-17 sqrt(alpha beta)
                big = big || g2 >= 1e-32 * ab;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
#ifdef __CUDA_ARCH__
                const double c = rsqrt(1.0 + t * t), s = c * t;
#else
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#endif
                for (int i = 0; i < 3; ++i) {
                    const double ap = a[i][p], aq = a[i][q];
                    a[i][p] = c * ap - s * aq;
                    a[i][q] = s * ap + c * aq;
                    const double vp = v[i][p], vq = v[i][q];
                    v[i][p] = c * vp - s * vq;
                    v[i][q] = s * vp + c * vq;
                }
            }
        if (!big) break;
    }
    if (vwarm)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) vwarm[3 * i + j] = v[i][j];
    double sig[3], u[3][3];
    int kmin = 0;
    double smax = 0.0;
    for (int j = 0; j < 3; ++j) {
        sig[j] = sqrt(a[0][j] * a[0][j] + a[1][j] * a[1][j] + a[2][j] * a[2][j]);
        if (sig[j] < sig[kmin]) kmin = j;
        smax = fmax(smax, sig[j]);
    }
    const double tiny = 1e-14 * smax;
    int nbad = 0;
    for (int j = 0; j < 3; ++j) {
        if (sig[j] > tiny && sig[j] > 0.0) {
            for (int i = 0; i < 3; ++i) u[i][j] = a[i][j] / sig[j];
        } else {
            ++nbad;
            for (int i = 0; i < 3; ++i) u[i][j] = 0.0;
        }
    }
    if (nbad == 1) {
        // rank 2: complete U with the cross product of the two good columns
        const int j = kmin, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        u[0][j] = u[1][j1] * u[2][j2] - u[2][j1] * u[1][j2];
        u[1][j] = u[2][j1] * u[0][j2] - u[0][j1] * u[2][j2];
        u[2][j] = u[0][j1] * u[1][j2] - u[1][j1] * u[0][j2];
    } else if (nbad >= 2) {
        // rank <= 1: no unique rotation; the identity step is the benign choice
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r[3 * i + j] = (i == j) ? 1.0 : 0.0;
        return;
    }
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                r[3 * i + j] = v[i][0] * u[j][0] + v[i][1] * u[j][1] + v[i][2] * u[j][2];
        const double det = r[0] * (r[4] * r[8] - r[5] * r[7]) - r[1] * (r[3] * r[8] - r[5] * r[6]) +
                           r[2] * (r[3] * r[7] - r[4] * r[6]);
        if (det >= 0.0) break;
        for (int i = 0; i < 3; ++i) v[i][kmin] = -v[i][kmin];
    }
}

// Unit eigenvector of the smallest eigenvalue of [[a, b], [b, c]].
// eigh returns [1, 0] for a zero / isotropic matrix (SURVEY.md section 8(c) fact
// (4)); the sign of the vector is irrelevant to the point-to-line normal
// equations (fact (5)).  The result is divided by max(norm, 1e-10) as
// icp.py:74-75 does.
ICPB_HDL void sym2_min_eigvec(double a, double b, double c, double* n) {
    double vx, vy;
    if (b == 0.0) {
        if (a <= c) { vx = 1.0; vy = 0.0; } else { vx = 0.0; vy = 1.0; }
    } else {
        const double half = 0.5 * (a - c);
        const double rad = hypot(half, b);
        // lambda_min - a = -(half + rad) ;  lambda_min - c = half - rad
        const double e1x = b, e1y = -(half + rad);        // (b, lambda - a)
        const double e2x = half - rad, e2y = b;           // (lambda - c, b)
        if (fabs(e1y) >= fabs(e2x)) { vx = e1x; vy = e1y; } else { vx = e2x; vy = e2y; }
    }
    const double nn = sqrt(vx * vx + vy * vy);
    const double d = nn > 1e-10 ? nn : 1e-10;
    n[0] = vx / d;
    n[1] = vy / d;
}

}  // namespace icpb
