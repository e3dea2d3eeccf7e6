// Small dense helpers shared by the RANSAC kernel (ransac.cu) and the seeding kernels (seed.cu):
// Hartley de-normalisation of a DLT solution, the null vector of the 9x9 DLT normal matrix, 3x3
// inverse, warp sums.  FP64, fully unrolled, register resident.
#pragma once
#include <math.h>

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void inv3(const double* M, double* I) {
    const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
    const double id = 1.0 / (M[0] * c00 + M[1] * c01 + M[2] * c02);
    I[0] = c00 * id; I[1] = (M[2] * M[7] - M[1] * M[8]) * id; I[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    I[3] = c01 * id; I[4] = (M[0] * M[8] - M[2] * M[6]) * id; I[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    I[6] = c02 * id; I[7] = (M[1] * M[6] - M[0] * M[7]) * id; I[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}
// H = Td^-1 * Hn * Ts for the Hartley similarity transforms T = [s 0 -s cx; 0 s -s cy; 0 0 1]
__device__ __forceinline__ void denormalise(const double* hn, double ss, double scx, double scy, double ds, double dcx, double dcy,
                                            double* H) {
    // Hn * Ts
    double M[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        M[3 * i] = hn[3 * i] * ss; M[3 * i + 1] = hn[3 * i + 1] * ss;
        M[3 * i + 2] = hn[3 * i + 2] - ss * (hn[3 * i] * scx + hn[3 * i + 1] * scy);
    }
    // Td^-1 = [1/s 0 cx; 0 1/s cy; 0 0 1]
    const double id = 1.0 / ds;
#pragma unroll
    for (int j = 0; j < 3; ++j) { H[j] = M[j] * id + dcx * M[6 + j]; H[3 + j] = M[3 + j] * id + dcy * M[6 + j]; H[6 + j] = M[6 + j]; }
}

// smallest eigenvector of the symmetric 9x9 G (upper triangle given) by inverse iteration on G + mu I
__device__ bool smallest_eigvec9(const double* Gu /*45*/, double* z) {
    double L[9][9];
    double tr = 0;
    {
        int o = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = i; j < 9; ++j) { L[j][i] = Gu[o]; if (i == j) tr += Gu[o]; ++o; }
    }
    const double mu = 1e-14 * tr + 1e-300;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        double s = L[j][j] + mu;
#pragma unroll
        for (int k = 0; k < j; ++k) s = fma(-L[j][k], L[j][k], s);
        if (!(s > 0.0)) return false;
        const double il = rsqrt(s);  // the diagonal holds 1 / l_jj: the solves below multiply instead of dividing
        L[j][j] = il;
#pragma unroll
        for (int i = j + 1; i < 9; ++i) {
            double t = L[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) t = fma(-L[i][k], L[j][k], t);
            L[i][j] = t * il;
        }
    }
    double w[9] = {1.0, 0.7, 0.3, -0.5, 0.9, 0.2, -0.8, 0.4, 1.0};
    // (G + mu I)^-1 is positive definite, so the normalised iterates converge without sign flips; the
    // iteration stops once they are stationary to rounding (a handful of steps: the smallest
    // eigenvalue of a DLT normal matrix is separated from the next by orders of magnitude)
    for (int it = 0; it < 16; ++it) {
        double prev[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) prev[i] = w[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) { double s = w[i];
#pragma unroll
            for (int k = 0; k < i; ++k) s = fma(-L[i][k], w[k], s); w[i] = s * L[i][i]; }
#pragma unroll
        for (int i = 8; i >= 0; --i) { double s = w[i];
#pragma unroll
            for (int k = i + 1; k < 9; ++k) s = fma(-L[k][i], w[k], s); w[i] = s * L[i][i]; }
        double n2 = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) n2 = fma(w[i], w[i], n2);
        const double in = rsqrt(n2);
        double d = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) { w[i] *= in; d = fmax(d, fabs(w[i] - prev[i])); }
        if (d < 4e-16) break;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) z[i] = w[i];
    return true;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}


// A^T A of the DLT design matrix (rows [-p, 0, u p], [0, -p, v p], p = (x, y, 1)) from the 24 monomial
// sums m[k][e] = sum w_k pp_e with w = (1, u, v, u^2 + v^2), pp = (x^2, xy, x, y^2, y, 1): upper triangle.
__device__ __forceinline__ void dlt_normal_matrix(const double (*m)[6], double* Gu /*45*/) {
    auto blk = [](const double* s, int i, int j) { const int a = i < j ? i : j, b = i < j ? j : i; return s[a == 0 ? b : (a == 1 ? 2 + b : 5)]; };
    int o = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = i; j < 9; ++j) {
            const int bi = i / 3, bj = j / 3, ii = i % 3, jj = j % 3;
            double v = 0.0;
            if (bi == bj) v = blk(bi == 2 ? m[3] : m[0], ii, jj);
            else if (bj == 2) v = -blk(bi == 0 ? m[1] : m[2], ii, jj);
            Gu[o++] = v;
        }
}

}  // namespace
