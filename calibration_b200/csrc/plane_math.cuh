// Per-hypothesis arithmetic of the plane RANSAC kernel (ransac_plane.cu), written once as host/device
// inline functions: the kernel is the only product caller; tests/host_emul compiles the same header with
// g++ and replays the reference's loop around it on the CPU suite (test-only, never shipped).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define CALP_HD __host__ __device__ __forceinline__
#else
#define CALP_HD inline
#endif

namespace calk {

CALP_HD double plane_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

// PlaneRansacEstimator::fit and ::is_degenerate (src/estimation/linear/planefit.cpp:14-33,52-62) test the
// same cross product: false = degenerate sample (|v1 x v2| < 1e-12).  P = (n, d), n = v1 x v2 / |v1 x v2|.
CALP_HD bool plane_from_points(double x0, double y0, double z0, double x1, double y1, double z1, double x2, double y2, double z2,
                               double* P) {
    const double v1x = x1 - x0, v1y = y1 - y0, v1z = z1 - z0;
    const double v2x = x2 - x0, v2y = y2 - y0, v2z = z2 - z0;
    double nx = v1y * v2z - v1z * v2y, ny = v1z * v2x - v1x * v2z, nz = v1x * v2y - v1y * v2x;
    const double norm = sqrt(nx * nx + ny * ny + nz * nz);
    if (!(norm >= 1e-12)) return false;
    nx /= norm; ny /= norm; nz /= norm;
    P[0] = nx; P[1] = ny; P[2] = nz; P[3] = -(nx * x0 + ny * y0 + nz * z0);
    return true;
}

// n . p + d; PlaneRansacEstimator::residual (planefit.cpp:35-38) is its absolute value
CALP_HD double plane_signed(const double* P, double x, double y, double z) { return fma(P[0], x, fma(P[1], y, fma(P[2], z, P[3]))); }

// fit_plane_svd (planefit.cpp:66-84): centroid c and the scatter matrix s (xx xy xz yy yz zz) of the CENTRED
// points are accumulated by the caller; this is the O(1) part — the eigenvector of the smallest eigenvalue
// of s by cyclic Jacobi rotations (= the last right singular vector of the centred n x 3 matrix),
// plane = (n, -n . c) / |n|, sign such that the normal component of largest magnitude is positive.
struct PlaneSums { double c[3]; double s[6]; };
CALP_HD bool plane_refit_solve(const PlaneSums& r, double* P) {
    double A[3][3] = {{r.s[0], r.s[1], r.s[2]}, {r.s[1], r.s[3], r.s[4]}, {r.s[2], r.s[4], r.s[5]}};
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 30; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        const double dia = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (!(off > 1e-22 * dia)) break;  // also leaves on NaN and on the zero matrix
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q = p + 1; q < 3; ++q) {
                const int o = 3 - p - q;  // the third index
                const double apq = A[p][q];
                if (apq != 0.0) {
                    const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
                    const double c = plane_rsqrt(fma(t, t, 1.0)), s = t * c;
                    A[p][p] = fma(-t, apq, A[p][p]); A[q][q] = fma(t, apq, A[q][q]);
                    A[p][q] = 0.0; A[q][p] = 0.0;
                    const double aop = A[o][p], aoq = A[o][q];
                    A[o][p] = A[p][o] = c * aop - s * aoq;
                    A[o][q] = A[q][o] = s * aop + c * aoq;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const double vp = V[k][p], vq = V[k][q];
                        V[k][p] = c * vp - s * vq; V[k][q] = s * vp + c * vq;
                    }
                }
            }
    }
    // column of the smallest eigenvalue (ties: the first)
    double nx = V[0][0], ny = V[1][0], nz = V[2][0], lam = A[0][0];
    if (A[1][1] < lam) { lam = A[1][1]; nx = V[0][1]; ny = V[1][1]; nz = V[2][1]; }
    if (A[2][2] < lam) { nx = V[0][2]; ny = V[1][2]; nz = V[2][2]; }
    const double d = -(nx * r.c[0] + ny * r.c[1] + nz * r.c[2]);
    const double nrm = sqrt(nx * nx + ny * ny + nz * nz);
    P[0] = nx / nrm; P[1] = ny / nrm; P[2] = nz / nrm; P[3] = d / nrm;
    double big = P[0];
    if (fabs(P[1]) > fabs(big)) big = P[1];
    if (fabs(P[2]) > fabs(big)) big = P[2];
    if (big < 0.0) { P[0] = -P[0]; P[1] = -P[1]; P[2] = -P[2]; P[3] = -P[3]; }
    return isfinite(P[0]) && isfinite(P[3]);
}

}  // namespace calk
