// Host-side arithmetic of the seeding stage (plain C++: also compiled by tests/host_emul): Zhang's closed form
// from the per-camera Gram matrix and the sanitize step.
#pragma once
#include <algorithm>
#include <cmath>

#include "../../include/calib_b200.h"

namespace {

// ---- host side ----
// Eigenvector of the smallest eigenvalue of a symmetric positive semi-definite 6x6 by cyclic Jacobi
// rotations with the relative stopping rule |a_pq| <= eps sqrt(a_pp a_qq): for graded matrices such
// as Zhang's Gram matrix (columns scale like f^2, f, 1) Jacobi resolves the small eigenpairs to the
// accuracy the column-scaled matrix allows (Demmel & Veselic), which an unscaled QR iteration would not.
void smallest_eigvec6(double A[6][6], double* vec) {
    double V[6][6];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 5; ++p) for (int q = p + 1; q < 6; ++q) {
            const double apq = A[p][q];
            if (apq == 0.0 || std::fabs(apq) <= 1e-18 * std::sqrt(std::fabs(A[p][p] * A[q][q]))) continue;
            rotated = true;
            const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
            const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < 6; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
            for (int k = 0; k < 6; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
            for (int k = 0; k < 6; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
        }
        if (!rotated) break;
    }
    int m = 0; for (int j = 1; j < 6; ++j) if (A[j][j] < A[m][m]) m = j;
    for (int i = 0; i < 6; ++i) vec[i] = V[i][m];
}

// K from b = (B11, B12, B22, B13, B23, B33), B = K^-T K^-1 up to scale and sign (zhang.cpp:29-90)
bool kmtx_from_conic(const double* bin, double* k5) {
    for (int sign = 0; sign < 2; ++sign) {
        double b[6]; for (int i = 0; i < 6; ++i) b[i] = sign ? -bin[i] : bin[i];
        bool fin = true; for (int i = 0; i < 6; ++i) fin = fin && std::isfinite(b[i]);
        if (!fin) continue;
        // B = U^T U with U upper triangular (Cholesky); U = K^-1 up to scale
        const double B[3][3] = {{b[0], b[1], b[3]}, {b[1], b[2], b[4]}, {b[3], b[4], b[5]}};
        double L[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        bool pd = true;
        for (int j = 0; j < 3 && pd; ++j) {
            double s = B[j][j]; for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
            if (!(s > 0.0)) { pd = false; break; }
            L[j][j] = std::sqrt(s);
            for (int i = j + 1; i < 3; ++i) { double t = B[i][j]; for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k]; L[i][j] = t / L[j][j]; }
        }
        if (!pd) continue;
        // U = L^T; K = U^-1 (upper triangular inverse), then K /= K22
        const double u00 = L[0][0], u01 = L[1][0], u02 = L[2][0], u11 = L[1][1], u12 = L[2][1], u22 = L[2][2];
        double k00 = 1.0 / u00, k11 = 1.0 / u11, k22 = 1.0 / u22;
        double k01 = -u01 * k00 * k11, k12 = -u12 * k11 * k22, k02 = (u01 * u12 - u02 * u11) * k00 * k11 * k22;
        if (!std::isfinite(k00) || !std::isfinite(k11) || !std::isfinite(k22) || !std::isfinite(k01) || !std::isfinite(k02) || !std::isfinite(k12)) continue;
        if (std::fabs(k22) < 1e-15) continue;
        k00 /= k22; k01 /= k22; k02 /= k22; k11 /= k22; k12 /= k22;
        if (k00 <= 0.0 || k11 <= 0.0) { k00 = -k00; k01 = -k01; k02 = -k02; k11 = -k11; k12 = -k12; }
        k5[0] = k00; k5[1] = k11; k5[2] = k02; k5[3] = k12; k5[4] = k01;
        return true;
    }
    return false;
}

void sanitize(double* k5, const cal_seed_options& o) {  // intrinsics_utils.h:12-62
    auto min_focal = [](double v, double mn) { return (!std::isfinite(v) || v < mn) ? mn : v; };
    auto principal = [](double v, double mn, double mx) { return (!std::isfinite(v) || v < mn || v > mx) ? 0.5 * (mn + mx) : v; };
    k5[0] = min_focal(k5[0], o.fx_min); k5[1] = min_focal(k5[1], o.fy_min);
    k5[2] = principal(k5[2], o.cx_min, o.cx_max); k5[3] = principal(k5[3], o.cy_min, o.cy_max);
    const double smin = std::min(o.skew_min, o.skew_max), smax = std::max(o.skew_min, o.skew_max);
    if (!std::isfinite(k5[4]) || k5[4] < smin || k5[4] > smax) k5[4] = std::min(std::max(0.0, smin), smax);
}

}  // namespace
