// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).
//
// CPU restatement of the linear seeding stage that feeds the refinement path
// (SURVEY §8(f)-1), following the reference line by line:
//   compute_planar_homographies (no-RANSAC branch)  src/estimation/linear/intrinsicsdlt.cpp:33-86
//   symmetric_rms_px                                src/estimation/linear/intrinsicsdlt.cpp:21-31
//   zhang_intrinsics_from_hs                        src/estimation/linear/zhang.cpp:9-208
//   sanitize_intrinsics                             include/calib/estimation/common/intrinsics_utils.h:12-68
//   pose_from_homography                            src/estimation/linear/posefromhomography.cpp:12-67
//   project_to_so3                                  include/calib/estimation/common/se3_utils.h:10-19
//   estimate_intrinsics                             src/estimation/linear/intrinsicsdlt.cpp:101-145
// Eigen::JacobiSVD is restated by a one-sided Jacobi SVD (oracle_math.hpp); both deliver the
// singular vectors to rounding, which is all these functions consume.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle_api.h"
#include "oracle_math.hpp"

extern "C" int orc_homography_dlt(int32_t n, const double* x, const double* y, const double* u, const double* v, double* hmtx);
extern "C" int orc_ransac_homography(int32_t n, const double* x, const double* y, const double* u, const double* v,
                                     const orc_ransac_options* o, const int32_t* sample_idx, orc_ransac_result* res, uint8_t* inlier_mask);

namespace {

// HomographyEstimator::residual (homographyestimator.cpp:80-93)
double sym_residual(const double* H, const double* Hi, double x, double y, double u, double v) {
    const double qx = H[0] * x + H[1] * y + H[2], qy = H[3] * x + H[4] * y + H[5], qz = H[6] * x + H[7] * y + H[8];
    const double du = u - qx / qz, dv = v - qy / qz;
    const double px = Hi[0] * u + Hi[1] * v + Hi[2], py = Hi[3] * u + Hi[4] * v + Hi[5], pz = Hi[6] * u + Hi[7] * v + Hi[8];
    const double dx = x - px / pz, dy = y - py / pz;
    return std::sqrt(0.5 * (du * du + dv * dv + dx * dx + dy * dy));
}

// project_to_so3 (se3_utils.h:10-19)
void project_to_so3(const double* M, double* R) {
    std::vector<double> A(M, M + 9), V, sv;
    orc::jacobi_svd(A, 3, 3, V, sv);
    // Eigen sorts singular values descending; sigma(2,2) = -1 acts on the smallest one
    int order[3] = {0, 1, 2};
    std::sort(order, order + 3, [&](int a, int b) { return sv[a] > sv[b]; });
    double U[9], Vm[9];
    for (int jj = 0; jj < 3; ++jj) {
        const int j = order[jj];
        for (int i = 0; i < 3; ++i) { U[3 * i + jj] = sv[j] > 0 ? A[3 * i + j] / sv[j] : 0.0; Vm[3 * i + jj] = V[3 * i + j]; }
    }
    if (!(sv[order[2]] > 0)) {  // rank deficient: complete U's last column as the cross product of the first two
        U[2] = U[3] * U[7] - U[6] * U[4]; U[5] = U[6] * U[1] - U[0] * U[7]; U[8] = U[0] * U[4] - U[3] * U[1];
    }
    double Vt[9]; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Vt[3 * i + j] = Vm[3 * j + i];
    orc::mat3_mul(U, Vt, R);
    if (orc::mat3_det(R) < 0.0) {
        for (int i = 0; i < 3; ++i) Vt[6 + i] = -Vt[6 + i];
        orc::mat3_mul(U, Vt, R);
    }
}

// LLT of a symmetric 3x3 (Eigen::LLT, lower); false if not positive definite
bool llt3(const double* B, double* L) {
    std::memset(L, 0, 9 * sizeof(double));
    for (int j = 0; j < 3; ++j) {
        double s = B[3 * j + j];
        for (int k = 0; k < j; ++k) s -= L[3 * j + k] * L[3 * j + k];
        if (!(s > 0.0)) return false;
        L[3 * j + j] = std::sqrt(s);
        for (int i = j + 1; i < 3; ++i) {
            double t = B[3 * i + j];
            for (int k = 0; k < j; ++k) t -= L[3 * i + k] * L[3 * j + k];
            L[3 * i + j] = t / L[3 * j + j];
        }
    }
    return true;
}

// kmtx_from_dual_conic::try_factor (zhang.cpp:40-82)
bool try_factor(const double* B, double* K) {
    for (int i = 0; i < 9; ++i) if (!std::isfinite(B[i])) return false;
    double L[9];
    if (!llt3(B, L)) return false;
    const double U[9] = {L[0], L[3], L[6], 0, L[4], L[7], 0, 0, L[8]};  // matrixU() = L^T
    orc::mat3_inv(U, K);
    for (int i = 0; i < 9; ++i) if (!std::isfinite(K[i])) return false;
    const double k22 = K[8];
    if (std::fabs(k22) < 1e-15) return false;
    for (int i = 0; i < 9; ++i) K[i] /= k22;
    if (K[0] <= 0.0 || K[4] <= 0.0) for (int i = 0; i < 9; ++i) K[i] = -K[i];
    return true;
}

bool kmtx_from_dual_conic(const double* b, double* K) {
    double B[9] = {b[0], b[1], b[3], b[1], b[2], b[4], b[3], b[4], b[5]};
    if (try_factor(B, K)) return true;
    for (int i = 0; i < 9; ++i) B[i] = -B[i];
    return try_factor(B, K);
}

// normalize_hmtx (zhang.cpp:121-147)
void normalize_hmtx(const double* Hin, double* H) {
    std::memcpy(H, Hin, 9 * sizeof(double));
    for (int i = 0; i < 9; ++i) if (!std::isfinite(H[i])) return;
    if (H[8] < 0.0) for (int i = 0; i < 9; ++i) H[i] = -H[i];
    const double h33 = H[8];
    if (std::fabs(h33) > 1e-12) { for (int i = 0; i < 9; ++i) H[i] /= h33; return; }
    double nf = 0; for (int i = 0; i < 9; ++i) nf += H[i] * H[i];
    nf = std::sqrt(nf);
    if (nf > 1e-12) for (int i = 0; i < 9; ++i) H[i] /= nf;
}

void v_ij(const double* H, int i, int j, double* v) {
    const double h0i = H[i], h1i = H[3 + i], h2i = H[6 + i], h0j = H[j], h1j = H[3 + j], h2j = H[6 + j];
    v[0] = h0i * h0j; v[1] = h0i * h1j + h1i * h0j; v[2] = h1i * h1j;
    v[3] = h0i * h2j + h2i * h0j; v[4] = h1i * h2j + h2i * h1j; v[5] = h2i * h2j;
}

}  // namespace

extern "C" {

// The two rows a homography contributes to Zhang's design matrix (zhang.cpp:149-181).
void orc_zhang_rows(const double* hmtx, double* rows12) {
    double H[9]; normalize_hmtx(hmtx, H);
    double v12[6], v11[6], v22[6], vr[6];
    v_ij(H, 0, 1, v12); v_ij(H, 0, 0, v11); v_ij(H, 1, 1, v22);
    auto nrm = [](double* r) { double s = 0; for (int i = 0; i < 6; ++i) s += r[i] * r[i]; s = std::sqrt(s); if (s > 0) for (int i = 0; i < 6; ++i) r[i] /= s; };
    nrm(v12);
    for (int i = 0; i < 6; ++i) vr[i] = v11[i] - v22[i];
    nrm(vr);
    std::memcpy(rows12, v12, sizeof v12); std::memcpy(rows12 + 6, vr, sizeof vr);
}

// zhang_intrinsics_from_hs (zhang.cpp:183-208).  kmtx5 = fx, fy, cx, cy, skew.  Returns 1 on success.
int orc_zhang_intrinsics(int64_t m, const double* hmtx /*[m][9]*/, double* kmtx5) {
    if (m < 4) return 0;
    std::vector<double> V((size_t)2 * m * 6);
    for (int64_t k = 0; k < m; ++k) orc_zhang_rows(hmtx + 9 * k, &V[(size_t)12 * k]);
    std::vector<double> W, sv;
    orc::jacobi_svd(V, (int)(2 * m), 6, W, sv);
    int jmin = 0; for (int j = 1; j < 6; ++j) if (sv[j] < sv[jmin]) jmin = j;
    double b[6]; for (int i = 0; i < 6; ++i) b[i] = W[i * 6 + jmin];
    double K[9];
    if (!kmtx_from_dual_conic(b, K)) {
        for (int i = 0; i < 6; ++i) b[i] = -b[i];
        if (!kmtx_from_dual_conic(b, K)) return 0;
    }
    kmtx5[0] = K[0]; kmtx5[1] = K[4]; kmtx5[2] = K[2]; kmtx5[3] = K[5]; kmtx5[4] = K[1];
    return 1;
}

// sanitize_intrinsics (intrinsics_utils.h:12-62).  bounds10 = fx_min, fx_max, fy_min, fy_max, cx_min,
// cx_max, cy_min, cy_max, skew_min, skew_max.  Returns 1 if a value was modified.
int orc_sanitize_intrinsics(double* k5, const double* b) {
    int modified = 0;
    auto min_focal = [&](double v, double mn) { if (!std::isfinite(v) || v < mn) { modified = 1; return mn; } return v; };
    auto principal = [&](double v, double mn, double mx) { if (!std::isfinite(v) || v < mn || v > mx) { modified = 1; return 0.5 * (mn + mx); } return v; };
    k5[0] = min_focal(k5[0], b[0]); k5[1] = min_focal(k5[1], b[2]);
    k5[2] = principal(k5[2], b[4], b[5]); k5[3] = principal(k5[3], b[6], b[7]);
    const double smin = std::min(b[8], b[9]), smax = std::max(b[8], b[9]);
    if (!std::isfinite(k5[4]) || k5[4] < smin || k5[4] > smax) { modified = 1; k5[4] = std::min(std::max(0.0, smin), smax); }  // std::clamp(0.0, ..)
    return modified;
}

// pose_from_homography (posefromhomography.cpp:12-67).  pose12 = R row-major, t.  Returns 1 on success.
int orc_pose_from_homography(const double* k5, const double* H, double* pose12, double* scale, double* cond) {
    const double I[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    std::memcpy(pose12, I, sizeof I);
    if (!std::isfinite(k5[0]) || !std::isfinite(k5[1]) || k5[2] <= 0 || k5[3] <= 0) return 0;
    if (!std::isfinite(H[8])) return 0;
    const double K[9] = {k5[0], k5[4], k5[2], 0, k5[1], k5[3], 0, 0, 1};
    double Ki[9]; orc::mat3_inv(K, Ki);
    double Hn[9]; orc::mat3_mul(Ki, H, Hn);
    const double n1 = std::sqrt(Hn[0] * Hn[0] + Hn[3] * Hn[3] + Hn[6] * Hn[6]);
    const double n2 = std::sqrt(Hn[1] * Hn[1] + Hn[4] * Hn[4] + Hn[7] * Hn[7]);
    if (!(n1 > 1e-15) || !(n2 > 1e-15)) return 0;
    const double s = 1.0 / ((n1 + n2) * 0.5);
    if (scale) *scale = s;
    if (cond) *cond = n1 > n2 ? n1 / n2 : n2 / n1;
    double M[9];
    for (int i = 0; i < 3; ++i) { M[3 * i] = s * Hn[3 * i]; M[3 * i + 1] = s * Hn[3 * i + 1]; }
    M[2] = M[3] * M[7] - M[6] * M[4]; M[5] = M[6] * M[1] - M[0] * M[7]; M[8] = M[0] * M[4] - M[3] * M[1];
    double R[9]; project_to_so3(M, R);
    double t[3] = {s * Hn[2], s * Hn[5], s * Hn[8]};
    if (t[2] <= 0) { for (int i = 0; i < 9; ++i) R[i] = -R[i]; for (int i = 0; i < 3; ++i) t[i] = -t[i]; }
    std::memcpy(pose12, R, sizeof R); std::memcpy(pose12 + 9, t, sizeof t);
    return 1;
}

// estimate_intrinsics (intrinsicsdlt.cpp:101-145) for ONE camera, homography_ransac = nullopt.
// Per view: success flag, homography (h33 = 1), symmetric_rms_px, pose (identity for failed views or
// when the decomposition fails).  bounds10 may be NULL (no bounds).  Returns 1 on success.
int orc_estimate_intrinsics(int64_t n_views, const int64_t* view_offset, const double* x, const double* y, const double* u,
                            const double* v, const double* bounds10, double* kmtx5, int32_t* view_success, double* hmtx,
                            double* sym_rms, double* poses) {
    const double I[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    std::vector<double> valid;
    for (int64_t k = 0; k < n_views; ++k) {
        const int64_t o = view_offset[k]; const int n = (int)(view_offset[k + 1] - o);
        double* H = hmtx + 9 * k;
        view_success[k] = 0; sym_rms[k] = 0.0;
        std::memcpy(poses + 12 * k, I, sizeof I);
        for (int i = 0; i < 9; ++i) H[i] = (i % 4 == 0) ? 1.0 : 0.0;
        if (n < 4) continue;
        if (orc_homography_dlt(n, x + o, y + o, u + o, v + o, H) != 0) continue;
        if (std::fabs(H[8]) > 1e-15) { const double h = H[8]; for (int i = 0; i < 9; ++i) H[i] /= h; }
        double Hi[9]; orc::mat3_inv(H, Hi);
        double s = 0; for (int i = 0; i < n; ++i) s += sym_residual(H, Hi, x[o + i], y[o + i], u[o + i], v[o + i]);
        sym_rms[k] = std::sqrt(s / (2.0 * n));
        view_success[k] = 1;
        valid.insert(valid.end(), H, H + 9);
    }
    if (n_views == 0) return 0;
    if (!orc_zhang_intrinsics((int64_t)valid.size() / 9, valid.data(), kmtx5)) return 0;
    if (bounds10) orc_sanitize_intrinsics(kmtx5, bounds10);
    for (int64_t k = 0; k < n_views; ++k)
        if (view_success[k]) orc_pose_from_homography(kmtx5, hmtx + 9 * k, poses + 12 * k, nullptr, nullptr);
    return 1;
}

// estimate_intrinsics with IntrinsicsEstimOptions::homography_ransac (intrinsicsdlt.cpp:50-64): per view
// ransac<HomographyEstimator>(view, opts), model / h33, symmetric rms over the inliers.
int orc_estimate_intrinsics_ransac(int64_t n_views, const int64_t* view_offset, const double* x, const double* y, const double* u,
                                   const double* v, const double* bounds10, const orc_ransac_options* ro, double* kmtx5,
                                   int32_t* view_success, double* hmtx, double* sym_rms, double* poses, uint8_t* inlier_mask) {
    const double I[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    std::vector<double> valid;
    for (int64_t k = 0; k < n_views; ++k) {
        const int64_t o = view_offset[k]; const int n = (int)(view_offset[k + 1] - o);
        double* H = hmtx + 9 * k;
        view_success[k] = 0; sym_rms[k] = 0.0;
        std::memcpy(poses + 12 * k, I, sizeof I);
        for (int i = 0; i < 9; ++i) H[i] = (i % 4 == 0) ? 1.0 : 0.0;
        if (n < 4) continue;
        orc_ransac_result r;
        std::vector<uint8_t> mask(n);
        orc_ransac_homography(n, x + o, y + o, u + o, v + o, ro, nullptr, &r, mask.data());
        if (inlier_mask) std::memcpy(inlier_mask + o, mask.data(), n);
        if (!r.success) continue;
        std::memcpy(H, r.hmtx, sizeof r.hmtx);
        if (std::fabs(H[8]) > 1e-15) { const double h = H[8]; for (int i = 0; i < 9; ++i) H[i] /= h; }
        double Hi[9]; orc::mat3_inv(H, Hi);
        double s = 0; int cnt = 0;
        for (int i = 0; i < n; ++i) if (mask[i]) { s += sym_residual(H, Hi, x[o + i], y[o + i], u[o + i], v[o + i]); ++cnt; }
        sym_rms[k] = cnt ? std::sqrt(s / (2.0 * cnt)) : INFINITY;
        view_success[k] = 1;
        valid.insert(valid.end(), H, H + 9);
    }
    if (n_views == 0) return 0;
    if (!orc_zhang_intrinsics((int64_t)valid.size() / 9, valid.data(), kmtx5)) return 0;
    if (bounds10) orc_sanitize_intrinsics(kmtx5, bounds10);
    for (int64_t k = 0; k < n_views; ++k)
        if (view_success[k]) orc_pose_from_homography(kmtx5, hmtx + 9 * k, poses + 12 * k, nullptr, nullptr);
    return 1;
}

}  // extern "C"
