// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
//
// CPU restatement of the arithmetic on the calibration refinement hot path
// (reference: VitalyVorobyev/calibration).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build or call this.
//
// This header holds the scalar machinery shared by the oracle sources:
//   * Dual<N>   forward-mode dual numbers = what ceres::Jet<double,N> computes
//               (the reference differentiates every residual with
//               ceres::AutoDiffCostFunction, e.g. src/estimation/residuals/
//               intrinsicresidual.h:44-46)
//   * camera models restated from include/calib/models/*.h
//   * pose algebra restated from src/estimation/detail/observationutils.h
//   * small dense linear algebra (Cholesky, one-sided Jacobi SVD) standing in
//     for the Eigen/Ceres calls (Eigen and Ceres are absent from this image).
//
// Parity status: pinned against the reference's own synthetic-recovery tests
// (tests/unit/*_test.cpp re-expressed in oracle/refdata.cpp), against this
// image's libstdc++ std::sample, and — for the RANSAC loop — bit for bit against
// the reference's own ransac<> template compiled from /root/reference
// (oracle/_ref, ref_ransac_harness.cpp; outputs frozen in tests/golden/).  The
// minimisers on noisy data are cross-checked against OpenCV and scipy
// (tests/test_oracle_crosscheck.py).  Parity against real Ceres ITERATES
// (iteration counts, covariance digits) is UNPINNED: Ceres cannot be built here
// — see DESIGN.md.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------
// Dual numbers (ceres::Jet semantics: comparisons look at the scalar part)
// ---------------------------------------------------------------------------
template <int N>
struct Dual {
    double v;
    double d[N];
    Dual() : v(0) { for (int i = 0; i < N; ++i) d[i] = 0; }
    Dual(double c) : v(c) { for (int i = 0; i < N; ++i) d[i] = 0; }  // NOLINT
    static Dual var(double c, int k) { Dual r(c); r.d[k] = 1.0; return r; }
};
template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a) {
    Dual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.v * b.d[i] + a.d[i] * b.v; return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
    // ceres::Jet: (a/b) with derivative (da - (a/b) db) / b
    Dual<N> r; const double inv = 1.0 / b.v; r.v = a.v * inv;
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv; return r; }
template <int N> inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator+(double a, const Dual<N>& b) { return b + a; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> inline Dual<N> operator-(double a, const Dual<N>& b) { return (-b) + a; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, double b) {
    Dual<N> r; r.v = a.v * b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b; return r; }
template <int N> inline Dual<N> operator*(double a, const Dual<N>& b) { return b * a; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, double b) { return a * (1.0 / b); }
template <int N> inline Dual<N> operator/(double a, const Dual<N>& b) { return Dual<N>(a) / b; }
template <int N> inline Dual<N>& operator+=(Dual<N>& a, const Dual<N>& b) { a = a + b; return a; }
template <int N> inline Dual<N>& operator-=(Dual<N>& a, const Dual<N>& b) { a = a - b; return a; }
template <int N> inline Dual<N>& operator*=(Dual<N>& a, const Dual<N>& b) { a = a * b; return a; }
template <int N> inline bool operator<(const Dual<N>& a, const Dual<N>& b) { return a.v < b.v; }
template <int N> inline bool operator>(const Dual<N>& a, const Dual<N>& b) { return a.v > b.v; }
template <int N> inline bool operator<(const Dual<N>& a, double b) { return a.v < b; }
template <int N> inline bool operator>(const Dual<N>& a, double b) { return a.v > b; }
template <int N> inline Dual<N> sqrt(const Dual<N>& a) {
    Dual<N> r; r.v = std::sqrt(a.v); const double s = 0.5 / r.v;
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s; return r; }
template <int N> inline Dual<N> sin(const Dual<N>& a) {
    Dual<N> r; r.v = std::sin(a.v); const double c = std::cos(a.v);
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * c; return r; }
template <int N> inline Dual<N> cos(const Dual<N>& a) {
    Dual<N> r; r.v = std::cos(a.v); const double s = -std::sin(a.v);
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s; return r; }
template <int N> inline Dual<N> abs(const Dual<N>& a) { return a.v < 0.0 ? -a : a; }
template <int N> inline Dual<N> atan2(const Dual<N>& y, const Dual<N>& x) {
    // ceres::Jet atan2: d = (x dy - y dx) / (x^2 + y^2)
    Dual<N> r; r.v = std::atan2(y.v, x.v); const double t = 1.0 / (x.v * x.v + y.v * y.v);
    for (int i = 0; i < N; ++i) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) * t; return r; }
inline double sqrt(double a) { return std::sqrt(a); }
inline double sin(double a) { return std::sin(a); }
inline double cos(double a) { return std::cos(a); }
inline double abs(double a) { return std::fabs(a); }
inline double atan2(double y, double x) { return std::atan2(y, x); }
inline double scalar(double a) { return a; }
template <int N> inline double scalar(const Dual<N>& a) { return a.v; }

// ---------------------------------------------------------------------------
// Pose algebra — src/estimation/detail/observationutils.h:15-41
// ---------------------------------------------------------------------------
// quat_array_to_rotmat (observationutils.h:20-24): Eigen::Quaternion(w,x,y,z)
// .toRotationMatrix() — polynomial in a NON-normalised quaternion.
template <class T> inline void quat_to_rotmat(const T* q, T* R /*row-major 3x3*/) {
    const T w = q[0], x = q[1], y = q[2], z = q[3];
    const T tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const T twx = tx * w, twy = ty * w, twz = tz * w;
    const T txx = tx * x, txy = ty * x, txz = tz * x;
    const T tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
    R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}
// invert_transform (observationutils.h:26-32)
template <class T> inline void invert_transform(const T* R, const T* t, T* Ri, T* ti) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Ri[3 * i + j] = R[3 * j + i];
    for (int i = 0; i < 3; ++i) ti[i] = -(Ri[3 * i] * t[0] + Ri[3 * i + 1] * t[1] + Ri[3 * i + 2] * t[2]);
}
// product (observationutils.h:34-41)
template <class T> inline void se3_product(const T* R1, const T* t1, const T* R2, const T* t2, T* R, T* t) {
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)
            R[3 * i + j] = R1[3 * i] * R2[j] + R1[3 * i + 1] * R2[3 + j] + R1[3 * i + 2] * R2[6 + j];
        t[i] = R1[3 * i] * t2[0] + R1[3 * i + 1] * t2[1] + R1[3 * i + 2] * t2[2] + t1[i];
    }
}

// ---------------------------------------------------------------------------
// Camera models — include/calib/models/{camera_matrix,distortion,pinhole,scheimpflug}.h
// intr layout [fx,fy,cx,cy,skew,k1,k2,k3,p1,p2(,tau_x,tau_y)] (pinhole.h:125-146,
// scheimpflug.h:241-259)
// ---------------------------------------------------------------------------
// apply_distortion (distortion.h:91-116) with 3 radial + 2 tangential coeffs.
template <class T> inline void bc5_distort(const T& x, const T& y, const T* k /*k1,k2,k3,p1,p2*/, T& xd, T& yd) {
    T r2 = x * x + y * y;
    T radial = T(1.0);
    T rpow = r2;
    for (int i = 0; i < 3; ++i) { radial += k[i] * rpow; rpow *= r2; }
    const T& p1 = k[3]; const T& p2 = k[4];
    xd = x * radial + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
    yd = y * radial + p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
}
// PinholeCamera::project(Vec3) (pinhole.h:102-107): hnormalized -> distort -> denormalize
// (camera_matrix.h:42-46).
template <class T> inline void pinhole_project(const T* intr, const T* P, T& u, T& v) {
    T x = P[0] / P[2], y = P[1] / P[2];
    T xd, yd; bc5_distort(x, y, intr + 5, xd, yd);
    u = intr[0] * xd + intr[4] * yd + intr[2];
    v = intr[1] * yd + intr[3];
}
// ScheimpflugCamera::project (scheimpflug.h:139-181)
template <class T> inline void scheimpflug_project(const T* intr, const T* P, T& u, T& v) {
    const T tx = intr[10], ty = intr[11];
    const T ctx = cos(tx), stx = sin(tx), cty = cos(ty), sty = sin(ty);
    // rot_sensor rows: [cy, sx*sy, cx*sy; 0, cx, -sx; -sy, sx*cy, cx*cy]; columns a, b, n
    const T a[3] = {cty, T(0.0), -sty};
    const T b[3] = {stx * sty, ctx, stx * cty};
    const T n[3] = {ctx * sty, -stx, ctx * cty};
    T sden = n[0] * P[0] + n[1] * P[1] + n[2] * P[2];
    T mx = (a[0] * P[0] + a[1] * P[1] + a[2] * P[2]) / sden;
    T my = (b[0] * P[0] + b[1] * P[1] + b[2] * P[2]) / sden;
    const T s0 = n[2];
    const T mx0 = a[2] / s0, my0 = b[2] / s0;
    T d[3] = {mx - mx0, my - my0, T(1.0)};
    T pu, pv; pinhole_project(intr, d, pu, pv);
    // apply_linear_intrinsics (pinhole.h:148-153): fx,fy,skew only, no cx,cy
    u = pu + (intr[0] * mx0 + intr[4] * my0);
    v = pv + (intr[1] * my0);
}

// ---------------------------------------------------------------------------
// Dense helpers (row-major) standing in for Eigen / Ceres linear algebra
// ---------------------------------------------------------------------------
// In-place lower Cholesky A = L L^T on the lower triangle; false if not PD.
inline bool cholesky(double* A, int n) {
    for (int j = 0; j < n; ++j) {
        double s = A[j * n + j];
        for (int k = 0; k < j; ++k) s -= A[j * n + k] * A[j * n + k];
        if (!(s > 0.0) || !std::isfinite(s)) return false;
        const double l = std::sqrt(s);
        A[j * n + j] = l;
        for (int i = j + 1; i < n; ++i) {
            double t = A[i * n + j];
            for (int k = 0; k < j; ++k) t -= A[i * n + k] * A[j * n + k];
            A[i * n + j] = t / l;
        }
    }
    return true;
}
inline void cholesky_solve(const double* L, int n, double* b) {
    for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i * n + k] * b[k]; b[i] = s / L[i * n + i]; }
    for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * b[k]; b[i] = s / L[i * n + i]; }
}
// One-sided Jacobi (Hestenes) SVD of A (m x n, row-major, m >= n is not
// required).  On exit the columns of A are U*S, V (n x n) holds the right
// singular vectors as columns, sv the singular values (unsorted).
inline void jacobi_svd(std::vector<double>& A, int m, int n, std::vector<double>& V, std::vector<double>& sv) {
    V.assign(static_cast<size_t>(n) * n, 0.0);
    for (int i = 0; i < n; ++i) V[i * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < n - 1; ++p) for (int q = p + 1; q < n; ++q) {
            double app = 0, aqq = 0, apq = 0;
            for (int i = 0; i < m; ++i) { const double x = A[i * n + p], y = A[i * n + q]; app += x * x; aqq += y * y; apq += x * y; }
            if (apq == 0.0) continue;
            if (std::fabs(apq) <= 1e-300 || std::fabs(apq) <= DBL_EPSILON * 1e-3 * std::sqrt(app * aqq)) continue;
            off = std::max(off, std::fabs(apq) / std::sqrt(app * aqq));
            const double zeta = (aqq - app) / (2.0 * apq);
            const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
            const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
            for (int i = 0; i < m; ++i) { const double x = A[i * n + p], y = A[i * n + q]; A[i * n + p] = c * x - s * y; A[i * n + q] = s * x + c * y; }
            for (int i = 0; i < n; ++i) { const double x = V[i * n + p], y = V[i * n + q]; V[i * n + p] = c * x - s * y; V[i * n + q] = s * x + c * y; }
        }
        if (off < 1e-15) break;
    }
    sv.assign(n, 0.0);
    for (int j = 0; j < n; ++j) { double s = 0; for (int i = 0; i < m; ++i) s += A[i * n + j] * A[i * n + j]; sv[j] = std::sqrt(s); }
}

inline void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
        C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
inline double mat3_det(const double* M) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}
// Eigen Matrix3d::inverse() = cofactor / determinant
inline void mat3_inv(const double* M, double* I) {
    const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
    const double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
    const double id = 1.0 / det;
    I[0] = c00 * id; I[1] = (M[2] * M[7] - M[1] * M[8]) * id; I[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    I[3] = c01 * id; I[4] = (M[0] * M[8] - M[2] * M[6]) * id; I[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    I[6] = c02 * id; I[7] = (M[1] * M[6] - M[0] * M[7]) * id; I[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}
// Eigen::Quaterniond(Matrix3d) (used by populate_quat_tran, observationutils.h:43-48)
inline void rotmat_to_quat(const double* R, double* q /*w,x,y,z*/) {
    double t = R[0] + R[4] + R[8];
    if (t > 0.0) {
        t = std::sqrt(t + 1.0); q[0] = 0.5 * t; t = 0.5 / t;
        q[1] = (R[7] - R[5]) * t; q[2] = (R[2] - R[6]) * t; q[3] = (R[3] - R[1]) * t;
    } else {
        int i = 0; if (R[4] > R[0]) i = 1; if (R[8] > R[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        q[1 + i] = 0.5 * t; t = 0.5 / t;
        q[0] = (R[3 * k + j] - R[3 * j + k]) * t;
        q[1 + j] = (R[3 * j + i] + R[3 * i + j]) * t;
        q[1 + k] = (R[3 * k + i] + R[3 * i + k]) * t;
    }
}

}  // namespace orc
