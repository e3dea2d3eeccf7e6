// Stand-alone host types for calib_b200_adapter.hpp — used ONLY where the reference's own headers
// (Eigen, calib/...) are absent, e.g. the image this library is developed in.  Inside the reference
// tree the adapter includes the reference's headers instead and this file is not seen.
//
// Two parts, no arithmetic of the hot path in either:
//   1. the handful of Eigen types the reference's API signatures and its unit tests use
//      (fixed small matrices, VectorXd / MatrixXd, Quaterniond, AngleAxisd, Translation3d,
//      Isometry3d), under the same names, so the adapter's source is the same in both modes;
//   2. the reference's option / result / observation / camera-model types with the same names and
//      fields, each citing the reference declaration it mirrors (paths relative to /root/reference).
//      The camera models carry a double-only project() — what the reference's tests use to render
//      synthetic pixels (tests/unit/utils.h:216-236); the refinement itself never calls it.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace Eigen {

using Index = std::ptrdiff_t;
enum : int { Dynamic = -1, ColMajor = 0, RowMajor = 1 };

template <class M>
struct CommaInit {  // m << a, b, c  fills row by row, as Eigen does
    M& m;
    Index k;
    CommaInit& operator,(double v) {
        m.rowmajor_at(k++) = v;
        return *this;
    }
};

template <int R, int C>
struct Mat {
    static_assert(R > 0 && C > 0);
    double a[R * C]{};
    Mat() = default;
    Mat(double x, double y) requires(R * C == 2) : a{x, y} {}
    Mat(double x, double y, double z) requires(R * C == 3) : a{x, y, z} {}
    Mat(double x, double y, double z, double w) requires(R * C == 4 && (R == 1 || C == 1)) : a{x, y, z, w} {}

    static Mat Zero() { return Mat{}; }
    static Mat Identity() requires(R == C) {
        Mat m;
        for (int i = 0; i < R; ++i) m(i, i) = 1.0;
        return m;
    }
    static Mat UnitX() requires(C == 1 && R >= 1) { Mat m; m.a[0] = 1.0; return m; }
    static Mat UnitY() requires(C == 1 && R >= 2) { Mat m; m.a[1] = 1.0; return m; }
    static Mat UnitZ() requires(C == 1 && R >= 3) { Mat m; m.a[2] = 1.0; return m; }

    static constexpr Index rows() { return R; }
    static constexpr Index cols() { return C; }
    static constexpr Index size() { return R * C; }
    double& rowmajor_at(Index k) { return a[k]; }
    double& operator()(Index i, Index j) { return a[i * C + j]; }
    const double& operator()(Index i, Index j) const { return a[i * C + j]; }
    double& operator()(Index i) requires(R == 1 || C == 1) { return a[i]; }
    const double& operator()(Index i) const requires(R == 1 || C == 1) { return a[i]; }
    double& operator[](Index i) requires(R == 1 || C == 1) { return a[i]; }
    const double& operator[](Index i) const requires(R == 1 || C == 1) { return a[i]; }
    double& x() requires(C == 1) { return a[0]; }
    double& y() requires(C == 1 && R >= 2) { return a[1]; }
    double& z() requires(C == 1 && R >= 3) { return a[2]; }
    double& w() requires(C == 1 && R >= 4) { return a[3]; }
    const double& x() const requires(C == 1) { return a[0]; }
    const double& y() const requires(C == 1 && R >= 2) { return a[1]; }
    const double& z() const requires(C == 1 && R >= 3) { return a[2]; }
    const double& w() const requires(C == 1 && R >= 4) { return a[3]; }
    const double* data() const { return a; }

    CommaInit<Mat> operator<<(double v) {
        a[0] = v;
        return CommaInit<Mat>{*this, 1};
    }

    Mat<C, R> transpose() const {
        Mat<C, R> t;
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C; ++j) t(j, i) = (*this)(i, j);
        return t;
    }
    Mat<R, 1> col(Index j) const {
        Mat<R, 1> c;
        for (int i = 0; i < R; ++i) c.a[i] = (*this)(i, j);
        return c;
    }
    double squaredNorm() const {
        double s = 0;
        for (double v : a) s += v * v;
        return s;
    }
    double norm() const { return std::sqrt(squaredNorm()); }
    Mat normalized() const {
        const double n = norm();
        return n > 0 ? (*this) / n : *this;
    }
    double dot(const Mat& o) const {
        double s = 0;
        for (int i = 0; i < R * C; ++i) s += a[i] * o.a[i];
        return s;
    }
    Mat cross(const Mat& o) const requires(R == 3 && C == 1) {
        return Mat(a[1] * o.a[2] - a[2] * o.a[1], a[2] * o.a[0] - a[0] * o.a[2], a[0] * o.a[1] - a[1] * o.a[0]);
    }
    double trace() const requires(R == C) {
        double s = 0;
        for (int i = 0; i < R; ++i) s += (*this)(i, i);
        return s;
    }
    template <int N>
    Mat<N, 1> head() const requires(C == 1 && N <= R) {
        Mat<N, 1> h;
        for (int i = 0; i < N; ++i) h.a[i] = a[i];
        return h;
    }
    Mat<R + 1, 1> homogeneous() const requires(C == 1) {
        Mat<R + 1, 1> h;
        for (int i = 0; i < R; ++i) h.a[i] = a[i];
        h.a[R] = 1.0;
        return h;
    }
    Mat<R - 1, 1> hnormalized() const requires(C == 1 && R >= 2) {
        Mat<R - 1, 1> h;
        for (int i = 0; i < R - 1; ++i) h.a[i] = a[i] / a[R - 1];
        return h;
    }
    // Eigen's fuzzy comparison: |a - b|^2 <= prec^2 min(|a|^2, |b|^2)
    bool isApprox(const Mat& o, double prec = 1e-12) const {
        return (*this - o).squaredNorm() <= prec * prec * std::min(squaredNorm(), o.squaredNorm());
    }

    Mat operator-() const { Mat r; for (int i = 0; i < R * C; ++i) r.a[i] = -a[i]; return r; }
    Mat& operator+=(const Mat& o) { for (int i = 0; i < R * C; ++i) a[i] += o.a[i]; return *this; }
    Mat& operator-=(const Mat& o) { for (int i = 0; i < R * C; ++i) a[i] -= o.a[i]; return *this; }
    Mat& operator*=(double s) { for (double& v : a) v *= s; return *this; }
    Mat& operator/=(double s) { for (double& v : a) v /= s; return *this; }
    friend Mat operator+(Mat l, const Mat& r) { return l += r; }
    friend Mat operator-(Mat l, const Mat& r) { return l -= r; }
    friend Mat operator*(Mat l, double s) { return l *= s; }
    friend Mat operator*(double s, Mat l) { return l *= s; }
    friend Mat operator/(Mat l, double s) { return l /= s; }
};
template <int R, int K, int C>
Mat<R, C> operator*(const Mat<R, K>& l, const Mat<K, C>& r) {
    Mat<R, C> o;
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += l(i, k) * r(k, j);
            o(i, j) = s;
        }
    return o;
}

struct VectorXd {
    std::vector<double> v;
    VectorXd() = default;
    explicit VectorXd(Index n) : v(static_cast<size_t>(n), 0.0) {}
    static VectorXd Zero(Index n) { return VectorXd(n); }
    Index size() const { return static_cast<Index>(v.size()); }
    double& rowmajor_at(Index k) { return v[static_cast<size_t>(k)]; }
    double& operator[](Index i) { return v[static_cast<size_t>(i)]; }
    const double& operator[](Index i) const { return v[static_cast<size_t>(i)]; }
    double& operator()(Index i) { return v[static_cast<size_t>(i)]; }
    const double& operator()(Index i) const { return v[static_cast<size_t>(i)]; }
    CommaInit<VectorXd> operator<<(double x) {
        v.at(0) = x;
        return CommaInit<VectorXd>{*this, 1};
    }
};
struct MatrixXd {  // row-major storage; only element access, which is all the result structs need
    Index r = 0, c = 0;
    std::vector<double> v;
    MatrixXd() = default;
    MatrixXd(Index rows, Index cols) : r(rows), c(cols), v(static_cast<size_t>(rows * cols), 0.0) {}
    static MatrixXd Zero(Index rows, Index cols) { return MatrixXd(rows, cols); }
    Index rows() const { return r; }
    Index cols() const { return c; }
    Index size() const { return r * c; }
    double& operator()(Index i, Index j) { return v[static_cast<size_t>(i * c + j)]; }
    const double& operator()(Index i, Index j) const { return v[static_cast<size_t>(i * c + j)]; }
    double trace() const {
        double s = 0;
        for (Index i = 0; i < std::min(r, c); ++i) s += (*this)(i, i);
        return s;
    }
};

template <class S, int R, int C, int O = 0>
using Matrix = std::conditional_t<(R == Dynamic && C == 1), VectorXd,
                                  std::conditional_t<(R == Dynamic || C == Dynamic), MatrixXd, Mat<(R > 0 ? R : 1), (C > 0 ? C : 1)>>>;
using Vector2d = Mat<2, 1>;
using Vector3d = Mat<3, 1>;
using Vector4d = Mat<4, 1>;
using Matrix3d = Mat<3, 3>;

struct AngleAxisd {
    double ang = 0;
    Vector3d ax = Vector3d::UnitZ();
    AngleAxisd() = default;
    AngleAxisd(double angle, const Vector3d& axis) : ang(angle), ax(axis) {}
    explicit AngleAxisd(const Matrix3d& m);  // through the quaternion, as Eigen does (defined below)
    double angle() const { return ang; }
    const Vector3d& axis() const { return ax; }
    Matrix3d toRotationMatrix() const {  // Rodrigues, unit axis assumed (as Eigen does)
        const double s = std::sin(ang), c = std::cos(ang);
        const Vector3d ca = ax * (1.0 - c);
        Matrix3d m;
        double t;
        t = ca.x() * ax.y(); m(0, 1) = t - s * ax.z(); m(1, 0) = t + s * ax.z();
        t = ca.x() * ax.z(); m(0, 2) = t + s * ax.y(); m(2, 0) = t - s * ax.y();
        t = ca.y() * ax.z(); m(1, 2) = t - s * ax.x(); m(2, 1) = t + s * ax.x();
        m(0, 0) = ca.x() * ax.x() + c; m(1, 1) = ca.y() * ax.y() + c; m(2, 2) = ca.z() * ax.z() + c;
        return m;
    }
};

struct Quaterniond {
    double qw = 1, qx = 0, qy = 0, qz = 0;
    Quaterniond() = default;
    Quaterniond(double w, double x, double y, double z) : qw(w), qx(x), qy(y), qz(z) {}
    explicit Quaterniond(const Matrix3d& m) {  // Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl<.., 3, 3>
        double q[3];
        double t = m.trace();
        if (t > 0.0) {
            t = std::sqrt(t + 1.0);
            qw = 0.5 * t;
            t = 0.5 / t;
            qx = (m(2, 1) - m(1, 2)) * t; qy = (m(0, 2) - m(2, 0)) * t; qz = (m(1, 0) - m(0, 1)) * t;
        } else {
            int i = 0;
            if (m(1, 1) > m(0, 0)) i = 1;
            if (m(2, 2) > m(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
            q[i] = 0.5 * t;
            t = 0.5 / t;
            qw = (m(k, j) - m(j, k)) * t;
            q[j] = (m(j, i) + m(i, j)) * t;
            q[k] = (m(k, i) + m(i, k)) * t;
            qx = q[0]; qy = q[1]; qz = q[2];
        }
    }
    double w() const { return qw; }
    double x() const { return qx; }
    double y() const { return qy; }
    double z() const { return qz; }
    double norm() const { return std::sqrt(qw * qw + qx * qx + qy * qy + qz * qz); }
    void normalize() {
        const double n = norm();
        qw /= n; qx /= n; qy /= n; qz /= n;
    }
    Quaterniond normalized() const { Quaterniond q = *this; q.normalize(); return q; }
    Matrix3d toRotationMatrix() const {
        const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
        const double twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx;
        const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
        Matrix3d r;
        r(0, 0) = 1 - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
        r(1, 0) = txy + twz; r(1, 1) = 1 - (txx + tzz); r(1, 2) = tyz - twx;
        r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = 1 - (txx + tyy);
        return r;
    }
};

inline AngleAxisd::AngleAxisd(const Matrix3d& m) {  // Eigen/src/Geometry/AngleAxis.h, operator=(QuaternionBase)
    const Quaterniond q(m);
    double n = std::sqrt(q.x() * q.x() + q.y() * q.y() + q.z() * q.z());
    if (n < std::numeric_limits<double>::epsilon()) n = std::sqrt(q.x() * q.x() + q.y() * q.y() + q.z() * q.z());
    if (n != 0.0) {
        ang = 2.0 * std::atan2(n, std::abs(q.w()));
        if (q.w() < 0.0) n = -n;
        ax = Vector3d(q.x() / n, q.y() / n, q.z() / n);
    } else {
        ang = 0.0;
        ax = Vector3d(1.0, 0.0, 0.0);
    }
}

struct Translation3d {
    Vector3d t;
    Translation3d(double x, double y, double z) : t(x, y, z) {}
    explicit Translation3d(const Vector3d& v) : t(v) {}
};

struct Isometry3d {
    Matrix3d R = Matrix3d::Identity();
    Vector3d t;
    Isometry3d() = default;
    Isometry3d(const Matrix3d& rot, const Vector3d& tr) : R(rot), t(tr) {}
    Isometry3d(const AngleAxisd& aa) : R(aa.toRotationMatrix()) {}
    Isometry3d(const Translation3d& tr) : t(tr.t) {}
    static Isometry3d Identity() { return Isometry3d{}; }
    Matrix3d& linear() { return R; }
    const Matrix3d& linear() const { return R; }
    Matrix3d rotation() const { return R; }
    Vector3d& translation() { return t; }
    const Vector3d& translation() const { return t; }
    Isometry3d inverse() const {
        const Matrix3d rt = R.transpose();
        return Isometry3d(rt, -(rt * t));
    }
    friend Isometry3d operator*(const Isometry3d& a, const Isometry3d& b) { return Isometry3d(a.R * b.R, a.R * b.t + a.t); }
    friend Vector3d operator*(const Isometry3d& a, const Vector3d& p) { return a.R * p + a.t; }
    bool isApprox(const Isometry3d& o, double prec = 1e-12) const {
        double d = (R - o.R).squaredNorm() + (t - o.t).squaredNorm();
        double n = std::min(R.squaredNorm() + t.squaredNorm() + 1.0, o.R.squaredNorm() + o.t.squaredNorm() + 1.0);
        return d <= prec * prec * n;
    }
};
inline Isometry3d operator*(const Translation3d& a, const Isometry3d& b) { return Isometry3d(a) * b; }
inline Isometry3d operator*(const Translation3d& a, const AngleAxisd& b) { return Isometry3d(a) * Isometry3d(b); }
inline Isometry3d operator*(const Isometry3d& a, const Translation3d& b) { return a * Isometry3d(b); }
inline Isometry3d operator*(const Isometry3d& a, const AngleAxisd& b) { return a * Isometry3d(b); }

}  // namespace Eigen

namespace calib {

// ---- models -------------------------------------------------------------------------------------
// include/calib/models/camera_matrix.h:14-21
template <typename Scalar>
struct CameraMatrixT final {
    Scalar fx = Scalar(0);
    Scalar fy = Scalar(0);
    Scalar cx = Scalar(0);
    Scalar cy = Scalar(0);
    Scalar skew = Scalar(0);
};
using CameraMatrix = CameraMatrixT<double>;
// camera_matrix.h:34-46
inline auto normalize(const CameraMatrix& cam, const Eigen::Vector2d& pixel) -> Eigen::Vector2d {
    const double y = (pixel.y() - cam.cy) / cam.fy;
    const double x = (pixel.x() - cam.cx - cam.skew * y) / cam.fx;
    return {x, y};
}
inline auto denormalize(const CameraMatrix& cam, const Eigen::Vector2d& n) -> Eigen::Vector2d {
    return {cam.fx * n.x() + cam.skew * n.y() + cam.cx, cam.fy * n.y() + cam.cy};
}
// camera_matrix.h:50-72
struct CalibrationBounds final {
    double fx_min = 0.0, fx_max = 2000.0, fy_min = 0.0, fy_max = 2000.0;
    double cx_min = 0.0, cx_max = 1280.0, cy_min = 0.0, cy_max = 720.0;
    double skew_min = -0.01, skew_max = 0.01;
};

// include/calib/models/distortion.h:91-116 (apply_distortion) and :141-155 (BrownConrady)
template <typename Scalar_>
struct BrownConrady final {
    using Scalar = Scalar_;
    Eigen::VectorXd coeffs;
    BrownConrady() = default;
    explicit BrownConrady(const Eigen::VectorXd& c) : coeffs(c) {}
    auto distort(const Eigen::Vector2d& n) const -> Eigen::Vector2d {
        if (coeffs.size() < 2) throw std::runtime_error("Insufficient distortion coefficients");
        const int nr = static_cast<int>(coeffs.size()) - 2;
        const double x = n.x(), y = n.y(), r2 = x * x + y * y;
        double radial = 1.0, rpow = r2;
        for (int i = 0; i < nr; ++i) { radial += coeffs[i] * rpow; rpow *= r2; }
        const double p1 = coeffs[nr], p2 = coeffs[nr + 1];
        return {x * radial + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x), y * radial + p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y};
    }
};
using BrownConradyd = BrownConrady<double>;

template <class CamT> struct CameraTraits;

// include/calib/models/pinhole.h:30-116
template <class DistortionT>
struct PinholeCamera final {
    using Scalar = typename DistortionT::Scalar;
    CameraMatrixT<Scalar> kmtx;
    DistortionT distortion;
    PinholeCamera() = default;
    PinholeCamera(const CameraMatrixT<Scalar>& m, const Eigen::VectorXd& coeffs) : kmtx(m), distortion(coeffs) {}
    auto apply_intrinsics(const Eigen::Vector2d& px) const -> Eigen::Vector2d { return normalize(kmtx, px); }
    auto remove_intrinsics(const Eigen::Vector2d& n) const -> Eigen::Vector2d { return denormalize(kmtx, n); }
    auto project(const Eigen::Vector2d& norm_xy) const -> Eigen::Vector2d {  // pinhole.h:96-100
        return denormalize(kmtx, distortion.distort(norm_xy));
    }
    auto project(const Eigen::Vector3d& xyz) const -> Eigen::Vector2d {  // pinhole.h:102-107
        return denormalize(kmtx, distortion.distort(xyz.hnormalized()));
    }
};
template <class DistortionT>
using Camera = PinholeCamera<DistortionT>;  // pinhole.h:164-165

// pinhole.h:117-160
template <class DistortionT>
struct CameraTraits<PinholeCamera<DistortionT>> {
    static constexpr size_t param_count = 10;
    static constexpr int k_num_dist_coeffs = 5;
    template <typename T = double>
    static auto from_array(const double* intr) -> PinholeCamera<DistortionT> {
        Eigen::VectorXd dist(k_num_dist_coeffs);
        for (int i = 0; i < k_num_dist_coeffs; ++i) dist[i] = intr[5 + i];
        return PinholeCamera<DistortionT>(CameraMatrix{intr[0], intr[1], intr[2], intr[3], intr[4]}, dist);
    }
    static void to_array(const PinholeCamera<DistortionT>& cam, std::array<double, param_count>& arr) {
        arr = {cam.kmtx.fx, cam.kmtx.fy, cam.kmtx.cx, cam.kmtx.cy, cam.kmtx.skew, 0, 0, 0, 0, 0};
        for (int i = 0; i < k_num_dist_coeffs; ++i) arr[5 + i] = cam.distortion.coeffs[i];
    }
    static auto apply_linear_intrinsics(const PinholeCamera<DistortionT>& cam, const Eigen::Vector2d& m) -> Eigen::Vector2d {
        return {cam.kmtx.fx * m.x() + cam.kmtx.skew * m.y(), cam.kmtx.fy * m.y()};  // pinhole.h:148-153 (no cx, cy)
    }
};

// include/calib/models/scheimpflug.h:18-21
struct ScheimpflugAngles final {
    double tau_x{0};
    double tau_y{0};
};
// include/calib/models/scheimpflug.h:35-181
template <class CameraT>
struct ScheimpflugCamera final {
    using Scalar = typename CameraT::Scalar;
    CameraT camera;
    Scalar tau_x = Scalar(0);
    Scalar tau_y = Scalar(0);
    ScheimpflugCamera() = default;
    ScheimpflugCamera(const CameraT& cam, ScheimpflugAngles angles) : camera(cam), tau_x(angles.tau_x), tau_y(angles.tau_y) {}
    ScheimpflugCamera(CameraT cam, double tx, double ty) : camera(std::move(cam)), tau_x(tx), tau_y(ty) {}
    auto project(const Eigen::Vector3d& xc) const -> Eigen::Vector2d {  // scheimpflug.h:139-181
        const double cx = std::cos(tau_x), sx = std::sin(tau_x), cy = std::cos(tau_y), sy = std::sin(tau_y);
        const Eigen::Vector3d axis(cy, 0.0, -sy), base(sx * sy, cx, sx * cy), normal(cx * sy, -sx, cx * cy);
        const double sden = normal.dot(xc);
        const double mx = axis.dot(xc) / sden, my = base.dot(xc) / sden;
        const double s0 = normal.z(), mx0 = axis.z() / s0, my0 = base.z() / s0;
        const Eigen::Vector2d px = camera.project(Eigen::Vector3d(mx - mx0, my - my0, 1.0));
        return px + CameraTraits<CameraT>::apply_linear_intrinsics(camera, {mx0, my0});
    }
};
// scheimpflug.h:234-261
template <class CameraT>
struct CameraTraits<ScheimpflugCamera<CameraT>> {
    static constexpr size_t param_count = CameraTraits<CameraT>::param_count + 2;
    static constexpr int k_tau_x_idx = CameraTraits<CameraT>::param_count;
    static constexpr int k_tau_y_idx = CameraTraits<CameraT>::param_count + 1;
    template <typename T = double>
    static auto from_array(const double* intr) -> ScheimpflugCamera<CameraT> {
        return ScheimpflugCamera<CameraT>(CameraTraits<CameraT>::from_array(intr), intr[k_tau_x_idx], intr[k_tau_y_idx]);
    }
    static void to_array(const ScheimpflugCamera<CameraT>& cam, std::array<double, param_count>& arr) {
        std::array<double, CameraTraits<CameraT>::param_count> inner{};
        CameraTraits<CameraT>::to_array(cam.camera, inner);
        std::copy(inner.begin(), inner.end(), arr.begin());
        arr[k_tau_x_idx] = cam.tau_x;
        arr[k_tau_y_idx] = cam.tau_y;
    }
};

// include/calib/models/cameramodel.h:29-45 (reduced to what the adapter needs)
template <typename Cam>
concept camera_model = requires(const Cam& cam, Eigen::Vector3d p3) {
    typename Cam::Scalar;
    { cam.project(p3) } -> std::same_as<Eigen::Vector2d>;
    CameraTraits<Cam>::param_count;
};

// ---- observations -------------------------------------------------------------------------------
// include/calib/estimation/linear/planarpose.h:22-27
struct PlanarObservation {
    Eigen::Vector2d object_xy;
    Eigen::Vector2d image_uv;
};
using PlanarView = std::vector<PlanarObservation>;
using MulticamPlanarView = std::vector<PlanarView>;  // include/calib/estimation/linear/extrinsics.h:20

// ---- options and results ------------------------------------------------------------------------
// include/calib/estimation/optim/optimize.h:16-40
enum class OptimizerType { DEFAULT, SPARSE_SCHUR, DENSE_SCHUR, DENSE_QR };
struct OptimOptions final {
    static constexpr double k_default_epsilon = 1e-9;
    static constexpr int k_default_max_iterations = 1000;
    OptimizerType optimizer = OptimizerType::DEFAULT;
    double huber_delta = 1.0;
    double epsilon = k_default_epsilon;
    int max_iterations = k_default_max_iterations;
    bool compute_covariance = true;
    bool verbose = false;
};
struct OptimResult final {
    bool success = false;
    Eigen::MatrixXd covariance;
    std::string report = "Empty";
    double final_cost = 0.0;
};
// include/calib/estimation/common/ransac.h:22-29
struct RansacOptions final {
    int max_iters = 1000;
    double thresh = 2.0;
    int min_inliers = 12;
    double confidence = 0.99;
    uint64_t seed = 1234567;
    bool refit_on_inliers = true;
};
// include/calib/estimation/optim/intrinsics.h:13-28
struct IntrinsicsOptimOptions final {
    OptimOptions core;
    int num_radial = 2;
    bool optimize_skew = false;
    std::optional<CalibrationBounds> bounds = std::nullopt;
    std::vector<int> fixed_distortion_indices;
    std::vector<double> fixed_distortion_values;
};
template <camera_model CameraT>
struct IntrinsicsOptimizationResult final {
    OptimResult core;
    CameraT camera;
    std::vector<Eigen::Isometry3d> c_se3_t;
    std::vector<double> view_errors;
};
// include/calib/estimation/optim/extrinsics.h:14-27
template <camera_model CameraT>
struct ExtrinsicOptimizationResult final {
    OptimResult core;
    std::vector<CameraT> cameras;
    std::vector<Eigen::Isometry3d> c_se3_r;
    std::vector<Eigen::Isometry3d> r_se3_t;
};
struct ExtrinsicOptions final {
    OptimOptions core;
    bool optimize_intrinsics = true;
    bool optimize_skew = false;
    bool optimize_extrinsics = true;
};
// include/calib/estimation/optim/bundle.h:21-45
struct BundleObservation final {
    PlanarView view;
    Eigen::Isometry3d b_se3_g;
    size_t camera_index = 0;
};
struct BundleOptions final {
    OptimOptions core;
    bool optimize_intrinsics = false;
    bool optimize_skew = false;
    bool optimize_target_pose = true;
    bool optimize_hand_eye = true;
};
template <camera_model CameraT>
struct BundleResult final {
    OptimResult core;
    std::vector<CameraT> cameras;
    std::vector<Eigen::Isometry3d> g_se3_c;
    Eigen::Isometry3d b_se3_t;
};
// include/calib/estimation/optim/handeye.h:16-19
struct HandeyeResult final {
    OptimResult core;
    Eigen::Isometry3d g_se3_c;
};
// include/calib/estimation/linear/homography.h:15-20
struct HomographyResult final {
    bool success{false};
    Eigen::Matrix3d hmtx = Eigen::Matrix3d::Identity();
    std::vector<int> inliers;
    double symmetric_rms_px{0.0};
};
// include/calib/estimation/linear/planefit.h:14-19
struct PlaneRansacResult final {
    bool success{false};
    Eigen::Vector4d plane{Eigen::Vector4d::Zero()};
    std::vector<int> inliers;
    double inlier_rms{std::numeric_limits<double>::infinity()};
};
// include/calib/estimation/linear/intrinsics.h:26-53
struct IntrinsicsEstimOptions final {
    std::optional<CalibrationBounds> bounds = std::nullopt;
    std::optional<RansacOptions> homography_ransac = std::nullopt;
    bool use_skew = false;
};
struct ViewEstimateData final {
    size_t view_index = 0;
    Eigen::Isometry3d c_se3_t = Eigen::Isometry3d::Identity();
    HomographyResult homography;
    double forward_rms_px = 0.0;
};
struct IntrinsicsEstimateResult final {
    bool success{false};
    CameraMatrix kmtx;
    std::vector<double> dist = {0, 0, 0, 0};
    std::vector<ViewEstimateData> views;
    std::string log;
};

}  // namespace calib
