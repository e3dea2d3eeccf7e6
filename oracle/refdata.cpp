// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).
//
// Reproduces, with this toolchain's libstdc++ <random>, the RNG-dependent part
// of the data the reference's own unit tests synthesise, so the tests in
// tests/ can re-express them with the same seeds, sizes and tolerances
// (SURVEY §4.1):
//   RNG / SimulatedHandEye::make_sequence   tests/unit/utils.h:163-221
//   generate_synthetic_data + outliers      tests/unit/homography_test.cpp:21-47,111-118
// and restates the linear seed the intrinsics tests start from:
//   estimate_planar_pose(view, CameraMatrix) src/estimation/linear/planarpose_linear.cpp:17-76
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

#include "oracle_math.hpp"

namespace {

// tests/unit/utils.h:163-181
struct RNG {
    std::mt19937 gen;
    explicit RNG(uint32_t seed) : gen(seed) {}
    double uni(double a, double b) { std::uniform_real_distribution<double> d(a, b); return d(gen); }
    double gauss(double s) { std::normal_distribution<double> n(0.0, s); return n(gen); }
    void rand_unit_axis(double* o) {
        double z = uni(-1.0, 1.0);
        double t = uni(0.0, 2.0 * M_PI);
        double r = std::sqrt(1.0 - z * z);
        o[0] = r * std::cos(t); o[1] = r * std::sin(t); o[2] = z;
    }
};

// Eigen::AngleAxisd(angle, axis).toRotationMatrix()
void angle_axis_to_R(const double* axis, double angle, double* R) {
    const double s = std::sin(angle), c = std::cos(angle);
    const double sa[3] = {s * axis[0], s * axis[1], s * axis[2]};
    const double ca[3] = {(1 - c) * axis[0], (1 - c) * axis[1], (1 - c) * axis[2]};
    double tmp;
    tmp = ca[0] * axis[1]; R[1] = tmp - sa[2]; R[3] = tmp + sa[2];
    tmp = ca[0] * axis[2]; R[2] = tmp + sa[1]; R[6] = tmp - sa[1];
    tmp = ca[1] * axis[2]; R[5] = tmp - sa[0]; R[7] = tmp + sa[0];
    R[0] = ca[0] * axis[0] + c; R[4] = ca[1] * axis[1] + c; R[8] = ca[2] * axis[2] + c;
}

}  // namespace

extern "C" {

// RNG rng(seed); n_pre x rand_unit_axis(); make_sequence(n_frames) (utils.h:203-221);
// n_post x rand_unit_axis().  b_se3_g: [n_frames][12] (R row-major, t).
void orc_ref_handeye_sequence(uint32_t seed, int n_pre, double* pre_axes, int n_frames, double* b_se3_g, int n_post,
                              double* post_axes) {
    RNG rng(seed);
    for (int i = 0; i < n_pre; ++i) rng.rand_unit_axis(pre_axes + 3 * i);
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
    for (int k = 0; k < n_frames; ++k) {
        std::memcpy(b_se3_g + 12 * k, R, sizeof R); std::memcpy(b_se3_g + 12 * k + 9, t, sizeof t);
        if (k + 1 < n_frames) {
            const double ang = rng.uni(5.0, 25.0) * M_PI / 180.0;
            double ax[3]; rng.rand_unit_axis(ax);
            // `Eigen::Vector3d dt(rng.uni(..), rng.uni(..), rng.uni(..))`: GCC evaluates
            // the constructor arguments right to left, so the first draw lands in z.
            double dt[3]; dt[2] = rng.uni(-0.10, 0.10); dt[1] = rng.uni(-0.10, 0.10); dt[0] = rng.uni(-0.10, 0.10);
            // make_pose(dt, ax, ang): axis_angle_to_R normalises the axis (utils.h:46-49)
            const double nrm = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
            double an[3] = {ax[0] / nrm, ax[1] / nrm, ax[2] / nrm}, dR[9];
            if (ang < 1e-16) { const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; std::memcpy(dR, I, sizeof I); }
            else angle_axis_to_R(an, ang, dR);
            double Rn[9], tn[3]; orc::se3_product(R, t, dR, dt, Rn, tn);  // T = T * d
            std::memcpy(R, Rn, sizeof R); std::memcpy(t, tn, sizeof t);
        }
    }
    for (int i = 0; i < n_post; ++i) rng.rand_unit_axis(post_axes + 3 * i);
}

// Gaussian pixel noise exactly as SimulatedHandEye::render_pixels draws it
// (utils.h:233-250): continues an RNG(seed) after `skip_sequence_frames` frames
// of make_sequence, two gauss() per point.
void orc_ref_gauss_stream(uint32_t seed, int skip_uniforms, double sigma, int64_t n, double* out) {
    RNG rng(seed);
    for (int i = 0; i < skip_uniforms; ++i) (void)rng.uni(0.0, 1.0);
    for (int64_t i = 0; i < n; ++i) out[i] = rng.gauss(sigma);
}

// generate_synthetic_data(view, H, n_points, noise) with std::mt19937(42)
// (homography_test.cpp:21-47), then n_out outliers from std::mt19937(seed_out)
// (homography_test.cpp:111-118).  xyuv: [(n_points + n_out)][4].
void orc_ref_homography_data(int n_points, double noise, int n_out, uint32_t seed_out, double* H, double* xyuv) {
    const double angle = 0.1, c = std::cos(angle), s = std::sin(angle);
    const double Ht[9] = {c, -s, 10.0, s, c, -5.0, 0.001, -0.002, 1.0};
    std::memcpy(H, Ht, sizeof Ht);
    std::mt19937 rng(42);
    std::uniform_real_distribution<double> dist(-100.0, 100.0);
    std::normal_distribution<double> nz(0.0, noise > 0 ? noise : 1.0);
    for (int i = 0; i < n_points; ++i) {
        // `const Vec2 point(dist(rng), dist(rng))`: argument evaluation order is
        // unspecified; GCC evaluates right to left, so the first draw is y.
        const double b = dist(rng), a = dist(rng);
        const double px = a, py = b;
        const double qx = Ht[0] * px + Ht[1] * py + Ht[2], qy = Ht[3] * px + Ht[4] * py + Ht[5], qz = Ht[6] * px + Ht[7] * py + Ht[8];
        double u = qx / qz, v = qy / qz;
        if (noise > 0) { const double n2 = nz(rng), n1 = nz(rng); u += n1; v += n2; }
        xyuv[4 * i] = px; xyuv[4 * i + 1] = py; xyuv[4 * i + 2] = u; xyuv[4 * i + 3] = v;
    }
    std::mt19937 rng2(seed_out);
    std::uniform_real_distribution<double> d2(-100.0, 100.0);
    for (int i = 0; i < n_out; ++i) {
        const double sy = d2(rng2), sx = d2(rng2);
        const double dy = d2(rng2), dx = d2(rng2);
        double* o = xyuv + 4 * (n_points + i);
        o[0] = sx; o[1] = sy; o[2] = dx; o[3] = dy;
    }
}

int orc_homography_dlt(int32_t n, const double* x, const double* y, const double* u, const double* v, double* hmtx);

// PlaneFit.RansacRejectsOutliers data (planefit_test.cpp:24-46): std::mt19937(1337), 100 points on the
// plane through (0,0,1) with normal (0.2,-0.3,1), then 40 outliers in [5,10]^3.  xyz: [140][3]; plane: 4.
void orc_ref_plane_data(double* plane, double* xyz) {
    std::mt19937 rng(1337);
    std::uniform_real_distribution<double> dist_xy(-1.0, 1.0);
    const double nn = std::sqrt(0.2 * 0.2 + 0.3 * 0.3 + 1.0);
    const double gt[4] = {0.2 / nn, -0.3 / nn, 1.0 / nn, -(0.2 / nn * 0.0 + -0.3 / nn * 0.0 + 1.0 / nn * 1.0)};
    std::memcpy(plane, gt, sizeof gt);
    for (int i = 0; i < 100; ++i) {
        const double x = dist_xy(rng);
        const double y = dist_xy(rng);
        xyz[3 * i] = x; xyz[3 * i + 1] = y; xyz[3 * i + 2] = (-gt[3] - gt[0] * x - gt[1] * y) / gt[2];
    }
    std::uniform_real_distribution<double> dist_out(5.0, 10.0);
    for (int i = 0; i < 40; ++i) {
        // emplace_back(dist_out(rng), dist_out(rng), dist_out(rng)): GCC evaluates the arguments right to left
        const double z = dist_out(rng), y = dist_out(rng), x = dist_out(rng);
        xyz[3 * (100 + i)] = x; xyz[3 * (100 + i) + 1] = y; xyz[3 * (100 + i) + 2] = z;
    }
}

// estimate_planar_pose(view, CameraMatrix) — planarpose_linear.cpp:54-76 with
// pose_from_homography_normalized :17-52.  K5 = fx, fy, cx, cy, skew.
void orc_ref_estimate_planar_pose(int32_t n, const double* x, const double* y, const double* u, const double* v,
                                  const double* K5, double* pose12) {
    const double I[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    std::memcpy(pose12, I, sizeof I);
    if (n < 4) return;
    // normalize (camera_matrix.h:34-40)
    std::vector<double> un(n), vn(n);
    for (int i = 0; i < n; ++i) { const double yc = (v[i] - K5[3]) / K5[1]; un[i] = (u[i] - K5[2] - K5[4] * yc) / K5[0]; vn[i] = yc; }
    double H[9];
    if (orc_homography_dlt(n, x, y, un.data(), vn.data(), H) != 0) return;
    if (std::fabs(H[8]) > 1e-15) for (int i = 0; i < 9; ++i) H[i] /= H[8];
    const double h1[3] = {H[0], H[3], H[6]}, h2[3] = {H[1], H[4], H[7]}, h3[3] = {H[2], H[5], H[8]};
    const double n1 = std::sqrt(h1[0] * h1[0] + h1[1] * h1[1] + h1[2] * h1[2]);
    const double n2 = std::sqrt(h2[0] * h2[0] + h2[1] * h2[1] + h2[2] * h2[2]);
    double s = std::sqrt(n1 * n2); if (s < 1e-12) s = 1.0;
    double r1[3], r2[3], r3[3];
    for (int i = 0; i < 3; ++i) { r1[i] = h1[i] / s; r2[i] = h2[i] / s; }
    r3[0] = r1[1] * r2[2] - r1[2] * r2[1]; r3[1] = r1[2] * r2[0] - r1[0] * r2[2]; r3[2] = r1[0] * r2[1] - r1[1] * r2[0];
    std::vector<double> A = {r1[0], r2[0], r3[0], r1[1], r2[1], r3[1], r1[2], r2[2], r3[2]}, V, sv;
    orc::jacobi_svd(A, 3, 3, V, sv);
    double U[9]; for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) U[3 * i + j] = A[3 * i + j] / sv[j];
    double Vt[9]; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Vt[3 * i + j] = V[3 * j + i];
    double R[9]; orc::mat3_mul(U, Vt, R);
    if (orc::mat3_det(R) < 0) {
        int m = 0; for (int j = 1; j < 3; ++j) if (sv[j] < sv[m]) m = j;  // Eigen sorts descending: col(2) is the smallest
        for (int i = 0; i < 3; ++i) Vt[3 * m + i] = -Vt[3 * m + i];
        orc::mat3_mul(U, Vt, R);
    }
    double t[3] = {h3[0] / s, h3[1] / s, h3[2] / s};
    if (R[8] < 0) { for (int i = 0; i < 9; ++i) R[i] = -R[i]; for (int i = 0; i < 3; ++i) t[i] = -t[i]; }
    std::memcpy(pose12, R, sizeof R); std::memcpy(pose12 + 9, t, sizeof t);
}

}  // extern "C"
