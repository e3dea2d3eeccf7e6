// Synthetic-data helpers of the C++ host tests: what the reference's tests/unit/utils.h provides
// (RNG :167-185, SimulatedHandEye :187-252, make_pose :52-58, make_circle_poses :79-96,
// make_bundle_observations :138-164, make_scheimpflug_observations :113-136), restated on the
// adapter's types.  std::mt19937 and the libstdc++ distributions are the ones the reference's CI
// (Ubuntu, GCC) draws from, so the seeds of its tests give the same scenes here.  TEST INFRASTRUCTURE.
#pragma once
#include <cmath>
#include <numbers>
#include <random>
#include <vector>

#include "calib_b200_adapter.hpp"

using calib::BrownConradyd;
using calib::BundleObservation;
using calib::PinholeCamera;
using calib::PlanarView;
using calib::ScheimpflugCamera;

static inline double deg2rad(double d) { return d * std::numbers::pi / 180.0; }
static inline double rad2deg(double r) { return r * 180.0 / std::numbers::pi; }

inline double rotation_angle(const Eigen::Matrix3d& R) {
    const double c = std::max(-1.0, std::min(1.0, (R.trace() - 1.0) * 0.5));
    return std::acos(c);
}

// The acos form above (the reference's, tests/unit/utils.h:29-32) cannot resolve angles below sqrt(2 eps) = 2.1e-8 rad:
// a trace one ulp short of 3 already reads as 2.1e-8.  Assertions tighter than that use the atan2 form.
inline double rotation_angle_small(const Eigen::Matrix3d& R) {
    const double ax = R(2, 1) - R(1, 2), ay = R(0, 2) - R(2, 0), az = R(1, 0) - R(0, 1);
    return std::atan2(0.5 * std::sqrt(ax * ax + ay * ay + az * az), (R.trace() - 1.0) * 0.5);
}

inline PlanarView make_view(const std::vector<Eigen::Vector2d>& obj, const std::vector<Eigen::Vector2d>& img) {
    PlanarView view(obj.size());
    for (size_t i = 0; i < obj.size(); ++i) view[i] = {obj[i], img[i]};
    return view;
}

inline Eigen::Matrix3d axis_angle_to_R(const Eigen::Vector3d& axis, double angle) {
    if (angle < 1e-16) return Eigen::Matrix3d::Identity();
    return Eigen::AngleAxisd(angle, axis.normalized()).toRotationMatrix();
}

inline Eigen::Isometry3d make_pose(const Eigen::Vector3d& t, const Eigen::Vector3d& axis, double angle) {
    Eigen::Isometry3d T = Eigen::Isometry3d::Identity();
    T.linear() = axis_angle_to_R(axis, angle);
    T.translation() = t;
    return T;
}

inline std::vector<Eigen::Isometry3d> make_circle_poses(int n, double radius, double z0, double z_step, double rot_step,
                                                        double axis_z = 1.0) {
    std::vector<Eigen::Isometry3d> poses;
    for (int i = 0; i < n; ++i) {
        const double angle = i * 2.0 * std::numbers::pi / n;
        Eigen::Isometry3d T = Eigen::Isometry3d::Identity();
        T.translation() = Eigen::Vector3d(radius * std::cos(angle), radius * std::sin(angle), z0 + z_step * i);
        const Eigen::Vector3d axis(std::cos(angle), std::sin(angle), axis_z);
        T.linear() = Eigen::AngleAxisd(rot_step * i, axis.normalized()).toRotationMatrix();
        poses.push_back(T);
    }
    return poses;
}

// one observation per (robot pose, camera): c_T_t = g_T_c^-1 * b_T_g^-1 * b_T_t, every board point projected
template <class CameraT>
inline std::vector<BundleObservation> make_observations(const std::vector<CameraT>& cams, const std::vector<Eigen::Isometry3d>& g_se3_cs,
                                                        const Eigen::Isometry3d& b_se3_t, const std::vector<Eigen::Vector2d>& obj,
                                                        const std::vector<Eigen::Isometry3d>& b_se3_gs) {
    std::vector<BundleObservation> obs;
    for (const auto& btg : b_se3_gs)
        for (size_t c = 0; c < cams.size(); ++c) {
            const Eigen::Isometry3d c_se3_t = g_se3_cs[c].inverse() * btg.inverse() * b_se3_t;
            std::vector<Eigen::Vector2d> img;
            for (const auto& xy : obj) img.push_back(cams[c].project(c_se3_t * Eigen::Vector3d(xy.x(), xy.y(), 0.0)));
            obs.push_back({make_view(obj, img), btg, c});
        }
    return obs;
}
template <class D>
inline auto make_bundle_observations(const std::vector<PinholeCamera<D>>& cams, const std::vector<Eigen::Isometry3d>& g, const Eigen::Isometry3d& b,
                                     const std::vector<Eigen::Vector2d>& obj, const std::vector<Eigen::Isometry3d>& poses) {
    return make_observations(cams, g, b, obj, poses);
}
template <class D>
inline auto make_scheimpflug_observations(const std::vector<ScheimpflugCamera<PinholeCamera<D>>>& cams, const std::vector<Eigen::Isometry3d>& g,
                                          const Eigen::Isometry3d& b, const std::vector<Eigen::Vector2d>& obj,
                                          const std::vector<Eigen::Isometry3d>& poses) {
    return make_observations(cams, g, b, obj, poses);
}

struct RNG final {
    std::mt19937 gen;
    explicit RNG(uint32_t seed = 0xC001C0DE) : gen(seed) {}
    double uni(double a, double b) { return std::uniform_real_distribution<double>(a, b)(gen); }
    double gauss(double stddev) { return std::normal_distribution<double>(0.0, stddev)(gen); }
    Eigen::Vector3d rand_unit_axis() {
        const double z = uni(-1.0, 1.0);
        const double t = uni(0.0, 2.0 * std::numbers::pi);
        const double r = std::sqrt(1.0 - z * z);
        return {r * std::cos(t), r * std::sin(t), z};
    }
};

struct SimulatedHandEye final {
    SimulatedHandEye(const Eigen::Isometry3d& g_se3_c, const Eigen::Isometry3d& b_se3_t, const PinholeCamera<BrownConradyd>& cam)
        : g_se3_c_gt(g_se3_c), b_se3_t_gt(b_se3_t), cam_gt(cam) {}
    Eigen::Isometry3d g_se3_c_gt, b_se3_t_gt;
    PinholeCamera<BrownConradyd> cam_gt;
    std::vector<Eigen::Isometry3d> c_se3_t;
    std::vector<Eigen::Vector3d> obj_pts;
    std::vector<BundleObservation> observations;

    std::vector<Eigen::Isometry3d> b_se3_g() const {
        std::vector<Eigen::Isometry3d> out;
        for (const auto& o : observations) out.push_back(o.b_se3_g);
        return out;
    }
    void make_sequence(size_t n_frames, RNG& rng) {
        c_se3_t.clear();
        observations.clear();
        Eigen::Isometry3d T = Eigen::Isometry3d::Identity();
        for (size_t k = 0; k < n_frames; ++k) {
            observations.push_back({PlanarView{}, T, 0});
            c_se3_t.push_back(g_se3_c_gt.inverse() * T.inverse() * b_se3_t_gt);
            if (k + 1 < n_frames) {  // utils.h:207-211: angle, axis, then `Vector3d dt(uni, uni, uni)` — GCC evaluates
                                     // constructor arguments right to left, so the first draw lands in z (stated
                                     // explicitly here so the scene does not depend on the compiler)
                const double ang = deg2rad(rng.uni(5.0, 25.0));
                const Eigen::Vector3d ax = rng.rand_unit_axis();
                const double dz = rng.uni(-0.10, 0.10), dy = rng.uni(-0.10, 0.10), dx = rng.uni(-0.10, 0.10);
                T = T * make_pose(Eigen::Vector3d(dx, dy, dz), ax, ang);
            }
        }
    }
    void make_target_grid(int rows, int cols, double spacing) {
        obj_pts.clear();
        const double x0 = -0.5 * (cols - 1) * spacing, y0 = -0.5 * (rows - 1) * spacing;
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) obj_pts.emplace_back(x0 + c * spacing, y0 + r * spacing, 0.0);
    }
    void render_pixels(double noise_px = 0.0, RNG* rng = nullptr) {
        for (size_t k = 0; k < observations.size(); ++k) {
            auto& view = observations[k].view;
            view.clear();
            for (const auto& P : obj_pts) {
                const Eigen::Vector3d Pc = c_se3_t[k] * P;
                if (Pc.z() <= 1e-6) continue;
                Eigen::Vector2d uv = cam_gt.project(Pc);
                if (noise_px > 0.0 && rng) {
                    uv.x() += rng->gauss(noise_px);
                    uv.y() += rng->gauss(noise_px);
                }
                view.push_back({Eigen::Vector2d(P.x(), P.y()), uv});
            }
        }
    }
};
