// End to end from the reference's own input type: calib::optimize_bundle(std::vector<BundleObservation>, ...) through the
// drop-in adapter (include/calib_b200_adapter.hpp) at BASELINE configs[4] scale — 8 cameras x N robot poses x 88 corners,
// every view its own heap-allocated PlanarView, exactly what a caller of the reference holds.  The timed region is the
// whole call: AoS -> staging (threaded, page-locked), upload, device layout, LM solve with covariance, results back in
// the reference's result types.  Prints one JSON object.  No Python, no torch in this process.
//   examples/_build/cpp_bundle_e2e [n_poses = 100000] [repeats = 3]
#include <chrono>
#include <cstdio>
#include <random>
#include <thread>

#include "calib_b200_adapter.hpp"

using namespace calib;
using Clock = std::chrono::steady_clock;

static Eigen::Isometry3d pose(const Eigen::Vector3d& t, const Eigen::Vector3d& axis, double angle) {
    Eigen::Isometry3d T = Eigen::Isometry3d::Identity();
    T.linear() = Eigen::AngleAxisd(angle, axis.normalized()).toRotationMatrix();
    T.translation() = t;
    return T;
}

int main(int argc, char** argv) {
    const int n_poses = argc > 1 ? std::atoi(argv[1]) : 100000, repeats = argc > 2 ? std::atoi(argv[2]) : 3, n_cams = 8;
    const int rows = 8, cols = 11;
    std::vector<PinholeCamera<BrownConradyd>> cams_gt(n_cams), cams0(n_cams);
    std::vector<Eigen::Isometry3d> g_se3_c_gt(n_cams), g_se3_c0(n_cams);
    for (int c = 0; c < n_cams; ++c) {
        cams_gt[c].kmtx = CameraMatrix{1000.0 * (1 + 0.005 * c), 1005.0 * (1 + 0.005 * c), 640, 360, 0.0};
        cams_gt[c].distortion.coeffs = Eigen::VectorXd::Zero(5);
        cams_gt[c].distortion.coeffs << -0.12, 0.02, 0.0005, -0.0007, 0.001;
        cams0[c] = cams_gt[c];
        cams0[c].kmtx.fx *= 0.99; cams0[c].kmtx.fy *= 1.01; cams0[c].kmtx.cx += 2.0; cams0[c].kmtx.cy -= 1.5;
        const double a = 2.0 * 3.141592653589793 * c / n_cams;
        g_se3_c_gt[c] = pose(Eigen::Vector3d(0.05 * std::cos(a), 0.05 * std::sin(a), 0.10), Eigen::Vector3d(std::cos(a), std::sin(a), 0.3), 0.12);
        g_se3_c0[c] = g_se3_c_gt[c];
        g_se3_c0[c].translation() += Eigen::Vector3d(0.004, -0.003, 0.002);
    }
    const Eigen::Isometry3d b_se3_t_gt = pose(Eigen::Vector3d(0.5, -0.1, 0.8), Eigen::Vector3d(1, 0, 0), 0.25);
    Eigen::Isometry3d b_se3_t0 = b_se3_t_gt; b_se3_t0.translation() += Eigen::Vector3d(0.003, 0.002, -0.004);

    // observations: the robot looks at the board from random poses; every camera of the rig sees all 88 corners
    std::vector<BundleObservation> obs(static_cast<size_t>(n_poses) * n_cams);
    const auto t_gen = Clock::now();
    {
        const unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&, t] {
                std::mt19937_64 gen(1234 + t);
                std::uniform_real_distribution<double> ang(-0.35, 0.35), sh(-0.08, 0.08), dz(0.55, 0.75);
                std::normal_distribution<double> noise(0.0, 0.2);
                for (int p = static_cast<int>(static_cast<int64_t>(n_poses) * t / nt); p < static_cast<int>(static_cast<int64_t>(n_poses) * (t + 1) / nt); ++p) {
                    // camera-centric view of the board, then the robot pose that produces it for camera 0's mount
                    const Eigen::Isometry3d c_se3_t = pose(Eigen::Vector3d(sh(gen), sh(gen), dz(gen)), Eigen::Vector3d(ang(gen), ang(gen), 0.2), ang(gen));
                    const Eigen::Isometry3d b_se3_g = b_se3_t_gt * c_se3_t.inverse() * g_se3_c_gt[0].inverse();
                    for (int c = 0; c < n_cams; ++c) {
                        BundleObservation& o = obs[static_cast<size_t>(p) * n_cams + c];
                        o.b_se3_g = b_se3_g; o.camera_index = static_cast<size_t>(c);
                        const Eigen::Isometry3d cam_se3_t = g_se3_c_gt[c].inverse() * b_se3_g.inverse() * b_se3_t_gt;
                        o.view.resize(rows * cols);
                        for (int r = 0; r < rows; ++r)
                            for (int q = 0; q < cols; ++q) {
                                const Eigen::Vector2d xy((q - 5) * 0.02, (r - 3.5) * 0.02);
                                Eigen::Vector2d uv = cams_gt[c].project(cam_se3_t * Eigen::Vector3d(xy.x(), xy.y(), 0.0));
                                o.view[static_cast<size_t>(r * cols + q)] = {xy, Eigen::Vector2d(uv.x() + noise(gen), uv.y() + noise(gen))};
                            }
                    }
                }
            });
        for (auto& t : th) t.join();
    }
    const double gen_s = std::chrono::duration<double>(Clock::now() - t_gen).count();
    const size_t n_obs = obs.size() * rows * cols;

    BundleOptions opts;
    opts.optimize_intrinsics = true; opts.optimize_target_pose = true; opts.optimize_hand_eye = true;
    opts.core.huber_delta = 1.0; opts.core.compute_covariance = true;
    std::printf("{\"what\": \"calib::optimize_bundle(std::vector<BundleObservation>) through calib_b200_adapter.hpp: AoS packing (threaded, page-locked "
                "staging), upload, layout, LM solve with covariance, results in the reference's types\", \"n_cams\": %d, \"n_poses\": %d, \"observations\": %zu, "
                "\"generate_s\": %.2f, \"host_threads\": %u, \"runs\": [", n_cams, n_poses, n_obs, gen_s, std::max(1u, std::min(32u, std::thread::hardware_concurrency())));
    int rc = 0;
    for (int k = 0; k < repeats; ++k) {
        try {
            const auto t0 = Clock::now();
            const auto res = optimize_bundle<PinholeCamera<BrownConradyd>>(obs, cams0, g_se3_c0, b_se3_t0, opts);
            const double dt = std::chrono::duration<double>(Clock::now() - t0).count();
            std::printf("%s{\"wall_s\": %.5f, \"converged\": %s, \"final_cost\": %.9e, \"fx0\": %.6f, \"covariance_rows\": %d}", k ? ", " : "", dt,
                        res.core.success ? "true" : "false", res.core.final_cost, res.cameras[0].kmtx.fx, static_cast<int>(res.core.covariance.rows()));
            if (!res.core.success || std::abs(res.cameras[0].kmtx.fx - 1000.0) > 1.0) rc = 1;
        } catch (const std::exception& e) {
            std::printf("%s{\"error\": \"%s\"}", k ? ", " : "", e.what());
            rc = 2;
        }
    }
    std::printf("]}\n");
    return rc;
}
