// C++20 client of the drop-in adapter (include/calib_b200_adapter.hpp): the calls a user of the reference makes —
// estimate_intrinsics, estimate_planar_pose, optimize_intrinsics — with the reference's own types and signatures,
// served by libcalib_b200.so.  Build:
//   g++ -std=c++20 -Iinclude examples/cpp_adapter_example.cpp -Lcalibration_b200/_build -lcalib_b200
//       -Wl,-rpath,$PWD/calibration_b200/_build -o examples/_build/cpp_adapter_example
#include <cstdio>
#include <random>

#include "calib_b200_adapter.hpp"

using namespace calib;

int main() {
    // ground truth: pinhole + Brown-Conrady camera, a 9 x 6 board (30 mm) seen from 20 random poses, 0.2 px noise
    PinholeCamera<BrownConradyd> cam_gt;
    cam_gt.kmtx = CameraMatrix{1000, 1005, 640, 360, 0.0};
    cam_gt.distortion.coeffs = Eigen::VectorXd::Zero(5);
    cam_gt.distortion.coeffs << -0.12, 0.02, 0.0005, -0.0007, 0.001;
    std::mt19937 gen(7);
    std::uniform_real_distribution<double> tilt(-0.7, 0.7), axis(-1.0, 1.0), shift(-0.15, 0.15), depth(0.6, 1.0);
    std::normal_distribution<double> noise(0.0, 0.2);
    std::vector<PlanarView> views;
    for (int v = 0; v < 20; ++v) {
        Eigen::Isometry3d c_se3_t = Eigen::Translation3d(shift(gen), shift(gen), depth(gen)) *
                                    Eigen::AngleAxisd(tilt(gen), Eigen::Vector3d(axis(gen), axis(gen), 0.2).normalized());
        PlanarView view;
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 9; ++c) {
                const Eigen::Vector2d xy((c - 4) * 0.03, (r - 2.5) * 0.03);
                Eigen::Vector2d uv = cam_gt.project(c_se3_t * Eigen::Vector3d(xy.x(), xy.y(), 0.0));
                uv += Eigen::Vector2d(noise(gen), noise(gen));
                view.push_back({xy, uv});
            }
        views.push_back(std::move(view));
    }

    try {
        // linear seed (batched per-view DLT + Zhang on the GPU), then the refinement (K1 + per-view Schur + host LM)
        const IntrinsicsEstimateResult seed = estimate_intrinsics(views);
        if (!seed.success) { std::puts("linear seed failed"); return 1; }
        PinholeCamera<BrownConradyd> guess(seed.kmtx, Eigen::VectorXd::Zero(5));
        std::vector<Eigen::Isometry3d> poses;
        for (const auto& view : views) poses.push_back(estimate_planar_pose(view, guess.kmtx));
        const auto res = optimize_intrinsics(views, guess, poses);
        std::printf("%s\n", res.core.report.c_str());
        std::printf("fx %.3f fy %.3f cx %.3f cy %.3f   k1 %.5f k2 %.5f   sigma(fx) at unit pixel noise %.3f\n", res.camera.kmtx.fx, res.camera.kmtx.fy,
                    res.camera.kmtx.cx, res.camera.kmtx.cy, res.camera.distortion.coeffs[0], res.camera.distortion.coeffs[1],
                    res.core.covariance.rows() ? std::sqrt(res.core.covariance(0, 0)) : -1.0);
        return res.core.success && std::abs(res.camera.kmtx.fx - 1000.0) < 10.0 ? 0 : 1;
    } catch (const std::exception& e) {  // std::invalid_argument / std::runtime_error, as the reference throws
        std::printf("error: %s\n", e.what());
        return 2;
    }
}
