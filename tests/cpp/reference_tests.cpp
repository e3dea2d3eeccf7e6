// C++ host tests of the drop-in boundary: the reference's own unit tests for the refinement path
// (tests/unit/*.cpp, cited per test), re-expressed against include/calib_b200_adapter.hpp — the same
// calib:: entry points, options, seeds, scenes and tolerances.  Where a reference test seeds the
// problem with a linear estimator that is outside the path (estimate_extrinsic_dlt,
// optimize_homography) the note on the test says what stands in.
//
// Linked against libcalib_b200.so this is a parity run on the GPU (tests/test_gpu_zzz_cpp_host.py);
// linked against tests/cpp/abi_standin.cpp (CPU oracle behind the same C symbols) it checks the
// adapter's packing, block order, unpacking and error mapping in the CPU suite.  TEST INFRASTRUCTURE.
#include <cstdlib>
#include <numeric>

#include "mini_gtest.hpp"
#include "ref_sim.hpp"

using namespace calib;
using Vec2 = Eigen::Vector2d;
using Mat3 = Eigen::Matrix3d;

namespace {

Camera<BrownConradyd> pinhole(double fx, double fy, double cx, double cy, double skew = 0.0) {
    Camera<BrownConradyd> cam;
    cam.kmtx = CameraMatrix{fx, fy, cx, cy, skew};
    cam.distortion.coeffs = Eigen::VectorXd::Zero(5);
    return cam;
}

std::vector<PlanarView> views_of(const SimulatedHandEye& sim) {
    std::vector<PlanarView> views;
    for (const auto& ob : sim.observations) views.push_back(ob.view);
    return views;
}

}  // namespace

// ---- tests/unit/intrinsics_optimize_test.cpp ------------------------------------------------------

TEST(OptimizeIntrinsics, RecoversIntrinsicsNoSkew) {  // :8-62
    RNG rng(7);
    const auto cam_gt = pinhole(1000, 1005, 640, 360);
    SimulatedHandEye sim{Eigen::Isometry3d::Identity(), Eigen::Translation3d(0.0, 0.0, 2.0) * Eigen::Isometry3d::Identity(), cam_gt};
    sim.make_sequence(15, rng);
    sim.make_target_grid(8, 11, 0.02);
    sim.render_pixels();
    const auto views = views_of(sim);

    auto guess_cam = cam_gt;
    guess_cam.kmtx.fx *= 0.97;
    guess_cam.kmtx.fy *= 1.03;
    guess_cam.kmtx.cx += 5.0;
    guess_cam.kmtx.cy -= 4.0;

    std::vector<Eigen::Isometry3d> init_poses;
    for (const auto& view : views) init_poses.push_back(estimate_planar_pose(view, guess_cam.kmtx));

    IntrinsicsOptimOptions opts;
    opts.num_radial = 3;
    opts.optimize_skew = false;
    auto res = optimize_intrinsics(views, guess_cam, init_poses, opts);

    const auto& k_final = res.camera.kmtx;
    const auto& k_gt = cam_gt.kmtx;
    EXPECT_NEAR(k_final.fx, k_gt.fx, 1e-6);
    EXPECT_NEAR(k_final.fy, k_gt.fy, 1e-6);
    EXPECT_NEAR(k_final.cx, k_gt.cx, 1e-6);
    EXPECT_NEAR(k_final.cy, k_gt.cy, 1e-6);
    EXPECT_NEAR(k_final.skew, k_gt.skew, 1e-9);
    EXPECT_LT(res.core.final_cost, 1e-6);
    // beyond the reference's assertions: result packaging (one pose per view, covariance in the block order of
    // IntrinsicBlocks::get_param_blocks, intrinsics.cpp:34-50: 10 + 4 n + 3 n ambient rows)
    EXPECT_TRUE(res.core.success);
    ASSERT_EQ(res.c_se3_t.size(), views.size());
    for (size_t v = 0; v < views.size(); ++v) {
        EXPECT_LT(rotation_angle_small(res.c_se3_t[v].linear().transpose() * sim.c_se3_t[v].linear()), 1e-8);
        EXPECT_LT((res.c_se3_t[v].translation() - sim.c_se3_t[v].translation()).norm(), 1e-8);
    }
    EXPECT_EQ(res.core.covariance.rows(), static_cast<Eigen::Index>(10 + 7 * views.size()));
}

TEST(OptimizeIntrinsics, RecoversSkew) {  // :64-113
    RNG rng(5);
    const auto cam_gt = pinhole(1000, 1005, 640, 360, 0.001);
    SimulatedHandEye sim{Eigen::Isometry3d::Identity(), Eigen::Translation3d(0.0, 0.0, 2.0) * Eigen::Isometry3d::Identity(), cam_gt};
    sim.make_sequence(15, rng);
    sim.make_target_grid(8, 11, 0.02);
    sim.render_pixels();
    const auto views = views_of(sim);

    auto guess_cam = cam_gt;
    guess_cam.kmtx.fx *= 0.95;
    guess_cam.kmtx.fy *= 1.05;
    guess_cam.kmtx.cx += 10.0;
    guess_cam.kmtx.cy -= 6.0;
    guess_cam.kmtx.skew = 0.0;

    std::vector<Eigen::Isometry3d> init_poses(views.size());
    std::transform(views.begin(), views.end(), init_poses.begin(), [&](const auto& view) { return estimate_planar_pose(view, guess_cam.kmtx); });

    IntrinsicsOptimOptions opts;
    opts.num_radial = 0;
    opts.optimize_skew = true;
    auto res = optimize_intrinsics(views, guess_cam, init_poses, opts);

    const auto& k_final = res.camera.kmtx;
    const auto& k_gt = cam_gt.kmtx;
    EXPECT_NEAR(k_final.fx, k_gt.fx, 1e-6);
    EXPECT_NEAR(k_final.fy, k_gt.fy, 1e-6);
    EXPECT_NEAR(k_final.cx, k_gt.cx, 1e-6);
    EXPECT_NEAR(k_final.cy, k_gt.cy, 1e-6);
    EXPECT_NEAR(k_final.skew, k_gt.skew, 1e-8);
}

// The adapter keeps ONE board while the views share it (shared-board form of cal_problem_desc) and expands to
// per-observation object points at the first view that differs: a view with corners missing must give the same
// calibration as the reference's per-observation packing would.
TEST(OptimizeIntrinsics, ViewsWithMissingCornersLeaveTheSharedBoardForm) {
    RNG rng(7);
    const auto cam_gt = pinhole(1000, 1005, 640, 360);
    SimulatedHandEye sim{Eigen::Isometry3d::Identity(), Eigen::Translation3d(0.0, 0.0, 2.0) * Eigen::Isometry3d::Identity(), cam_gt};
    sim.make_sequence(15, rng);
    sim.make_target_grid(8, 11, 0.02);
    sim.render_pixels();
    auto views = views_of(sim);
    views[4].erase(views[4].begin() + 10, views[4].begin() + 17);  // seven corners undetected in view 4
    views[9].resize(60);
    auto guess_cam = cam_gt;
    guess_cam.kmtx.fx *= 0.97;
    guess_cam.kmtx.cy -= 4.0;
    std::vector<Eigen::Isometry3d> init_poses;
    for (const auto& view : views) init_poses.push_back(estimate_planar_pose(view, guess_cam.kmtx));
    auto res = optimize_intrinsics(views, guess_cam, init_poses);
    EXPECT_NEAR(res.camera.kmtx.fx, cam_gt.kmtx.fx, 1e-6);
    EXPECT_NEAR(res.camera.kmtx.fy, cam_gt.kmtx.fy, 1e-6);
    EXPECT_NEAR(res.camera.kmtx.cx, cam_gt.kmtx.cx, 1e-6);
    EXPECT_NEAR(res.camera.kmtx.cy, cam_gt.kmtx.cy, 1e-6);
    for (size_t v = 0; v < views.size(); ++v) EXPECT_LT((res.c_se3_t[v].translation() - sim.c_se3_t[v].translation()).norm(), 1e-8);
    // the packing itself: shared while the boards agree, per-observation from the first view that differs
    b200::Soa s;
    for (size_t v = 0; v < 4; ++v) s.add(views[v], 0, static_cast<int>(v));
    if (std::getenv("CALIB_B200_PER_OBSERVATION") == nullptr) EXPECT_TRUE(s.shared && s.x.empty() && s.board_x.size() == 88);
    else EXPECT_TRUE(!s.shared && s.x.size() == 4 * 88);   // the switch forces object points per observation from the start
    s.add(views[4], 0, 4);
    EXPECT_FALSE(s.shared);
    EXPECT_EQ(s.x.size(), static_cast<size_t>(4 * 88 + 81));
    EXPECT_EQ(s.x.size(), s.u.size());
    for (size_t i = 0; i < s.x.size(); ++i) {
        const auto& ob = views[i / 88 < 4 ? i / 88 : 4][i / 88 < 4 ? i % 88 : i - 4 * 88];
        EXPECT_TRUE(s.x[i] == ob.object_xy.x() && s.y[i] == ob.object_xy.y() && s.u[i] == ob.image_uv.x());
    }
}

// Non-convergence is not an error (solve_problem, src/estimation/detail/ceresutils.h:40-42): success = false, the
// parameters reached so far are returned; compute_covariance = false leaves OptimResult::covariance empty (:109-118 of intrinsics.cpp).
TEST(OptimizeIntrinsics, NonConvergenceStaysInBand) {
    RNG rng(7);
    const auto cam_gt = pinhole(1000, 1005, 640, 360);
    SimulatedHandEye sim{Eigen::Isometry3d::Identity(), Eigen::Translation3d(0.0, 0.0, 2.0) * Eigen::Isometry3d::Identity(), cam_gt};
    sim.make_sequence(15, rng);
    sim.make_target_grid(8, 11, 0.02);
    sim.render_pixels();
    const auto views = views_of(sim);
    auto guess_cam = cam_gt;
    guess_cam.kmtx.fx *= 0.97;
    guess_cam.kmtx.fy *= 1.03;
    std::vector<Eigen::Isometry3d> init_poses;
    for (const auto& view : views) init_poses.push_back(estimate_planar_pose(view, guess_cam.kmtx));
    IntrinsicsOptimOptions opts;
    opts.core.max_iterations = 2;
    opts.core.compute_covariance = false;
    const auto res = optimize_intrinsics(views, guess_cam, init_poses, opts);
    EXPECT_FALSE(res.core.success);
    EXPECT_TRUE(res.core.report.find("NO_CONVERGENCE") != std::string::npos);
    EXPECT_EQ(res.core.covariance.rows(), static_cast<Eigen::Index>(0));
    EXPECT_GT(res.core.final_cost, 0.0);
    EXPECT_LT(std::abs(res.camera.kmtx.fx - 1000.0), 30.0);   // two iterations in: better than the start, not yet there
    ASSERT_EQ(res.c_se3_t.size(), views.size());
}

TEST(OptimizeIntrinsics, InsufficientViewsThrow) {  // validate_input, src/estimation/optim/intrinsics.cpp:92-96
    const auto cam = pinhole(1000, 1000, 640, 360);
    std::vector<PlanarView> views(3, PlanarView(12));
    std::vector<Eigen::Isometry3d> poses(3);
    EXPECT_THROW(optimize_intrinsics(views, cam, poses), std::invalid_argument);
}

// ---- tests/unit/bundle_test.cpp -------------------------------------------------------------------

static void recovers_x_and_intrinsics(double skew, bool optimize_skew, double skew_tol) {  // :9-81, :83-154
    RNG rng(7);
    const Eigen::Isometry3d g_se3_c_gt = make_pose(Eigen::Vector3d(0.03, 0.00, 0.12), Eigen::Vector3d(0, 1, 0), deg2rad(8.0));
    const Eigen::Isometry3d b_se3_t_gt = make_pose(Eigen::Vector3d(0.5, -0.1, 0.8), Eigen::Vector3d(1, 0, 0), deg2rad(14.0));
    const auto cam_gt = pinhole(1000, 1005, 640, 360, skew);

    SimulatedHandEye sim{g_se3_c_gt, b_se3_t_gt, cam_gt};
    sim.make_sequence(25, rng);
    sim.make_target_grid(8, 11, 0.02);
    sim.render_pixels();

    auto cam0 = pinhole(cam_gt.kmtx.fx * 0.97, cam_gt.kmtx.fy * 1.03, cam_gt.kmtx.cx + 5.0, cam_gt.kmtx.cy - 4.0, optimize_skew ? 0.0 : skew);
    Eigen::Isometry3d g_se3_c0 = g_se3_c_gt;
    g_se3_c0.translation() += Eigen::Vector3d(-0.01, 0.006, -0.004);
    g_se3_c0.linear() = axis_angle_to_R(Eigen::Vector3d(0.3, 0.7, -0.2).normalized(), deg2rad(2.0)) * g_se3_c0.linear();

    BundleOptions opts;
    opts.optimize_intrinsics = true;
    opts.optimize_skew = optimize_skew;
    opts.core.optimizer = OptimizerType::DENSE_QR;
    opts.core.huber_delta = -1;
    opts.core.verbose = false;

    auto result = optimize_bundle<Camera<BrownConradyd>>(sim.observations, {cam0}, {g_se3_c0}, b_se3_t_gt, opts);
    const auto& X = result.g_se3_c[0];
    const auto& Kf = result.cameras[0].kmtx;
    const auto& K_gt = cam_gt.kmtx;

    EXPECT_LT(rad2deg(rotation_angle(X.linear().transpose() * g_se3_c_gt.linear())), 1e-6);
    EXPECT_LT((X.translation() - g_se3_c_gt.translation()).norm(), 1e-6);
    EXPECT_NEAR(Kf.fx, K_gt.fx, 1e-6);
    EXPECT_NEAR(Kf.fy, K_gt.fy, 1e-6);
    EXPECT_NEAR(Kf.cx, K_gt.cx, 1e-6);
    EXPECT_NEAR(Kf.cy, K_gt.cy, 1e-6);
    EXPECT_NEAR(Kf.skew, K_gt.skew, skew_tol);
    EXPECT_LT(rad2deg(rotation_angle(result.b_se3_t.linear().transpose() * b_se3_t_gt.linear())), 1e-6);
    EXPECT_LT((result.b_se3_t.translation() - b_se3_t_gt.translation()).norm(), 1e-6);
}
TEST(OptimizeBundle, RecoversXAndIntrinsics_NoDistortion) { recovers_x_and_intrinsics(0.0, false, 1e-9); }
TEST(OptimizeBundle, RecoversXAndIntrinsics_NoDistortionSkew) { recovers_x_and_intrinsics(0.001, true, 1e-6); }

TEST(ReprojectionRefine, DistortionRecoveryOptional) {  // :156-210
    RNG rng(137);
    auto cam_gt = pinhole(900, 905, 640, 360);
    cam_gt.distortion.coeffs << -0.12, 0.02, 0.0005, -0.0007, 0.001;
    const Eigen::Isometry3d g_se3_c_gt = make_pose(Eigen::Vector3d(0.03, 0.00, 0.12), Eigen::Vector3d(0, 1, 0), deg2rad(8.0));
    const Eigen::Isometry3d b_se3_t_gt = make_pose(Eigen::Vector3d(0.5, -0.1, 80), Eigen::Vector3d(1, 0, 0), deg2rad(14.0));

    SimulatedHandEye sim{g_se3_c_gt, b_se3_t_gt, cam_gt};
    sim.make_sequence(22, rng);
    sim.make_target_grid(7, 10, 0.022);
    sim.render_pixels();

    auto cam0 = cam_gt;
    cam0.distortion.coeffs = Eigen::VectorXd::Zero(5);
    Eigen::Isometry3d X0 = g_se3_c_gt;
    X0.translation() += Eigen::Vector3d(0.01, 0.006, -0.003);
    X0.linear() = axis_angle_to_R(Eigen::Vector3d(0.1, 0.8, 0.1).normalized(), deg2rad(2.0)) * X0.linear();

    BundleOptions opts;
    opts.optimize_intrinsics = true;
    opts.optimize_hand_eye = true;
    opts.optimize_target_pose = true;
    opts.core.optimizer = OptimizerType::DENSE_QR;

    auto result = optimize_bundle<Camera<BrownConradyd>>(sim.observations, {cam0}, {X0}, b_se3_t_gt, opts);
    const auto& X = result.g_se3_c[0];
    const auto& dist = result.cameras[0].distortion.coeffs;
    EXPECT_LT(rad2deg(rotation_angle(X.linear().transpose() * g_se3_c_gt.linear())), 0.1);
    EXPECT_LT((X.translation() - g_se3_c_gt.translation()).norm(), 0.02);
    for (int i = 0; i < 5; ++i) EXPECT_NEAR(dist[i], cam_gt.distortion.coeffs[i], 1e-5);
}

// Large bundle inputs take the staged, threaded packing (b200::BundleStage): the descriptor it fills must equal the serial one
// (b200::Soa) array for array, and a view that differs from the board must be noticed.
TEST(OptimizeBundle, StagedPackingEqualsTheSerialPacking) {
    const int board = 88, n = 12000;   // 1.056 M observations: above b200::kStageThreshold
    std::vector<BundleObservation> obs(n);
    RNG rng(11);
    auto d = [&](RNG& r) { return r.uni(-1.0, 1.0); };
    for (int b = 0; b < n; ++b) {
        obs[b].camera_index = static_cast<size_t>(b % 3);
        obs[b].b_se3_g = make_pose(Eigen::Vector3d(d(rng), d(rng), d(rng)), Eigen::Vector3d(d(rng), d(rng), 1.0), d(rng));
        obs[b].view.resize(board);
        for (int i = 0; i < board; ++i) obs[b].view[i] = {Vec2((i % 11) * 0.02, (i / 11) * 0.02), Vec2(d(rng) * 600.0, d(rng) * 300.0)};
    }
    ASSERT_TRUE(obs.size() * board >= b200::kStageThreshold);
    b200::BundleStage st;
    ASSERT_TRUE(st.pack(obs));
    b200::Soa s;
    for (const auto& o : obs) {
        s.add(o.view, static_cast<int>(o.camera_index), -1);
        const Eigen::Matrix3d R = o.b_se3_g.linear(); const Eigen::Vector3d t = o.b_se3_g.translation();
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) s.bTg.push_back(R(i, j));
        for (int i = 0; i < 3; ++i) s.bTg.push_back(t(i));
    }
    cal_problem_desc a{}, c{};
    st.fill(a); s.fill(c);
    EXPECT_EQ(a.n_blocks, c.n_blocks); EXPECT_EQ(a.n_obs, c.n_obs);
    if (std::getenv("CALIB_B200_PER_OBSERVATION") == nullptr) {
        EXPECT_EQ(a.board_n, c.board_n);
        EXPECT_TRUE(std::equal(a.board_x, a.board_x + board, c.board_x) && std::equal(a.board_y, a.board_y + board, c.board_y));
    }
    EXPECT_TRUE(std::equal(a.img_u, a.img_u + a.n_obs, c.img_u) && std::equal(a.img_v, a.img_v + a.n_obs, c.img_v));
    EXPECT_TRUE(std::equal(a.block_offset, a.block_offset + n + 1, c.block_offset) && std::equal(a.block_cam, a.block_cam + n, c.block_cam));
    EXPECT_TRUE(std::equal(a.block_b_se3_g, a.block_b_se3_g + 12 * static_cast<size_t>(n), c.block_b_se3_g));
    obs[7777].view[40].object_xy.x() += 1e-3;   // one corner off the board: the staged form does not apply
    b200::BundleStage st2;
    EXPECT_FALSE(st2.pack(obs));
    obs[7777].view.resize(60);                   // a ragged view neither
    b200::BundleStage st3;
    EXPECT_FALSE(st3.pack(obs));
}

TEST(OptimizeBundle, InputValidation) {  // :212-227
    std::vector<BundleObservation> observations(2);
    Camera<BrownConradyd> cam(CameraMatrix{100.0, 100.0, 64.0, 48.0}, Eigen::VectorXd::Zero(5));
    const Eigen::Isometry3d X0 = Eigen::Isometry3d::Identity();
    BundleOptions opts;
    EXPECT_THROW({ optimize_bundle<Camera<BrownConradyd>>(observations, {cam, cam}, {X0}, Eigen::Isometry3d::Identity(), opts); },
                 std::invalid_argument);
    // validate_input, src/estimation/optim/bundle.cpp:136-145
    EXPECT_THROW({ optimize_bundle<Camera<BrownConradyd>>(observations, {}, {}, X0, opts); }, std::invalid_argument);
    EXPECT_THROW({ optimize_bundle<Camera<BrownConradyd>>({}, {cam}, {X0}, X0, opts); }, std::invalid_argument);
    // every observation empty: "No observations provided" comes back from cal_refine_create as CAL_ERR_INVALID_ARGUMENT
    EXPECT_THROW({ optimize_bundle<Camera<BrownConradyd>>(observations, {cam}, {X0}, X0, opts); }, std::invalid_argument);
}

namespace {
struct SmallRig {  // the scene shared by :229-349 and scheimpflug_bundle_test.cpp
    Eigen::Isometry3d g_se3_c = Eigen::Isometry3d::Identity(), b_se3_t = Eigen::Isometry3d::Identity();
    SmallRig() {
        g_se3_c.linear() = Eigen::AngleAxisd(0.05, Eigen::Vector3d::UnitY()).toRotationMatrix();
        g_se3_c.translation() = Eigen::Vector3d(0.1, 0.0, 0.05);
        b_se3_t.translation() = Eigen::Vector3d(0.2, 0.0, 0.0);
    }
};
}  // namespace

TEST(OptimizeBundle, SingleCameraHandEye) {  // :229-262
    Camera<BrownConradyd> cam(CameraMatrix{100.0, 100.0, 64.0, 48.0}, Eigen::VectorXd::Zero(5));
    const SmallRig rig;
    std::vector<Vec2> obj{{-0.1, -0.1}, {0.1, -0.1}, {0.1, 0.1}, {-0.1, 0.1}, {0.5, 0.5}, {-1.0, -1.0}, {2.0, 2.0}, {2.5, 0.5}, {9, 0}};
    std::vector<Camera<BrownConradyd>> cams{cam};
    auto observations = make_bundle_observations(cams, {rig.g_se3_c}, rig.b_se3_t, obj, make_circle_poses(8, 0.1, 0.3, 0.05, 0.1, 0.5));
    Eigen::Isometry3d init_g_se3_c = rig.g_se3_c;
    init_g_se3_c.translation() += Eigen::Vector3d(0.01, -0.01, 0.02);

    BundleOptions opts;
    opts.optimize_intrinsics = false;
    opts.optimize_target_pose = false;
    opts.optimize_hand_eye = true;
    auto res = optimize_bundle<Camera<BrownConradyd>>(observations, cams, {init_g_se3_c}, rig.b_se3_t, opts);

    EXPECT_LT((res.g_se3_c[0].translation() - rig.g_se3_c.translation()).norm(), 1e-3);
    Eigen::AngleAxisd diff(res.g_se3_c[0].linear() * rig.g_se3_c.linear().transpose());
    EXPECT_LT(diff.angle(), 1e-3);
    EXPECT_LT(res.core.final_cost, 0.01);
}

TEST(OptimizeBundle, SingleCameraTargetPose) {  // :264-293
    Camera<BrownConradyd> cam(CameraMatrix{100.0, 100.0, 64.0, 48.0}, Eigen::VectorXd::Zero(5));
    const SmallRig rig;
    std::vector<Vec2> obj{{-0.1, -0.1}, {0.1, -0.1}, {0.1, 0.1}, {-0.1, 0.1}, {0.5, 0.5}, {-1.0, -1.0}, {2.0, 2.0}, {2.5, 0.5}};
    std::vector<Camera<BrownConradyd>> cams{cam};
    auto observations = make_bundle_observations(cams, {rig.g_se3_c}, rig.b_se3_t, obj, make_circle_poses(8, 0.1, 0.3, 0.05, 0.1, 0.5));
    Eigen::Isometry3d init_b_se3_t = rig.b_se3_t;
    init_b_se3_t.translation() += Eigen::Vector3d(0.01, -0.02, 0.03);

    BundleOptions opts;
    opts.optimize_intrinsics = false;
    opts.optimize_target_pose = true;
    opts.optimize_hand_eye = false;
    auto res = optimize_bundle<Camera<BrownConradyd>>(observations, cams, {rig.g_se3_c}, init_b_se3_t, opts);

    EXPECT_LT((res.b_se3_t.translation() - rig.b_se3_t.translation()).norm(), 1e-3);
    Eigen::AngleAxisd diff(res.b_se3_t.linear() * rig.b_se3_t.linear().transpose());
    EXPECT_LT(diff.angle(), 1e-3);
}

TEST(OptimizeBundle, TwoCamerasHandEyeExtrinsics) {  // :295-349
    Camera<BrownConradyd> cam0(CameraMatrix{100.0, 100.0, 64.0, 48.0}, Eigen::VectorXd::Zero(5));
    Camera<BrownConradyd> cam1 = cam0;
    const SmallRig rig;
    const Eigen::Isometry3d g_se3_c0 = rig.g_se3_c;
    Eigen::Isometry3d c1_se3_c0 = Eigen::Isometry3d::Identity();
    c1_se3_c0.translation() = Eigen::Vector3d(0.05, 0.0, 0.0);
    c1_se3_c0.linear() = Eigen::AngleAxisd(0.1, Eigen::Vector3d::UnitZ()).toRotationMatrix();
    const Eigen::Isometry3d g_se3_c1 = g_se3_c0 * c1_se3_c0.inverse();

    std::vector<Vec2> obj{{-0.1, -0.1}, {0.1, -0.1}, {0.1, 0.1}, {-0.1, 0.1}, {0.5, 0.5}, {-1.0, -1.0}, {2.0, 2.0}, {2.5, 0.5}};
    std::vector<Camera<BrownConradyd>> cams{cam0, cam1};
    auto observations = make_bundle_observations(cams, {g_se3_c0, g_se3_c1}, rig.b_se3_t, obj, make_circle_poses(8, 0.1, 0.3, 0.05, 0.1, 0.5));
    Eigen::Isometry3d init_g_se3_c1 = g_se3_c1;
    init_g_se3_c1.translation() += Eigen::Vector3d(0.01, -0.01, 0.0);
    init_g_se3_c1.linear() = g_se3_c1.linear() * Eigen::AngleAxisd(0.01, Eigen::Vector3d::UnitZ()).toRotationMatrix();
    Eigen::Isometry3d init_g_se3_c0 = g_se3_c0;
    init_g_se3_c0.translation() += Eigen::Vector3d(-0.01, 0.02, -0.02);

    BundleOptions opts;
    opts.optimize_intrinsics = false;
    opts.optimize_target_pose = false;
    opts.optimize_hand_eye = true;
    auto res = optimize_bundle<Camera<BrownConradyd>>(observations, cams, {init_g_se3_c0, init_g_se3_c1}, rig.b_se3_t, opts);

    EXPECT_LT((res.g_se3_c[0].translation() - g_se3_c0.translation()).norm(), 1e-3);
    EXPECT_LT(Eigen::AngleAxisd(res.g_se3_c[0].linear() * g_se3_c0.linear().transpose()).angle(), 1e-3);
    EXPECT_LT((res.g_se3_c[1].translation() - g_se3_c1.translation()).norm(), 1e-3);
    EXPECT_LT(Eigen::AngleAxisd(res.g_se3_c[1].linear() * g_se3_c1.linear().transpose()).angle(), 1e-3);
    // covariance in BundleBlocks::get_param_blocks order (bundle.cpp:48-68): 2 x 10 + 2 x 4 + 2 x 3 + 4 + 3
    EXPECT_EQ(res.core.covariance.rows(), static_cast<Eigen::Index>(41));
}

// ---- tests/unit/scheimpflug_bundle_test.cpp -------------------------------------------------------

TEST(ScheimpflugBundle, IntrinsicsWithFixedHandeye) {  // :13-56
    PinholeCamera<BrownConradyd> cam(CameraMatrix{100.0, 100.0, 64.0, 48.0}, Eigen::VectorXd::Zero(5));
    const double taux = 0.02, tauy = -0.015;
    ScheimpflugCamera<PinholeCamera<BrownConradyd>> sc(cam, {taux, tauy});
    const SmallRig rig;
    std::vector<Vec2> obj{{-0.1, -0.1}, {0.1, -0.1}, {0.1, 0.1}, {-0.1, 0.1}, {0.05, 0.0}, {-0.05, 0.0}, {0.0, 0.05}, {0.0, -0.05}};
    auto obs = make_scheimpflug_observations<BrownConradyd>({sc}, {rig.g_se3_c}, rig.b_se3_t, obj, make_circle_poses(8, 0.1, 0.3, 0.05, 0.1, 0.5));
    sc.tau_x += 0.01;
    sc.tau_y -= 0.01;

    BundleOptions opts;
    opts.optimize_intrinsics = true;
    opts.optimize_target_pose = false;
    opts.optimize_hand_eye = false;
    opts.core.optimizer = OptimizerType::DENSE_QR;
    auto res = optimize_bundle(obs, std::vector<ScheimpflugCamera<PinholeCamera<BrownConradyd>>>{sc}, {rig.g_se3_c}, rig.b_se3_t, opts);

    EXPECT_LT((res.g_se3_c[0].translation() - rig.g_se3_c.translation()).norm(), 1e-6);
    EXPECT_LT(Eigen::AngleAxisd(res.g_se3_c[0].linear() * rig.g_se3_c.linear().transpose()).angle(), 1e-6);
    EXPECT_NEAR(res.cameras[0].tau_x, taux, 1e-6);
    EXPECT_NEAR(res.cameras[0].tau_y, tauy, 1e-6);
}

TEST(ScheimpflugBundle, HandeyeWithFixedIntrinsics) {  // :58-94
    PinholeCamera<BrownConradyd> cam(CameraMatrix{100.0, 100.0, 64.0, 48.0}, Eigen::VectorXd::Zero(5));
    const double taux = 0.02, tauy = -0.015;
    ScheimpflugCamera<PinholeCamera<BrownConradyd>> sc(cam, {taux, tauy});
    const SmallRig rig;
    std::vector<Vec2> obj{{-0.1, -0.1}, {0.1, -0.1}, {0.1, 0.1}, {-0.1, 0.1}, {0.05, 0.0}, {-0.05, 0.0}, {0.0, 0.05}, {0.0, -0.05}};
    auto observations =
        make_scheimpflug_observations<BrownConradyd>({sc}, {rig.g_se3_c}, rig.b_se3_t, obj, make_circle_poses(8, 0.1, 0.3, 0.05, 0.1, 0.5));
    Eigen::Isometry3d init_g_se3_c = rig.g_se3_c;
    init_g_se3_c.translation() += Eigen::Vector3d(0.01, -0.01, 0.02);

    BundleOptions opts;
    opts.optimize_intrinsics = false;
    opts.optimize_target_pose = false;
    opts.optimize_hand_eye = true;
    auto res = optimize_bundle(observations, std::vector<ScheimpflugCamera<PinholeCamera<BrownConradyd>>>{sc}, {init_g_se3_c}, rig.b_se3_t, opts);

    EXPECT_LT((res.g_se3_c[0].translation() - rig.g_se3_c.translation()).norm(), 1e-6);
    EXPECT_LT(Eigen::AngleAxisd(res.g_se3_c[0].linear() * rig.g_se3_c.linear().transpose()).angle(), 1e-6);
    EXPECT_NEAR(res.cameras[0].tau_x, taux, 1e-6);
    EXPECT_NEAR(res.cameras[0].tau_y, tauy, 1e-6);
}

// ---- tests/unit/scheimpflug_test.cpp: the model that renders the scenes above (calib_b200_mini.hpp) ----

TEST(ScheimpflugCamera, ZeroTiltMatchesPinhole) {  // :11-29
    const auto cam = pinhole(800, 820, 320, 240);
    ScheimpflugCamera<PinholeCamera<BrownConradyd>> sc(cam, {0.0, 0.0});
    const Eigen::Vector3d Xc(0.2, -0.1, 1.0);
    EXPECT_NEAR(sc.project(Xc).x(), cam.project(Xc).x(), 1e-9);
    EXPECT_NEAR(sc.project(Xc).y(), cam.project(Xc).y(), 1e-9);
}

TEST(ScheimpflugCamera, PrincipalRay) {  // :31-51
    const auto cam = pinhole(600, 600, 400, 300);
    const double taux = 0.1, tauy = -0.2;
    ScheimpflugCamera<PinholeCamera<BrownConradyd>> sc(cam, {taux, tauy});
    const Eigen::Vector2d uv = sc.project(Eigen::Vector3d(0.0, 0.0, 1.0));
    const Eigen::Vector2d expected_uv = cam.project(Eigen::Vector2d(-std::tan(tauy) / std::cos(taux), std::tan(taux)));
    EXPECT_NEAR(uv.x(), expected_uv.x(), 1e-9);
    EXPECT_NEAR(uv.y(), expected_uv.y(), 1e-9);
}

// ---- tests/unit/extrinsics_test.cpp ---------------------------------------------------------------

namespace {
struct StereoScene {
    std::vector<Camera<BrownConradyd>> cameras_gt;
    std::vector<Eigen::Isometry3d> cam_gt, target_gt;
    std::vector<MulticamPlanarView> views;
    StereoScene(std::vector<Eigen::Isometry3d> targets, const std::vector<Vec2>& points) : target_gt(std::move(targets)) {
        Eigen::VectorXd dist(5);
        dist << 0.0, 0.0, 0.0, 0.0, 0.0;
        const CameraMatrix kmtx{100.0, 100.0, 0.0, 0.0};
        cameras_gt = {Camera<BrownConradyd>{kmtx, dist}, Camera<BrownConradyd>{kmtx, dist}};
        cam_gt = {Eigen::Isometry3d::Identity(), Eigen::Translation3d(1.0, 0.0, 0.0) * Eigen::Isometry3d::Identity()};
        for (size_t v = 0; v < target_gt.size(); ++v) {
            MulticamPlanarView view(2);
            for (int c = 0; c < 2; ++c) {
                const Eigen::Isometry3d T = cam_gt[c] * target_gt[v];
                for (const auto& xy : points) {
                    const Eigen::Vector3d P = T * Eigen::Vector3d(xy.x(), xy.y(), 0.0);
                    view[c].push_back({xy, denormalize(cameras_gt[c].kmtx, Vec2(P.x() / P.z(), P.y() / P.z()))});
                }
            }
            views.push_back(std::move(view));
        }
    }
};
const std::vector<Vec2> k_eight_points = {{0.0, 0.0}, {1.0, 0.0}, {1.0, 1.0}, {0.0, 1.0}, {0.5, 0.5}, {-1.0, -1.0}, {2.0, 2.0}, {2.5, 0.5}};
std::vector<Camera<BrownConradyd>> perturbed_intrinsics() {
    return {Camera<BrownConradyd>{CameraMatrix{90.0, 95.0, 1.0, -1.0}, Eigen::VectorXd::Zero(5)},
            Camera<BrownConradyd>{CameraMatrix{105.0, 98.0, -0.5, 0.5}, Eigen::VectorXd::Zero(5)}};
}
}  // namespace

TEST(Extrinsics, RecoverCameraAndTargetPoses) {  // :9-73
    StereoScene s({Eigen::Translation3d(0.0, 0.0, 5.0) * Eigen::Isometry3d::Identity(),
                   Eigen::Translation3d(0.5, -0.2, 4.0) * Eigen::AngleAxisd(0.3, Eigen::Vector3d::UnitY()),
                   Eigen::Translation3d(-0.3, 0.4, 6.0) * Eigen::AngleAxisd(-0.2, Eigen::Vector3d::UnitX())},
                  {{0.0, 0.0}, {1.0, 0.0}, {1.0, 1.0}, {0.0, 1.0}});
    std::vector<Eigen::Isometry3d> cam_init = {s.cam_gt[0], Eigen::Translation3d(1.2, -0.1, 0.05) * Eigen::AngleAxisd(0.05, Eigen::Vector3d::UnitZ())};
    std::vector<Eigen::Isometry3d> target_init = {
        s.target_gt[0] * Eigen::Translation3d(0.1, 0.0, 0.0) * Eigen::AngleAxisd(0.02, Eigen::Vector3d::UnitZ()),
        s.target_gt[1] * Eigen::Translation3d(-0.05, 0.1, 0.05) * Eigen::AngleAxisd(-0.03, Eigen::Vector3d::UnitY()),
        s.target_gt[2] * Eigen::Translation3d(0.02, -0.02, -0.1) * Eigen::AngleAxisd(0.01, Eigen::Vector3d::UnitX())};

    ExtrinsicOptions opts;
    opts.optimize_intrinsics = false;
    auto result = optimize_extrinsics(s.views, s.cameras_gt, cam_init, target_init, opts);

    EXPECT_LT(result.core.final_cost, 1e-6);
    ASSERT_EQ(result.c_se3_r.size(), static_cast<size_t>(2));
    ASSERT_EQ(result.r_se3_t.size(), s.target_gt.size());
    EXPECT_TRUE(result.c_se3_r[1].translation().isApprox(s.cam_gt[1].translation(), 1e-3));
    EXPECT_TRUE(result.c_se3_r[1].linear().isApprox(s.cam_gt[1].linear(), 1e-3));
    for (size_t v = 0; v < s.target_gt.size(); ++v) {  // with optimize_intrinsics = false no target pose is held (extrinsics.cpp:118-127)
        EXPECT_TRUE(result.r_se3_t[v].translation().isApprox(s.target_gt[v].translation(), 1e-3));
        EXPECT_TRUE(result.r_se3_t[v].linear().isApprox(s.target_gt[v].linear(), 1e-3));
    }
    EXPECT_TRUE(result.c_se3_r[0].isApprox(Eigen::Isometry3d::Identity(), 1e-12));  // camera 0 is the gauge (extrinsics.cpp:136-139)
}

// RecoverAllParameters / FirstTargetPoseFixed start from estimate_extrinsic_dlt (a linear seed outside the path);
// the seed here is the ground truth moved by a fixed small motion, which the assertions do not depend on.
TEST(Extrinsics, RecoverAllParameters) {  // :75-140
    StereoScene s({Eigen::Translation3d(0.0, 0.0, 5.0) * Eigen::Isometry3d::Identity(),
                   Eigen::Translation3d(0.5, -0.2, 4.0) * Eigen::AngleAxisd(0.3, Eigen::Vector3d::UnitY())},
                  k_eight_points);
    std::vector<Eigen::Isometry3d> c_se3_r = {Eigen::Isometry3d::Identity(),
                                              s.cam_gt[1] * Eigen::Translation3d(0.02, -0.01, 0.01) * Eigen::AngleAxisd(0.01, Eigen::Vector3d::UnitZ())};
    std::vector<Eigen::Isometry3d> r_se3_t = {s.target_gt[0],  // "Anchor the first target pose to its ground truth to fix the scale."
                                              s.target_gt[1] * Eigen::Translation3d(0.01, 0.02, -0.02) * Eigen::AngleAxisd(0.01, Eigen::Vector3d::UnitY())};
    ExtrinsicOptions opts;
    auto res = optimize_extrinsics(s.views, perturbed_intrinsics(), c_se3_r, r_se3_t, opts);

    EXPECT_LT(res.core.final_cost, 1e-6);
    ASSERT_EQ(res.cameras.size(), static_cast<size_t>(2));
    EXPECT_NEAR(res.cameras[0].kmtx.fx, 100.0, 1e-3);
    EXPECT_NEAR(res.cameras[0].kmtx.fy, 100.0, 1e-3);
    EXPECT_TRUE(res.c_se3_r[1].translation().isApprox(s.cam_gt[1].translation(), 1e-3));
    EXPECT_TRUE(res.r_se3_t[0].translation().isApprox(s.target_gt[0].translation(), 1e-3));
    EXPECT_GT(res.core.covariance.trace(), 0.0);
    // ExtrinsicBlocks::get_param_blocks order (extrinsics.cpp:50-69): 2 x 10 + 2 x 4 + 2 x 3 + 2 x 4 + 2 x 3
    EXPECT_EQ(res.core.covariance.rows(), static_cast<Eigen::Index>(48));
}

TEST(Extrinsics, FirstTargetPoseFixed) {  // :142-199
    StereoScene s({Eigen::Translation3d(0.0, 0.0, 5.0) * Eigen::Isometry3d::Identity(),
                   Eigen::Translation3d(0.5, -0.2, 4.0) * Eigen::AngleAxisd(0.3, Eigen::Vector3d::UnitY())},
                  k_eight_points);
    std::vector<Eigen::Isometry3d> c_se3_r = {Eigen::Isometry3d::Identity(),
                                              s.cam_gt[1] * Eigen::Translation3d(0.02, -0.01, 0.01) * Eigen::AngleAxisd(0.01, Eigen::Vector3d::UnitZ())};
    std::vector<Eigen::Isometry3d> r_se3_t = {s.target_gt[0], s.target_gt[1]};
    r_se3_t[0].translation() = Eigen::Vector3d(0.0, 0.0, 3.0);  // deliberately wrong scale; must stay unchanged
    ExtrinsicOptions opts;
    auto res = optimize_extrinsics(s.views, perturbed_intrinsics(), c_se3_r, r_se3_t, opts);
    EXPECT_TRUE(res.r_se3_t[0].translation().isApprox(r_se3_t[0].translation(), 1e-12));
    EXPECT_GT(res.core.final_cost, 0.1);
}

TEST(Extrinsics, MismatchedPoseVectorsThrow) {  // src/estimation/optim/extrinsics.cpp:163-171
    StereoScene s({Eigen::Translation3d(0.0, 0.0, 5.0) * Eigen::Isometry3d::Identity()}, k_eight_points);
    EXPECT_THROW(optimize_extrinsics(s.views, s.cameras_gt, {Eigen::Isometry3d::Identity()}, s.target_gt), std::invalid_argument);
    EXPECT_THROW(optimize_extrinsics(s.views, s.cameras_gt, s.cam_gt, {}), std::invalid_argument);
    auto short_views = s.views;
    short_views[0].pop_back();  // a view that lists one camera only
    EXPECT_THROW(optimize_extrinsics(short_views, s.cameras_gt, s.cam_gt, s.target_gt), std::invalid_argument);
}

// ---- tests/unit/handeye_test.cpp ------------------------------------------------------------------

TEST(CeresAXXBRefine, ImprovesOverInitializer) {  // :101-152
    RNG rng(2024);
    // make_pose(t, rng.rand_unit_axis(), angle): one draw per pose, in statement order
    const Eigen::Vector3d ax1 = rng.rand_unit_axis();
    const Eigen::Isometry3d X_gt = make_pose(Eigen::Vector3d(0.02, -0.01, 0.09), ax1, deg2rad(10.0));
    const Eigen::Vector3d ax2 = rng.rand_unit_axis();
    const Eigen::Isometry3d b_se3_t_gt = make_pose(Eigen::Vector3d(0.25, 0.05, 0.55), ax2, deg2rad(18.0));
    const auto cam_gt = pinhole(950, 960, 640, 360);

    SimulatedHandEye sim{X_gt, b_se3_t_gt, cam_gt};
    sim.make_sequence(18, rng);
    sim.make_target_grid(6, 9, 0.025);
    sim.render_pixels(0.0, nullptr);
    const auto base_se3_gripper = sim.b_se3_g();
    const auto& camera_se3_target = sim.c_se3_t;

    Eigen::Isometry3d X0 = X_gt;
    {
        const Eigen::Vector3d ax = rng.rand_unit_axis();
        X0.linear() = axis_angle_to_R(ax, deg2rad(2.0)) * X0.linear();
        X0.translation() += Eigen::Vector3d(0.01, -0.005, 0.004);
    }
    const double err0_rot = rad2deg(rotation_angle(X0.linear().transpose() * X_gt.linear()));
    const double err0_tr = (X0.translation() - X_gt.translation()).norm();

    OptimOptions ro;
    ro.optimizer = OptimizerType::DENSE_QR;
    ro.max_iterations = 60;
    ro.huber_delta = 1.0;
    auto res = optimize_handeye(base_se3_gripper, camera_se3_target, X0, ro);
    const Eigen::Isometry3d Xr = res.g_se3_c;

    const double err1_rot = rad2deg(rotation_angle(Xr.linear().transpose() * X_gt.linear()));
    const double err1_tr = (Xr.translation() - X_gt.translation()).norm();
    EXPECT_LT(err1_rot, err0_rot);
    EXPECT_LT(err1_tr, err0_tr);
    EXPECT_LT(err1_rot, 0.05);
    EXPECT_LT(err1_tr, 0.002);
    EXPECT_EQ(res.core.covariance.rows(), static_cast<Eigen::Index>(7));
}

TEST(CeresAXXBRefine, ThrowsOnDegenerateSmallMotions) {  // build_all_pairs, src/estimation/linear/handeyedlt.cpp:56-58,76-79 (cf. :54-60 of the test file)
    std::vector<Eigen::Isometry3d> b_se3_g(5, Eigen::Isometry3d::Identity()), c_se3_t(5, Eigen::Isometry3d::Identity());
    EXPECT_THROW(optimize_handeye(b_se3_g, c_se3_t, Eigen::Isometry3d::Identity()), std::runtime_error);
    c_se3_t.pop_back();
    EXPECT_THROW(optimize_handeye(b_se3_g, c_se3_t, Eigen::Isometry3d::Identity()), std::runtime_error);
}

// ---- tests/unit/homography_test.cpp (the estimate_homography halves; optimize_homography is outside the path) ----

namespace {
Vec2 apply_homography(const Mat3& H, const Vec2& p) { return (H * p.homogeneous()).hnormalized(); }

void generate_synthetic_data(PlanarView& view, Mat3& true_H, int n_points = 50, double noise_level = -1) {  // :21-47
    const double angle = 0.1, c = std::cos(angle), s = std::sin(angle);
    true_H << c, -s, 10.0, s, c, -5.0, 0.001, -0.002, 1.0;
    std::mt19937 rng(42);
    std::uniform_real_distribution<double> dist(-100.0, 100.0);
    std::normal_distribution<double> noise(0.0, noise_level > 0 ? noise_level : 1.0);
    view.resize(static_cast<size_t>(n_points));
    for (auto& ob : view) {
        // `Vec2 point(dist(rng), dist(rng))`: GCC evaluates the arguments right to left — the first draw is y
        const double py = dist(rng), px = dist(rng);
        const Vec2 point(px, py);
        Vec2 pixel = apply_homography(true_H, point);
        if (noise_level > 0) {
            const double ny = noise(rng), nx = noise(rng);
            pixel += Vec2(nx, ny);
        }
        ob = {point, pixel};
    }
}
void add_outliers(PlanarView& view, uint32_t seed, int n) {  // :111-118
    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> dist(-100.0, 100.0);
    for (int i = 0; i < n; ++i) {
        const double sy = dist(rng), sx = dist(rng);
        const double dy = dist(rng), dx = dist(rng);
        view.push_back({Vec2(sx, sy), Vec2(dx, dy)});
    }
}
}  // namespace

TEST(HomographyTest, ExactHomography) {  // :50-72
    Mat3 H_true = Mat3::Identity();
    H_true(0, 2) = 10.0;
    H_true(1, 2) = -5.0;
    std::vector<Vec2> src = {{0.0, 0.0}, {1.0, 0.0}, {0.0, 1.0}, {1.0, 1.0}};
    PlanarView view;
    for (const auto& p : src) view.push_back({p, apply_homography(H_true, p)});
    const auto hres = estimate_homography(view);
    ASSERT_TRUE(hres.success);
    ASSERT_TRUE(hres.hmtx.isApprox(H_true, 1e-6));
    EXPECT_EQ(hres.inliers.size(), view.size());
}

TEST(HomographyTest, NoisyHomography) {  // :74-94
    PlanarView view;
    Mat3 H_true;
    generate_synthetic_data(view, H_true, 50, 0.1);
    const auto hres = estimate_homography(view);
    ASSERT_TRUE(hres.success);
    EXPECT_LT(hres.symmetric_rms_px, 0.25);
    ASSERT_TRUE(hres.hmtx.isApprox(H_true, 1e-2));
}

TEST(HomographyTest, InsufficientPoints) {  // HomographyEstimator::fit -> nullopt below four points
    PlanarView view{{{0.0, 0.0}, {10.0, 0.0}}, {{1.0, 0.0}, {11.0, 0.0}}, {{0.0, 1.0}, {10.0, 1.0}}};
    EXPECT_FALSE(estimate_homography(view).success);
    EXPECT_FALSE(estimate_homography(view, RansacOptions{}).success);
}

TEST(HomographyTest, RansacRecoversHomographyWithOutliers) {  // :104-134
    PlanarView view;
    Mat3 H_true;
    generate_synthetic_data(view, H_true, 100, 0.0);
    add_outliers(view, 7, 30);
    RansacOptions opts;
    opts.thresh = 1.0;
    opts.min_inliers = 90;
    opts.seed = 123;
    const auto hres = estimate_homography(view, opts);
    ASSERT_TRUE(hres.success);
    EXPECT_GE(hres.inliers.size(), static_cast<size_t>(95));
    EXPECT_LT(hres.symmetric_rms_px, 1e-3);
    EXPECT_TRUE(hres.hmtx.isApprox(H_true, 1e-2));
    for (int idx : hres.inliers) EXPECT_LT(idx, 100);  // the scene's outliers are far from the model
}

TEST(HomographyTest, RansacFailsWithTooFewInliers) {  // :137-160
    PlanarView view;
    Mat3 H_true;
    generate_synthetic_data(view, H_true, 4, 0.0);
    add_outliers(view, 3, 50);
    RansacOptions opts;
    opts.thresh = 0.5;
    opts.min_inliers = 10;
    opts.seed = 42;
    const auto hres = estimate_homography(view, opts);
    EXPECT_FALSE(hres.success);
}

// ---- tests/unit/planefit_test.cpp -----------------------------------------------------------------

TEST(PlaneFit, RansacRejectsOutliers) {  // :24-74
    std::mt19937 rng(1337);
    std::uniform_real_distribution<double> dist_xy(-1.0, 1.0);
    const Eigen::Vector3d n = Eigen::Vector3d(0.2, -0.3, 1.0).normalized();
    const Eigen::Vector4d ground_truth(n.x(), n.y(), n.z(), -n.dot(Eigen::Vector3d(0.0, 0.0, 1.0)));
    std::vector<Eigen::Vector3d> points;
    const int inliers = 100;
    for (int i = 0; i < inliers; ++i) {
        const double x = dist_xy(rng);
        const double y = dist_xy(rng);
        points.emplace_back(x, y, (-ground_truth[3] - ground_truth[0] * x - ground_truth[1] * y) / ground_truth[2]);
    }
    std::uniform_real_distribution<double> dist_out(5.0, 10.0);
    for (int i = 0; i < 40; ++i) {  // emplace_back(dist_out(rng), dist_out(rng), dist_out(rng)): right to left
        const double z = dist_out(rng), y = dist_out(rng), x = dist_out(rng);
        points.emplace_back(x, y, z);
    }
    RansacOptions opts;
    opts.max_iters = 2000;
    opts.thresh = 0.01;
    opts.min_inliers = 80;
    opts.confidence = 0.999;
    auto result = fit_plane_ransac(points, opts);
    ASSERT_TRUE(result.success);
    EXPECT_GE(result.inliers.size(), static_cast<size_t>(inliers));
    const Eigen::Vector4d estimated = (result.plane.head<3>().dot(ground_truth.head<3>()) < 0.0) ? -result.plane : result.plane;
    for (int i = 0; i < 4; ++i) EXPECT_NEAR(estimated[i], ground_truth[i], 1e-3);
    EXPECT_LT(result.inlier_rms, 1e-3);
    size_t counted = 0;
    for (int idx : result.inliers) {
        const auto& p = points[static_cast<size_t>(idx)];
        if (std::abs(result.plane.head<3>().dot(p) + result.plane[3]) < opts.thresh) ++counted;
    }
    EXPECT_EQ(counted, result.inliers.size());
}

TEST(PlaneFit, TooFewPointsFail) {  // src/estimation/linear/planefit.cpp:88-90
    EXPECT_FALSE(fit_plane_ransac({Eigen::Vector3d(0, 0, 0), Eigen::Vector3d(1, 0, 0)}).success);
}

// ---- tests/unit/intrinsics_estimate_test.cpp, planarpose_test.cpp ---------------------------------

TEST(EstimateIntrinsics, RecoversCameraMatrix) {  // :11-55
    RNG rng(10);
    const auto cam_gt = pinhole(900, 920, 640, 360);
    SimulatedHandEye sim{Eigen::Isometry3d::Identity(), Eigen::Translation3d(0.0, 0.0, 2.0) * Eigen::Isometry3d::Identity(), cam_gt};
    sim.make_sequence(8, rng);
    sim.make_target_grid(6, 9, 0.03);
    sim.render_pixels();
    const auto views = views_of(sim);

    auto res = estimate_intrinsics(views);
    ASSERT_TRUE(res.success);
    EXPECT_NEAR(res.kmtx.fx, cam_gt.kmtx.fx, 1e-6);
    EXPECT_NEAR(res.kmtx.fy, cam_gt.kmtx.fy, 1e-6);
    EXPECT_NEAR(res.kmtx.cx, cam_gt.kmtx.cx, 1e-6);
    EXPECT_NEAR(res.kmtx.cy, cam_gt.kmtx.cy, 1e-6);
    EXPECT_NEAR(res.kmtx.skew, cam_gt.kmtx.skew, 1e-9);
    ASSERT_EQ(res.views.size(), views.size());
    for (size_t i = 0; i < res.views.size(); ++i) {
        const auto& est = res.views[i].c_se3_t;
        const auto& gt = sim.c_se3_t[i];
        EXPECT_TRUE(gt.linear().isApprox(est.linear(), 1e-6) || gt.linear().isApprox(-est.linear(), 1e-6));
        EXPECT_GT(std::abs(gt.translation().normalized().dot(est.translation().normalized())), 0.999);
        EXPECT_EQ(res.views[i].view_index, i);
    }
}

TEST(EstimateIntrinsics, FailsWithTooFewViews) {  // :57-82
    RNG rng(5);
    const auto cam_gt = pinhole(800, 805, 320, 240);
    SimulatedHandEye sim{Eigen::Isometry3d::Identity(), Eigen::Translation3d(0.0, 0.0, 2.0) * Eigen::Isometry3d::Identity(), cam_gt};
    sim.make_sequence(3, rng);
    sim.make_target_grid(5, 7, 0.04);
    sim.render_pixels();
    EXPECT_FALSE(estimate_intrinsics(views_of(sim)).success);
    EXPECT_FALSE(estimate_intrinsics({}).success);
}

TEST(EstimateIntrinsics, HomographyRansacOption) {  // IntrinsicsEstimOptions::homography_ransac, src/estimation/linear/intrinsicsdlt.cpp:50-64
    RNG rng(10);
    const auto cam_gt = pinhole(900, 920, 640, 360);
    SimulatedHandEye sim{Eigen::Isometry3d::Identity(), Eigen::Translation3d(0.0, 0.0, 2.0) * Eigen::Isometry3d::Identity(), cam_gt};
    sim.make_sequence(8, rng);
    sim.make_target_grid(6, 9, 0.03);
    sim.render_pixels();
    auto views = views_of(sim);
    for (auto& v : views) { v[5].image_uv += Vec2(40.0, -25.0); v[17].image_uv += Vec2(-30.0, 35.0); }  // two gross outliers per view
    IntrinsicsEstimOptions opts;
    opts.homography_ransac = RansacOptions{};
    auto res = estimate_intrinsics(views, opts);
    ASSERT_TRUE(res.success);
    EXPECT_NEAR(res.kmtx.fx, cam_gt.kmtx.fx, 1e-5);
    EXPECT_NEAR(res.kmtx.fy, cam_gt.kmtx.fy, 1e-5);
    ASSERT_EQ(res.views.size(), views.size());
    for (const auto& v : res.views) {
        EXPECT_EQ(v.homography.inliers.size(), static_cast<size_t>(52));
        for (int idx : v.homography.inliers) EXPECT_TRUE(idx != 5 && idx != 17);
    }
}

TEST(PlanarPoseTest, DLTEstimation) {  // planarpose_test.cpp:60-94 with the data of :15-35
    const CameraMatrix intrinsics{1000, 1000, 500, 500};
    Eigen::Isometry3d true_pose = Eigen::Isometry3d::Identity();
    true_pose.linear() = Eigen::AngleAxisd(0.1, Eigen::Vector3d(1, 1, 1).normalized()).toRotationMatrix();
    true_pose.translation() = Eigen::Vector3d(0.1, 0.2, 2.0);
    PlanarView view;
    for (int i = -5; i <= 5; i += 2)
        for (int j = -5; j <= 5; j += 2) {
            const Vec2 obj_pt(i * 0.1, j * 0.1);
            view.push_back({obj_pt, denormalize(intrinsics, (true_pose * Eigen::Vector3d(obj_pt.x(), obj_pt.y(), 0.0)).hnormalized())});
        }
    const Eigen::Isometry3d estimated_pose = estimate_planar_pose(view, intrinsics);
    EXPECT_TRUE(true_pose.linear().isApprox(estimated_pose.linear(), 1e-1) || true_pose.linear().isApprox(-estimated_pose.linear(), 1e-1));
    EXPECT_GT(std::abs(true_pose.translation().normalized().dot(estimated_pose.translation().normalized())), 0.9);
    // noise-free data: the DLT pose is exact
    EXPECT_LT(rotation_angle_small(true_pose.linear().transpose() * estimated_pose.linear()), 1e-9);
    EXPECT_LT((true_pose.translation() - estimated_pose.translation()).norm(), 1e-9);
    // fewer than four points: identity (planarpose_linear.cpp:55-57)
    view.resize(3);
    EXPECT_TRUE(estimate_planar_pose(view, intrinsics).isApprox(Eigen::Isometry3d::Identity()));
}

int main(int argc, char** argv) { return mini_gtest::run(argc, argv); }
