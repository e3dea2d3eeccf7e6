// Source-compatible C++20 adapter: re-exposes the reference's refinement entry
// points on top of the C ABI of calib_b200.h, so the reference's callers
// (facades, stages, tests) keep compiling unchanged.
//
//   calib::optimize_intrinsics<CameraT>   include/calib/estimation/optim/intrinsics.h:35-39
//   calib::optimize_extrinsics<CameraT>   include/calib/estimation/optim/extrinsics.h:29-34
//   calib::optimize_bundle<CameraT>       include/calib/estimation/optim/bundle.h:58-63
//   calib::optimize_handeye               include/calib/estimation/optim/handeye.h:40-43
//   calib::estimate_homography (RANSAC)   include/calib/estimation/linear/homography.h:22-24
//   calib::fit_plane_ransac               include/calib/estimation/linear/planefit.h:23-24
//
// Two build modes, one source:
//   * inside the reference tree (Eigen and the calib/ headers are on the include path) the option,
//     result and camera-model types are the reference's own;
//   * stand-alone (this repository's image has neither Eigen nor Ceres) the same names come from
//     calib_b200_mini.hpp, and the linear-stage entry points that the reference defines in
//     calib_estimation_linear (estimate_homography, estimate_planar_pose, estimate_intrinsics,
//     fit_plane_ransac) are provided here under their reference names as well.  This is the mode the
//     C++ host tests of this repository compile (tests/cpp/reference_tests.cpp).
// It contains no arithmetic of the hot path: packing into SoA/CSR, the C call, unpacking, and the
// error mapping (CAL_ERR_INVALID_ARGUMENT -> std::invalid_argument, CAL_ERR_RUNTIME ->
// std::runtime_error, anything else -> std::runtime_error).
#pragma once

#include <algorithm>
#include <array>
#include <atomic>
#include <cstdlib>
#include <optional>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#if __has_include(<Eigen/Core>) && __has_include("calib/estimation/optim/bundle.h")
#define CALIB_B200_REFERENCE_TREE 1
#include "calib/estimation/linear/handeye.h"
#include "calib/estimation/linear/homography.h"
#include "calib/estimation/linear/planefit.h"
#include "calib/estimation/optim/bundle.h"
#include "calib/estimation/optim/extrinsics.h"
#include "calib/estimation/optim/handeye.h"
#include "calib/estimation/optim/intrinsics.h"
#include "calib/models/scheimpflug.h"
#else
#define CALIB_B200_REFERENCE_TREE 0
#include "calib_b200_mini.hpp"
#endif
#include "calib_b200.h"

// default arguments live on the reference's own declarations inside its tree
#if CALIB_B200_REFERENCE_TREE
#define CALIB_B200_DEFAULT(x)
#else
#define CALIB_B200_DEFAULT(x) = x
#endif

namespace calib::b200 {

inline void check(cal_status s) {
    if (s == CAL_OK) return;
    const std::string msg = cal_last_error();
    if (s == CAL_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}

// row-major C buffers -> the matrix types of the result structs (element-wise: no Eigen::Map, so the same
// lines compile against Eigen and against calib_b200_mini.hpp)
inline Eigen::Matrix3d mat3_from_rowmajor(const double* p) {
    Eigen::Matrix3d m;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m(i, j) = p[3 * i + j];
    return m;
}
inline Eigen::MatrixXd dense_from_rowmajor(const double* p, Eigen::Index n) {
    Eigen::MatrixXd m(n, n);
    for (Eigen::Index i = 0; i < n; ++i) for (Eigen::Index j = 0; j < n; ++j) m(i, j) = p[i * n + j];
    return m;
}

template <class CameraT> constexpr int model_of() {
    return CameraTraits<CameraT>::param_count == 12 ? CAL_MODEL_SCHEIMPFLUG_BC5 : CAL_MODEL_PINHOLE_BC5;
}

// populate_quat_tran (src/estimation/detail/observationutils.h:43-48)
inline void push_pose(const Eigen::Isometry3d& T, double* q, double* t) {
    Eigen::Quaterniond q0(T.linear());
    q[0] = q0.w(); q[1] = q0.x(); q[2] = q0.y(); q[3] = q0.z();
    t[0] = T.translation().x(); t[1] = T.translation().y(); t[2] = T.translation().z();
}
// restore_pose (observationutils.h:50-62)
inline Eigen::Isometry3d pop_pose(const double* q, const double* t) {
    Eigen::Quaterniond qq(q[0], q[1], q[2], q[3]);
    qq.normalize();
    Eigen::Isometry3d T = Eigen::Isometry3d::Identity();
    T.linear() = qq.toRotationMatrix();
    T.translation() = Eigen::Vector3d(t[0], t[1], t[2]);
    return T;
}

// SoA / CSR packing of the caller's views.  Every view of a calibration usually observes the same board, so
// object_xy repeats from view to view: while that holds only ONE board is kept (the shared-board form of
// cal_problem_desc, half the bytes to the GPU); the first view that differs expands what was packed so far to
// the per-observation form.  Exact comparison — a board with one corner missing in some view is not shared.
struct Soa {
    std::vector<double> x, y, u, v;   // x, y stay empty while the views share a board
    std::vector<double> board_x, board_y;
    bool shared = std::getenv("CALIB_B200_PER_OBSERVATION") == nullptr;   // set to force the per-observation form
    std::vector<int64_t> off{0};
    std::vector<int32_t> cam, view;
    std::vector<double> bTg;
    bool on_board(const PlanarView& pv) const {
        if (pv.size() != board_x.size()) return false;
        for (size_t i = 0; i < pv.size(); ++i)
            if (pv[i].object_xy.x() != board_x[i] || pv[i].object_xy.y() != board_y[i]) return false;
        return true;
    }
    void add(const PlanarView& pv, int c, int vw) {
        if (shared) {
            if (cam.empty() && board_x.empty()) {
                for (const auto& ob : pv) { board_x.push_back(ob.object_xy.x()); board_y.push_back(ob.object_xy.y()); }
            } else if (!on_board(pv)) {
                shared = false;
                for (size_t b = 0; b < cam.size(); ++b) { x.insert(x.end(), board_x.begin(), board_x.end()); y.insert(y.end(), board_y.begin(), board_y.end()); }
            }
        }
        for (const auto& ob : pv) {
            if (!shared) { x.push_back(ob.object_xy.x()); y.push_back(ob.object_xy.y()); }
            u.push_back(ob.image_uv.x()); v.push_back(ob.image_uv.y());
        }
        off.push_back(static_cast<int64_t>(u.size())); cam.push_back(c); view.push_back(vw);
    }
    void fill(cal_problem_desc& d) const {
        d.n_blocks = static_cast<int64_t>(cam.size()); d.n_obs = static_cast<int64_t>(u.size());
        if (shared && !board_x.empty()) {
            d.board_x = board_x.data(); d.board_y = board_y.data(); d.board_n = static_cast<int32_t>(board_x.size());
        } else {
            d.obj_x = x.data(); d.obj_y = y.data();
        }
        d.img_u = u.data(); d.img_v = v.data();
        d.block_offset = off.data(); d.block_cam = cam.data(); d.block_view = view.data();
        d.block_b_se3_g = bTg.empty() ? nullptr : bTg.data();
    }
};

// Large inputs (millions of observations): the AoS views — one heap vector per view — are packed by several threads
// straight into page-locked staging borrowed from the library (cal_host_borrow: a process-wide pool, so a long-running
// caller page-locks once), in ONE pass over the observations that also checks whether every view shows the same board
// (then only u, v travel: 16 instead of 32 bytes per observation).  A view that differs sends the call down the general
// path (Soa::add).  Same cal_problem_desc as the serial packing, bit for bit.
struct BundleStage {
    void* block = nullptr;
    double *u = nullptr, *v = nullptr, *bTg = nullptr, *bx = nullptr, *by = nullptr;
    int64_t* off = nullptr; int32_t* cam = nullptr;
    int64_t n_blocks = 0, n_obs = 0; int32_t board_n = 0;
    ~BundleStage() { if (block) cal_host_return(block); }
    // false: the views do not share one board (or a view is empty): nothing usable was produced
    bool pack(const std::vector<BundleObservation>& obs) {
        n_blocks = static_cast<int64_t>(obs.size());
        board_n = static_cast<int32_t>(obs[0].view.size());
        if (board_n == 0) return false;
        for (const auto& o : obs) if (static_cast<int32_t>(o.view.size()) != board_n) return false;   // ragged: general path
        n_obs = n_blocks * board_n;
        const size_t bytes = sizeof(double) * (2 * static_cast<size_t>(n_obs) + 12 * static_cast<size_t>(n_blocks) + 2 * static_cast<size_t>(board_n)) +
                             sizeof(int64_t) * (static_cast<size_t>(n_blocks) + 1) + sizeof(int32_t) * static_cast<size_t>(n_blocks) + 64;
        check(cal_host_borrow(bytes, &block));
        u = static_cast<double*>(block); v = u + n_obs; bTg = v + n_obs; bx = bTg + 12 * n_blocks; by = bx + board_n;
        off = reinterpret_cast<int64_t*>(by + board_n); cam = reinterpret_cast<int32_t*>(off + n_blocks + 1);
        for (int32_t i = 0; i < board_n; ++i) { bx[i] = obs[0].view[static_cast<size_t>(i)].object_xy.x(); by[i] = obs[0].view[static_cast<size_t>(i)].object_xy.y(); }
        const unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
        std::vector<std::thread> th;
        std::atomic<bool> same{true};
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&, t] {
                const int64_t b0 = n_blocks * t / nt, b1 = n_blocks * (t + 1) / nt;
                bool ok = true;
                for (int64_t b = b0; b < b1; ++b) {
                    const BundleObservation& o = obs[static_cast<size_t>(b)];
                    double* pu = u + b * board_n; double* pv = v + b * board_n;
                    for (int32_t i = 0; i < board_n; ++i) {
                        const auto& p = o.view[static_cast<size_t>(i)];
                        ok = ok && p.object_xy.x() == bx[i] && p.object_xy.y() == by[i];
                        pu[i] = p.image_uv.x(); pv[i] = p.image_uv.y();
                    }
                    off[b] = b * board_n; cam[b] = static_cast<int32_t>(o.camera_index);
                    const Eigen::Matrix3d R = o.b_se3_g.linear(); const Eigen::Vector3d tr = o.b_se3_g.translation();
                    double* g = bTg + 12 * b;
                    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) g[3 * i + j] = R(i, j);
                    for (int i = 0; i < 3; ++i) g[9 + i] = tr(i);
                }
                if (!ok) same.store(false);
            });
        for (auto& t : th) t.join();
        off[n_blocks] = n_obs;
        return same.load();
    }
    void fill(cal_problem_desc& d) const {
        d.n_blocks = n_blocks; d.n_obs = n_obs;
        d.board_x = bx; d.board_y = by; d.board_n = board_n;
        d.img_u = u; d.img_v = v; d.block_offset = off; d.block_cam = cam; d.block_b_se3_g = bTg;
    }
};
constexpr size_t kStageThreshold = size_t{1} << 20;   // observations from which the staged, threaded packing pays

inline cal_optim_options to_c(const OptimOptions& o) {
    return cal_optim_options{static_cast<int32_t>(o.optimizer), o.max_iterations, o.epsilon, o.compute_covariance ? 1 : 0,
                             o.verbose ? 1 : 0, 0, 0};
}

inline void run(const cal_problem_desc& d, const OptimOptions& core, std::vector<double>& x, OptimResult& out) {
    cal_refine_handle* h = nullptr;
    check(cal_refine_create(&d, /*device=*/0, &h));
    const cal_optim_options co = to_c(core);
    cal_optim_result r{};
    std::vector<double> cov(core.compute_covariance ? x.size() * x.size() : 0);
    const cal_status s = cal_refine_solve(h, &co, x.data(), &r, cov.empty() ? nullptr : cov.data());
    cal_refine_destroy(h);
    check(s);
    out.success = r.success != 0; out.final_cost = r.final_cost; out.report = r.report;
    if (r.covariance_ok) {
        out.covariance = dense_from_rowmajor(cov.data(), static_cast<Eigen::Index>(x.size()));
    }
}

}  // namespace calib::b200

namespace calib {

template <camera_model CameraT>
auto optimize_bundle(const std::vector<BundleObservation>& observations, const std::vector<CameraT>& initial_cameras,
                     const std::vector<Eigen::Isometry3d>& init_g_se3_c, const Eigen::Isometry3d& init_b_se3_t,
                     const BundleOptions& opts CALIB_B200_DEFAULT({})) -> BundleResult<CameraT> {
    constexpr int P = CameraTraits<CameraT>::param_count;
    if (initial_cameras.empty()) throw std::invalid_argument("No camera intrinsics provided");  // bundle.cpp:139-141
    if (observations.empty()) throw std::invalid_argument("No observations provided");            // bundle.cpp:142-144
    if (init_g_se3_c.size() != initial_cameras.size())
        throw std::invalid_argument("optimize_bundle: one initial hand-eye pose per camera required");
    b200::Soa s;
    b200::BundleStage stage;
    cal_problem_desc d{};
    d.kind = CAL_KIND_BUNDLE; d.model = b200::model_of<CameraT>(); d.n_cams = static_cast<int>(initial_cameras.size());
    d.optimize_intrinsics = opts.optimize_intrinsics; d.optimize_skew = opts.optimize_skew;
    d.optimize_target_pose = opts.optimize_target_pose; d.optimize_hand_eye = opts.optimize_hand_eye;
    d.huber_delta = opts.core.huber_delta;
    const bool staged = s.shared && observations.size() * observations[0].view.size() >= b200::kStageThreshold && stage.pack(observations);
    if (staged) {
        stage.fill(d);
    } else {
        for (const auto& ob : observations) {
            s.add(ob.view, static_cast<int>(ob.camera_index), -1);
            const Eigen::Matrix3d R = ob.b_se3_g.linear(); const Eigen::Vector3d t = ob.b_se3_g.translation();
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) s.bTg.push_back(R(i, j));
            for (int i = 0; i < 3; ++i) s.bTg.push_back(t(i));
        }
        s.fill(d);
    }
    const size_t nc = initial_cameras.size();
    std::vector<double> x(nc * (P + 7) + 7);
    for (size_t c = 0; c < nc; ++c) {
        std::array<double, P> a{}; CameraTraits<CameraT>::to_array(initial_cameras[c], a);
        std::copy(a.begin(), a.end(), x.begin() + c * P);
        b200::push_pose(init_g_se3_c[c], &x[nc * P + 4 * c], &x[nc * (P + 4) + 3 * c]);
    }
    b200::push_pose(init_b_se3_t, &x[nc * (P + 7)], &x[nc * (P + 7) + 4]);
    BundleResult<CameraT> result;
    b200::run(d, opts.core, x, result.core);
    result.cameras.resize(nc); result.g_se3_c.resize(nc);
    for (size_t c = 0; c < nc; ++c) {
        result.cameras[c] = CameraTraits<CameraT>::template from_array<double>(&x[c * P]);
        result.g_se3_c[c] = b200::pop_pose(&x[nc * P + 4 * c], &x[nc * (P + 4) + 3 * c]);
    }
    result.b_se3_t = b200::pop_pose(&x[nc * (P + 7)], &x[nc * (P + 7) + 4]);
    return result;
}

template <camera_model CameraT>
auto optimize_intrinsics(const std::vector<PlanarView>& views, const CameraT& init_camera,
                         std::vector<Eigen::Isometry3d> init_c_se3_t, const IntrinsicsOptimOptions& opts CALIB_B200_DEFAULT({}))
    -> IntrinsicsOptimizationResult<CameraT> {
    constexpr int P = CameraTraits<CameraT>::param_count;
    if (views.size() < 4)  // validate_input, intrinsics.cpp:92-96 (before the start poses are touched)
        throw std::invalid_argument("Insufficient views for calibration (at least 4 required).");
    if (init_c_se3_t.size() != views.size())  // the reference indexes one pose block per view (intrinsics.cpp:68-75)
        throw std::invalid_argument("optimize_intrinsics: one initial pose per view required");
    b200::Soa s;
    for (size_t v = 0; v < views.size(); ++v) s.add(views[v], 0, static_cast<int>(v));
    cal_problem_desc d{};
    d.kind = CAL_KIND_INTRINSICS; d.model = b200::model_of<CameraT>(); d.n_cams = 1; d.n_views = static_cast<int>(views.size());
    d.optimize_intrinsics = 1; d.optimize_skew = opts.optimize_skew; d.huber_delta = opts.core.huber_delta;
    s.fill(d);
    const size_t nv = views.size();
    std::vector<double> x(P + 7 * nv);
    std::array<double, P> a{}; CameraTraits<CameraT>::to_array(init_camera, a);
    std::copy(a.begin(), a.end(), x.begin());
    for (size_t v = 0; v < nv; ++v) b200::push_pose(init_c_se3_t[v], &x[P + 4 * v], &x[P + 4 * nv + 3 * v]);
    IntrinsicsOptimizationResult<CameraT> result;
    b200::run(d, opts.core, x, result.core);
    result.camera = CameraTraits<CameraT>::template from_array<double>(x.data());
    result.c_se3_t.resize(nv);
    for (size_t v = 0; v < nv; ++v) result.c_se3_t[v] = b200::pop_pose(&x[P + 4 * v], &x[P + 4 * nv + 3 * v]);
    return result;
}

template <camera_model CameraT>
auto optimize_extrinsics(const std::vector<MulticamPlanarView>& views, const std::vector<CameraT>& init_cameras,
                         const std::vector<Eigen::Isometry3d>& init_c_se3_r, const std::vector<Eigen::Isometry3d>& init_r_se3_t,
                         const ExtrinsicOptions& opts CALIB_B200_DEFAULT({})) -> ExtrinsicOptimizationResult<CameraT> {
    constexpr int P = CameraTraits<CameraT>::param_count;
    const size_t nc = init_cameras.size(), nv = views.size();
    if (init_c_se3_r.size() != nc || init_r_se3_t.size() != nv)  // extrinsics.cpp:163-171
        throw std::invalid_argument("Incompatible pose vector sizes for joint optimization");
    b200::Soa s;
    for (size_t v = 0; v < nv; ++v)
        if (views[v].size() < nc)  // the reference indexes views[v][cam] for every camera (extrinsics.cpp:91-96)
            throw std::invalid_argument("optimize_extrinsics: every MulticamPlanarView needs one PlanarView per camera");
    for (size_t v = 0; v < nv; ++v)
        for (size_t c = 0; c < nc; ++c)
            if (!views[v][c].empty()) s.add(views[v][c], static_cast<int>(c), static_cast<int>(v));  // extrinsics.cpp:94-96
    cal_problem_desc d{};
    d.kind = CAL_KIND_EXTRINSICS; d.model = b200::model_of<CameraT>(); d.n_cams = static_cast<int>(nc); d.n_views = static_cast<int>(nv);
    d.optimize_intrinsics = opts.optimize_intrinsics; d.optimize_skew = opts.optimize_skew;
    d.optimize_extrinsics = opts.optimize_extrinsics; d.huber_delta = opts.core.huber_delta;
    s.fill(d);
    std::vector<double> x(nc * (P + 7) + 7 * nv);
    const size_t oq = nc * P, ot = oq + 4 * nc, vq = ot + 3 * nc, vt = vq + 4 * nv;
    for (size_t c = 0; c < nc; ++c) {
        std::array<double, P> a{}; CameraTraits<CameraT>::to_array(init_cameras[c], a);
        std::copy(a.begin(), a.end(), x.begin() + c * P);
        b200::push_pose(init_c_se3_r[c], &x[oq + 4 * c], &x[ot + 3 * c]);
    }
    for (size_t v = 0; v < nv; ++v) b200::push_pose(init_r_se3_t[v], &x[vq + 4 * v], &x[vt + 3 * v]);
    ExtrinsicOptimizationResult<CameraT> result;
    b200::run(d, opts.core, x, result.core);
    result.cameras.resize(nc); result.c_se3_r.resize(nc); result.r_se3_t.resize(nv);
    for (size_t c = 0; c < nc; ++c) {
        result.cameras[c] = CameraTraits<CameraT>::template from_array<double>(&x[c * P]);
        result.c_se3_r[c] = b200::pop_pose(&x[oq + 4 * c], &x[ot + 3 * c]);
    }
    for (size_t v = 0; v < nv; ++v) result.r_se3_t[v] = b200::pop_pose(&x[vq + 4 * v], &x[vt + 3 * v]);
    return result;
}

inline auto optimize_handeye(const std::vector<Eigen::Isometry3d>& base_se3_gripper,
                             const std::vector<Eigen::Isometry3d>& camera_se3_target,
                             const Eigen::Isometry3d& init_gripper_se3_ref, const OptimOptions& options CALIB_B200_DEFAULT({})) -> HandeyeResult {
    // build_all_pairs (linear/handeyedlt.cpp:51-81) runs on the device: the n (n - 1) / 2 motion pairs are
    // formed on the fly in every pass (0.5 deg / reject-parallel / 1e-3 are optimize_handeye's own arguments,
    // handeye.cpp:63-64 with the defaults of linear/handeye.h)
    if (base_se3_gripper.size() != camera_se3_target.size()) throw std::runtime_error("Inconsistent hand-eye input sizes");
    const auto np = static_cast<int64_t>(base_se3_gripper.size());
    std::vector<double> g(12 * base_se3_gripper.size()), c(12 * camera_se3_target.size());
    auto pack = [](const Eigen::Isometry3d& T, double* o) {
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o[3 * i + j] = T.linear()(i, j);
        for (int i = 0; i < 3; ++i) o[9 + i] = T.translation()(i);
    };
    for (int64_t i = 0; i < np; ++i) { pack(base_se3_gripper[i], &g[12 * i]); pack(camera_se3_target[i], &c[12 * i]); }
    cal_axxb_handle* h = nullptr;
    int64_t kept = 0;
    b200::check(cal_axxb_create_from_poses(np, g.data(), c.data(), 0.5, 1, 1e-3, options.huber_delta, 0, &h, &kept));
    double x[7]; b200::push_pose(init_gripper_se3_ref, x, x + 4);
    const cal_optim_options co = b200::to_c(options);
    cal_optim_result r{}; double cov[49];
    const cal_status s = cal_axxb_solve(h, &co, x, &r, options.compute_covariance ? cov : nullptr);
    cal_axxb_destroy(h);
    b200::check(s);
    HandeyeResult result;
    result.core.success = r.success != 0; result.core.final_cost = r.final_cost; result.core.report = r.report;
    if (r.covariance_ok) result.core.covariance = b200::dense_from_rowmajor(cov, 7);
    result.g_se3_c = b200::pop_pose(x, x + 4);
    return result;
}

// estimate_intrinsics(views, opts) (linear/intrinsics.h:58-59, src/estimation/linear/intrinsicsdlt.cpp:101-145):
// all views in one batched call; with opts.homography_ransac every view's homography comes from the batched
// RANSAC kernel (intrinsicsdlt.cpp:50-64), else from the DLT over all its points.
inline auto estimate_intrinsics_b200(const std::vector<PlanarView>& views, const IntrinsicsEstimOptions& opts) -> IntrinsicsEstimateResult {
    IntrinsicsEstimateResult result;
    if (views.empty()) return result;
    std::vector<int64_t> off(views.size() + 1, 0);
    for (size_t k = 0; k < views.size(); ++k) off[k + 1] = off[k] + static_cast<int64_t>(views[k].size());
    const auto n_obs = static_cast<size_t>(off.back());
    if (n_obs == 0) return result;  // no view has a homography: Zhang fails (intrinsicsdlt.cpp:116-120)
    std::vector<double> x(n_obs), y(n_obs), u(n_obs), v(n_obs);
    for (size_t k = 0; k < views.size(); ++k)
        for (size_t i = 0; i < views[k].size(); ++i) {
            const auto o = static_cast<size_t>(off[k]) + i;
            x[o] = views[k][i].object_xy.x(); y[o] = views[k][i].object_xy.y(); u[o] = views[k][i].image_uv.x(); v[o] = views[k][i].image_uv.y();
        }
    std::vector<int32_t> cam(views.size(), 0), ok(views.size());
    cal_seed_options so{};
    if (opts.bounds) {
        const auto& b = *opts.bounds;
        so.use_bounds = 1; so.fx_min = b.fx_min; so.fx_max = b.fx_max; so.fy_min = b.fy_min; so.fy_max = b.fy_max;
        so.cx_min = b.cx_min; so.cx_max = b.cx_max; so.cy_min = b.cy_min; so.cy_max = b.cy_max; so.skew_min = b.skew_min; so.skew_max = b.skew_max;
    }
    cal_ransac_options ro{};
    std::vector<uint8_t> mask;
    if (opts.homography_ransac) {
        const auto& r = *opts.homography_ransac;
        ro = cal_ransac_options{r.max_iters, r.min_inliers, r.thresh, r.confidence, r.seed, r.refit_on_inliers ? 1 : 0, 0};
        mask.resize(n_obs);
    }
    double k5[5]; int32_t cam_ok = 0;
    std::vector<double> H(9 * views.size()), rms(views.size()), poses(12 * views.size());
    b200::check(cal_seed_intrinsics_ransac(static_cast<int64_t>(views.size()), off.data(), cam.data(), x.data(), y.data(), u.data(), v.data(), 1,
                                           &so, opts.homography_ransac ? &ro : nullptr, 0, k5, &cam_ok, ok.data(), H.data(), rms.data(),
                                           poses.data(), mask.empty() ? nullptr : mask.data()));
    if (!cam_ok) return result;
    result.success = true;
    result.kmtx = CameraMatrix{k5[0], k5[1], k5[2], k5[3], k5[4]};
    for (size_t k = 0; k < views.size(); ++k) {
        if (!ok[k]) continue;  // estimate_intrinsics keeps only the views whose homography succeeded (:109-113)
        ViewEstimateData ved;
        ved.view_index = k;
        ved.forward_rms_px = rms[k];
        ved.homography.success = true;
        ved.homography.hmtx = b200::mat3_from_rowmajor(&H[9 * k]);
        ved.homography.symmetric_rms_px = rms[k];
        for (size_t i = 0; i < views[k].size(); ++i)
            if (mask.empty() || mask[static_cast<size_t>(off[k]) + i]) ved.homography.inliers.push_back(static_cast<int>(i));
        ved.c_se3_t.linear() = b200::mat3_from_rowmajor(&poses[12 * k]);
        ved.c_se3_t.translation() = Eigen::Vector3d(poses[12 * k + 9], poses[12 * k + 10], poses[12 * k + 11]);
        result.views.push_back(std::move(ved));
    }
    return result;
}

// estimate_planar_pose(view, CameraMatrix) (src/estimation/linear/planarpose_linear.cpp:54-76) for many views in one call
inline auto estimate_planar_poses_b200(const std::vector<PlanarView>& views, const CameraMatrix& k) -> std::vector<Eigen::Isometry3d> {
    std::vector<Eigen::Isometry3d> out(views.size(), Eigen::Isometry3d::Identity());
    std::vector<int64_t> off(views.size() + 1, 0);
    for (size_t i = 0; i < views.size(); ++i) off[i + 1] = off[i] + static_cast<int64_t>(views[i].size());
    const auto n_obs = static_cast<size_t>(off.back());
    if (n_obs == 0) return out;  // fewer than 4 points: identity (planarpose_linear.cpp:55-57)
    std::vector<double> x(n_obs), y(n_obs), u(n_obs), v(n_obs);
    for (size_t i = 0, o = 0; i < views.size(); ++i)
        for (const auto& ob : views[i]) { x[o] = ob.object_xy.x(); y[o] = ob.object_xy.y(); u[o] = ob.image_uv.x(); v[o] = ob.image_uv.y(); ++o; }
    std::vector<int32_t> cam(views.size(), 0);
    const double k5[5] = {k.fx, k.fy, k.cx, k.cy, k.skew};
    std::vector<double> poses(12 * views.size());
    b200::check(cal_seed_planar_poses(static_cast<int64_t>(views.size()), off.data(), cam.data(), x.data(), y.data(), u.data(), v.data(), 1, k5, 0,
                                      poses.data(), nullptr));
    for (size_t i = 0; i < views.size(); ++i) {
        out[i].linear() = b200::mat3_from_rowmajor(&poses[12 * i]);
        out[i].translation() = Eigen::Vector3d(poses[12 * i + 9], poses[12 * i + 10], poses[12 * i + 11]);
    }
    return out;
}

// RANSAC branch of estimate_homography (optim/homography.cpp:45-73); the plain DLT branch stays on the host.
inline auto estimate_homography_ransac_b200(const PlanarView& data, const RansacOptions& ro) -> HomographyResult {
    const auto n = static_cast<int32_t>(data.size());
    if (n < 4) return HomographyResult{};  // ransac<> returns an unsuccessful result below k_min_samples (common/ransac.h:126-128)
    std::vector<double> x(n), y(n), u(n), v(n);
    for (int i = 0; i < n; ++i) { x[i] = data[i].object_xy.x(); y[i] = data[i].object_xy.y(); u[i] = data[i].image_uv.x(); v[i] = data[i].image_uv.y(); }
    const cal_ransac_options co{ro.max_iters, ro.min_inliers, ro.thresh, ro.confidence, ro.seed, ro.refit_on_inliers ? 1 : 0, 0};
    cal_ransac_result r{}; std::vector<uint8_t> mask(n);
    b200::check(cal_ransac_homography_batch(1, n, x.data(), y.data(), u.data(), v.data(), &co, 0, 0, &r, mask.data()));
    HomographyResult out;
    out.success = r.success != 0;
    if (out.success) {
        out.hmtx = b200::mat3_from_rowmajor(r.hmtx);
        for (int i = 0; i < n; ++i) if (mask[i]) out.inliers.push_back(i);
        out.symmetric_rms_px = r.symmetric_rms_px;
    }
    return out;
}

// fit_plane_ransac (src/estimation/linear/planefit.cpp:86-104) for one point set, and for a batch of point
// sets of equal size in one kernel launch (seed of set p = opts.seed + p when seed_per_problem).
inline auto fit_plane_ransac_b200(const std::vector<std::vector<Eigen::Vector3d>>& sets, const RansacOptions& ro,
                                  bool seed_per_problem = false) -> std::vector<PlaneRansacResult> {
    std::vector<PlaneRansacResult> out(sets.size());
    if (sets.empty()) return out;
    const auto n = static_cast<int32_t>(sets.front().size());
    for (const auto& s : sets)
        if (static_cast<int32_t>(s.size()) != n) throw std::invalid_argument("fit_plane_ransac_b200: point sets must have equal size");
    if (n < 3) return out;  // planefit.cpp:88-90
    const size_t np = sets.size();
    std::vector<double> x(np * n), y(np * n), z(np * n);
    for (size_t p = 0; p < np; ++p)
        for (int i = 0; i < n; ++i) { x[p * n + i] = sets[p][i].x(); y[p * n + i] = sets[p][i].y(); z[p * n + i] = sets[p][i].z(); }
    const cal_ransac_options co{ro.max_iters, ro.min_inliers, ro.thresh, ro.confidence, ro.seed, ro.refit_on_inliers ? 1 : 0, 0};
    std::vector<cal_plane_ransac_result> r(np); std::vector<uint8_t> mask(np * n);
    b200::check(cal_ransac_plane_batch(static_cast<int64_t>(np), n, x.data(), y.data(), z.data(), &co, seed_per_problem ? 1 : 0, 0,
                                       r.data(), mask.data()));
    for (size_t p = 0; p < np; ++p) {
        if (!r[p].success) continue;
        out[p].success = true;
        out[p].plane = Eigen::Vector4d(r[p].plane[0], r[p].plane[1], r[p].plane[2], r[p].plane[3]);
        for (int i = 0; i < n; ++i) if (mask[p * n + i]) out[p].inliers.push_back(i);
        out[p].inlier_rms = r[p].inlier_rms;
    }
    return out;
}
inline auto fit_plane_ransac_b200(const std::vector<Eigen::Vector3d>& pts, const RansacOptions& ro) -> PlaneRansacResult {
    return fit_plane_ransac_b200(std::vector<std::vector<Eigen::Vector3d>>{pts}, ro).front();
}

#if !CALIB_B200_REFERENCE_TREE
// Stand-alone mode only: the linear-stage entry points under their reference names (inside the reference tree
// these are defined by calib_estimation_linear / optim/homography.cpp, whose bodies call the *_b200 functions above).

// estimate_homography (include/calib/estimation/linear/homography.h:22-24, src/estimation/optim/homography.cpp:62-73)
inline auto estimate_homography(const PlanarView& data, std::optional<RansacOptions> ransac_opts = std::nullopt) -> HomographyResult {
    if (ransac_opts.has_value()) return estimate_homography_ransac_b200(data, *ransac_opts);
    // estimate_homography_dlt (homography.cpp:30-43): HomographyEstimator::fit over all points = the per-view DLT of the seeding kernel
    HomographyResult out;
    const auto n = static_cast<int64_t>(data.size());
    if (n < 4) return out;  // HomographyEstimator::fit -> nullopt (homographyestimator.cpp:126-128)
    std::vector<double> x(n), y(n), u(n), v(n);
    for (int64_t i = 0; i < n; ++i) { x[i] = data[i].object_xy.x(); y[i] = data[i].object_xy.y(); u[i] = data[i].image_uv.x(); v[i] = data[i].image_uv.y(); }
    const int64_t off[2] = {0, n};
    const int32_t cam = 0;
    double k5[5], H[9], rms = 0.0;
    int32_t cam_ok = 0, ok = 0;
    b200::check(cal_seed_intrinsics(1, off, &cam, x.data(), y.data(), u.data(), v.data(), 1, nullptr, 0, k5, &cam_ok, &ok, H, &rms, nullptr));
    if (!ok) return out;
    out.success = true;
    out.hmtx = b200::mat3_from_rowmajor(H);
    out.symmetric_rms_px = rms;
    out.inliers.resize(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) out.inliers[static_cast<size_t>(i)] = static_cast<int>(i);
    return out;
}
// estimate_planar_pose (include/calib/estimation/linear/planarpose.h:33, src/estimation/linear/planarpose_linear.cpp:54-76)
inline auto estimate_planar_pose(PlanarView view, const CameraMatrix& intrinsics) -> Eigen::Isometry3d {
    if (view.size() < 4) return Eigen::Isometry3d::Identity();
    return estimate_planar_poses_b200(std::vector<PlanarView>{std::move(view)}, intrinsics).front();
}
// estimate_intrinsics (include/calib/estimation/linear/intrinsics.h:58-59)
inline auto estimate_intrinsics(const std::vector<PlanarView>& views, const IntrinsicsEstimOptions& opts = {}) -> IntrinsicsEstimateResult {
    return estimate_intrinsics_b200(views, opts);
}
// fit_plane_ransac (include/calib/estimation/linear/planefit.h:23-24)
inline auto fit_plane_ransac(const std::vector<Eigen::Vector3d>& pts, const RansacOptions& opts = {}) -> PlaneRansacResult {
    return fit_plane_ransac_b200(pts, opts);
}
#endif

}  // namespace calib
