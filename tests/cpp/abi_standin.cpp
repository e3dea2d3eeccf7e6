// TEST INFRASTRUCTURE ONLY — never shipped, never loaded by the package.
//
// A stand-in for the subset of the C ABI (include/calib_b200.h) that the C++ host layer
// (include/calib_b200_adapter.hpp) calls, answering from the CPU oracle (oracle/liboracle.so) instead
// of the GPU.  Purpose: the CPU suite (`-m "not gpu"`) can run the adapter's packing / unpacking /
// block order / error mapping end to end — the code a maintainer drops into the reference — with no
// device present.  It says nothing about the CUDA path: the same test binary is linked against the
// real libcalib_b200.so in the GPU suite (tests/test_gpu_zzz_cpp_host.py), and that is the parity run.
// With -DSTANDIN_SIMT the linear-stage entry points (RANSAC homography / plane, per-view DLT, Zhang, poses) are
// answered by the product's OWN kernel sources executed on the CPU under the lock-step SIMT shim
// (tests/host_emul/libransac_simt.so, libseed_simt.so) instead of the oracle, so the C++ tests drive the
// real device code of that stage in the CPU suite.
// The argument checks below restate those of the product's entry points (refine_host.cu:118-161) so
// the error-convention tests behave the same on both sides.
#include <cstring>
#include <string>
#include <vector>

#include "calib_b200.h"
#include "oracle_api.h"

extern "C" {
int orc_estimate_intrinsics(int64_t, const int64_t*, const double*, const double*, const double*, const double*, const double*, double*,
                            int32_t*, double*, double*, double*);
int orc_estimate_intrinsics_ransac(int64_t, const int64_t*, const double*, const double*, const double*, const double*, const double*,
                                   const orc_ransac_options*, double*, int32_t*, double*, double*, double*, uint8_t*);
void orc_ref_estimate_planar_pose(int32_t, const double*, const double*, const double*, const double*, const double*, double*);
}

#ifdef STANDIN_SIMT
extern "C" {
int simt_ransac_homography(int64_t, int32_t, const double*, const double*, const double*, const double*, const cal_ransac_options*, int,
                           cal_ransac_result*, uint8_t*);
int simt_ransac_plane(int64_t, int32_t, const double*, const double*, const double*, const cal_ransac_options*, int, cal_plane_ransac_result*,
                      uint8_t*);
int simt_seed_intrinsics(int64_t, const int64_t*, const int32_t*, const double*, const double*, const double*, const double*, int32_t,
                         const cal_seed_options*, double*, int32_t*, int32_t*, double*, double*, double*);
int simt_seed_planar_poses(int64_t, const int64_t*, const int32_t*, const double*, const double*, const double*, const double*, const double*,
                           double*, int32_t*);
}
#endif

#include <cstddef>
// the oracle's descriptor is the prefix of the product's (which adds the optional shared-board fields)
static_assert(offsetof(cal_problem_desc, board_x) == sizeof(orc_problem_desc));
static_assert(sizeof(cal_optim_options) == sizeof(orc_optim_options));
static_assert(sizeof(cal_optim_result) == sizeof(orc_optim_result));
static_assert(sizeof(cal_ransac_options) == sizeof(orc_ransac_options));
static_assert(sizeof(cal_ransac_result) == sizeof(orc_ransac_result));
static_assert(sizeof(cal_plane_ransac_result) == sizeof(orc_plane_result));

// -DSTANDIN_NO_REFINE: the refinement entry points (and cal_last_error / cal_device_count) come from the CPU build of
// the product's own host code and kernels (tests/host_emul/_build/libcalib_b200_simt.so, tests/test_product_on_cpu.py)
// instead of the oracle; this file then only answers the linear stage and AX = XB.
#ifdef STANDIN_NO_REFINE
extern "C" void cal_set_last_error_(const char* msg);
namespace {
cal_status fail(cal_status s, const char* m) { cal_set_last_error_(m); return s; }
}  // namespace
#else
namespace {
thread_local std::string g_err;
cal_status fail(cal_status s, const char* m) { g_err = m; return s; }
}  // namespace
#endif

#ifndef STANDIN_NO_REFINE
struct cal_refine_handle {
    cal_problem_desc d;
    std::vector<double> x, y, u, v, bTg;
    std::vector<int64_t> off;
    std::vector<int32_t> cam, view;
};
#endif
struct cal_axxb_handle {
    std::vector<double> ra, rb, ta, tb;
    double huber;
};

extern "C" {

#ifndef STANDIN_NO_REFINE
const char* cal_last_error(void) { return g_err.c_str(); }
int cal_device_count(void) { return 0; }

cal_status cal_refine_create(const cal_problem_desc* desc, int, cal_refine_handle** out) {
    if (!desc || !out) return fail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    const cal_problem_desc& d = *desc;
    if (d.n_cams <= 0) return fail(CAL_ERR_INVALID_ARGUMENT, "No camera intrinsics provided");
    if (d.kind == CAL_KIND_INTRINSICS && d.n_views < 4)
        return fail(CAL_ERR_INVALID_ARGUMENT, "Insufficient views for calibration (at least 4 required).");
    if (d.n_blocks <= 0 || d.n_obs <= 0) return fail(CAL_ERR_INVALID_ARGUMENT, "No observations provided");
    auto* h = new cal_refine_handle;
    h->d = d;
    if (d.board_n > 0) {  // shared-board form: the oracle wants per-observation object points
        if (!d.board_x || !d.board_y) { delete h; return fail(CAL_ERR_INVALID_ARGUMENT, "null observation arrays"); }
        for (int64_t b = 0; b < d.n_blocks; ++b) {
            if (d.block_offset[b + 1] - d.block_offset[b] != d.board_n) {
                delete h;
                return fail(CAL_ERR_INVALID_ARGUMENT, "shared-board form: every residual block must hold exactly board_n observations");
            }
            h->x.insert(h->x.end(), d.board_x, d.board_x + d.board_n); h->y.insert(h->y.end(), d.board_y, d.board_y + d.board_n);
        }
    } else {
        h->x.assign(d.obj_x, d.obj_x + d.n_obs); h->y.assign(d.obj_y, d.obj_y + d.n_obs);
    }
    h->u.assign(d.img_u, d.img_u + d.n_obs); h->v.assign(d.img_v, d.img_v + d.n_obs);
    h->off.assign(d.block_offset, d.block_offset + d.n_blocks + 1);
    h->cam.assign(d.block_cam, d.block_cam + d.n_blocks);
    if (d.block_view) h->view.assign(d.block_view, d.block_view + d.n_blocks); else h->view.assign(d.n_blocks, 0);
    if (d.block_b_se3_g) h->bTg.assign(d.block_b_se3_g, d.block_b_se3_g + 12 * d.n_blocks);
    h->d.obj_x = h->x.data(); h->d.obj_y = h->y.data(); h->d.img_u = h->u.data(); h->d.img_v = h->v.data();
    h->d.block_offset = h->off.data(); h->d.block_cam = h->cam.data(); h->d.block_view = h->view.data();
    h->d.block_b_se3_g = h->bTg.empty() ? nullptr : h->bTg.data();
    *out = h;
    return CAL_OK;
}
void cal_refine_destroy(cal_refine_handle* h) { delete h; }
#if !defined(STANDIN_NO_REFINE)   // (the CPU build of the product's host code brings its own)
cal_status cal_host_borrow(size_t bytes, void** out) { *out = std::malloc(bytes ? bytes : 1); return *out ? CAL_OK : CAL_ERR_CUDA; }
void cal_host_return(void* p) { std::free(p); }
#endif

cal_status cal_refine_solve(cal_refine_handle* h, const cal_optim_options* o, double* x, cal_optim_result* r, double* cov) {
    orc_optim_result rr{};
    const int rc = orc_refine_solve(reinterpret_cast<const orc_problem_desc*>(&h->d), reinterpret_cast<const orc_optim_options*>(o), x, &rr,
                                    cov, 0);
    std::memcpy(r, &rr, sizeof rr);
    return rc == 0 ? CAL_OK : fail(CAL_ERR_RUNTIME, "oracle solve failed");
}

#endif  // STANDIN_NO_REFINE

cal_status cal_axxb_create_from_poses(int64_t n, const double* bg, const double* ct, double min_angle_deg, int, double, double huber, int,
                                      cal_axxb_handle** out, int64_t* kept) {
    if (n < 2) return fail(CAL_ERR_RUNTIME, "Inconsistent hand-eye input sizes");
    const int64_t cnt = orc_build_all_pairs(n, bg, ct, min_angle_deg, nullptr, nullptr, nullptr, nullptr);
    if (cnt <= 0) return fail(CAL_ERR_RUNTIME, "No valid motion pairs after filtering. Try lowering min_angle_deg or check your data.");
    auto* h = new cal_axxb_handle;
    h->ra.resize(9 * cnt); h->rb.resize(9 * cnt); h->ta.resize(3 * cnt); h->tb.resize(3 * cnt);
    h->huber = huber;
    orc_build_all_pairs(n, bg, ct, min_angle_deg, h->ra.data(), h->rb.data(), h->ta.data(), h->tb.data());
    if (kept) *kept = cnt;
    *out = h;
    return CAL_OK;
}
void cal_axxb_destroy(cal_axxb_handle* h) { delete h; }
cal_status cal_axxb_solve(cal_axxb_handle* h, const cal_optim_options* o, double* x7, cal_optim_result* r, double* cov49) {
    const orc_axxb_desc d{static_cast<int64_t>(h->ta.size() / 3), h->ra.data(), h->rb.data(), h->ta.data(), h->tb.data(), h->huber};
    orc_optim_result rr{};
    const int rc = orc_axxb_solve(&d, reinterpret_cast<const orc_optim_options*>(o), x7, &rr, cov49);
    std::memcpy(r, &rr, sizeof rr);
    return rc == 0 ? CAL_OK : fail(CAL_ERR_RUNTIME, "oracle solve failed");
}

cal_status cal_ransac_homography_batch(int64_t np, int32_t n, const double* x, const double* y, const double* u, const double* v,
                                       const cal_ransac_options* o, int per, int, cal_ransac_result* res, uint8_t* mask) {
#ifdef STANDIN_SIMT
    std::vector<uint8_t> m(static_cast<size_t>(np) * n);
    if (simt_ransac_homography(np, n, x, y, u, v, o, per, res, m.data())) return fail(CAL_ERR_RUNTIME, "simt launch failed");
    if (mask) std::memcpy(mask, m.data(), m.size());
    return CAL_OK;
#endif
    orc_ransac_homography_batch(np, n, x, y, u, v, reinterpret_cast<const orc_ransac_options*>(o), per,
                                reinterpret_cast<orc_ransac_result*>(res), mask, 0);
    return CAL_OK;
}
cal_status cal_ransac_plane_batch(int64_t np, int32_t n, const double* x, const double* y, const double* z, const cal_ransac_options* o, int per,
                                  int, cal_plane_ransac_result* res, uint8_t* mask) {
#ifdef STANDIN_SIMT
    std::vector<uint8_t> m(static_cast<size_t>(np) * n);
    if (simt_ransac_plane(np, n, x, y, z, o, per, res, m.data())) return fail(CAL_ERR_RUNTIME, "simt launch failed");
    if (mask) std::memcpy(mask, m.data(), m.size());
    return CAL_OK;
#endif
    orc_ransac_plane_batch(np, n, x, y, z, reinterpret_cast<const orc_ransac_options*>(o), per, reinterpret_cast<orc_plane_result*>(res), mask,
                           0);
    return CAL_OK;
}

cal_status cal_seed_intrinsics_ransac(int64_t nv, const int64_t* off, const int32_t*, const double* x, const double* y, const double* u,
                                      const double* v, int32_t n_cams, const cal_seed_options* so, const cal_ransac_options* ro, int, double* kmtx,
                                      int32_t* cam_ok, int32_t* view_ok, double* hmtx, double* rms, double* poses, uint8_t* mask) {
    if (n_cams != 1) return fail(CAL_ERR_INVALID_ARGUMENT, "stand-in: one camera");
    double b[10];
    if (so && so->use_bounds) {
        const double t[10] = {so->fx_min, so->fx_max, so->fy_min, so->fy_max, so->cx_min, so->cx_max, so->cy_min, so->cy_max, so->skew_min, so->skew_max};
        std::memcpy(b, t, sizeof t);
    }
    const double* bp = (so && so->use_bounds) ? b : nullptr;
    std::vector<int32_t> ok(nv); std::vector<double> H(9 * nv), r(nv), P(12 * nv);
    for (int i = 0; i < 5; ++i) kmtx[i] = 0.0;
#ifdef STANDIN_SIMT
    if (!ro) {
        std::vector<int32_t> cam0(nv, 0);
        simt_seed_intrinsics(nv, off, cam0.data(), x, y, u, v, 1, so, kmtx, cam_ok, ok.data(), H.data(), r.data(), P.data());
    } else
#endif
    *cam_ok = ro ? orc_estimate_intrinsics_ransac(nv, off, x, y, u, v, bp, reinterpret_cast<const orc_ransac_options*>(ro), kmtx, ok.data(),
                                                  H.data(), r.data(), P.data(), mask)
                 : orc_estimate_intrinsics(nv, off, x, y, u, v, bp, kmtx, ok.data(), H.data(), r.data(), P.data());
    if (view_ok) std::memcpy(view_ok, ok.data(), sizeof(int32_t) * nv);
    if (hmtx) std::memcpy(hmtx, H.data(), sizeof(double) * 9 * nv);
    if (rms) std::memcpy(rms, r.data(), sizeof(double) * nv);
    if (poses) std::memcpy(poses, P.data(), sizeof(double) * 12 * nv);
    return CAL_OK;
}
cal_status cal_seed_intrinsics(int64_t nv, const int64_t* off, const int32_t* cam, const double* x, const double* y, const double* u,
                               const double* v, int32_t n_cams, const cal_seed_options* so, int dev, double* kmtx, int32_t* cam_ok,
                               int32_t* view_ok, double* hmtx, double* rms, double* poses) {
    return cal_seed_intrinsics_ransac(nv, off, cam, x, y, u, v, n_cams, so, nullptr, dev, kmtx, cam_ok, view_ok, hmtx, rms, poses, nullptr);
}
cal_status cal_seed_planar_poses(int64_t nv, const int64_t* off, const int32_t* cam, const double* x, const double* y, const double* u,
                                 const double* v, int32_t, const double* kmtx, int, double* poses, int32_t*) {
#ifdef STANDIN_SIMT
    std::vector<int32_t> ok(nv);
    simt_seed_planar_poses(nv, off, cam, x, y, u, v, kmtx, poses, ok.data());
    return CAL_OK;
#endif
    for (int64_t k = 0; k < nv; ++k) {
        const int64_t o = off[k];
        orc_ref_estimate_planar_pose(static_cast<int32_t>(off[k + 1] - o), x + o, y + o, u + o, v + o, kmtx + 5 * (cam ? cam[k] : 0), poses + 12 * k);
    }
    return CAL_OK;
}

}  // extern "C"
