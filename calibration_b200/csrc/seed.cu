// Linear seeding stage that feeds the refinement path, batched over views.
//
//   cal_seed_intrinsics     = estimate_intrinsics(views, opts) per camera, no-RANSAC branch
//                             (reference src/estimation/linear/intrinsicsdlt.cpp:101-145):
//                             per-view Hartley-normalised DLT homography (compute_planar_homographies
//                             :33-86 -> HomographyEstimator::fit, homographyestimator.cpp:16-78,123-143),
//                             symmetric_rms_px (:21-31), Zhang's closed form (zhang.cpp:9-208),
//                             sanitize_intrinsics (intrinsics_utils.h:12-62) and per-view
//                             pose_from_homography (posefromhomography.cpp:12-67)
//   cal_seed_planar_poses   = estimate_planar_pose(view, CameraMatrix) per view
//                             (src/estimation/linear/planarpose_linear.cpp:17-76)
//
// Device work decomposition.  k_view_dlt gives 32 consecutive views to one warp.  The O(points)
// passes of a view (centroids, mean distances, the 24 monomial sums that make up A^T A of the DLT
// design matrix, the symmetric transfer error) run warp-cooperatively with coalesced loads and
// shuffle reductions, one view after the other; lane j keeps the sums of view j.  The O(1) dense
// work per view (null vector of the 9x9 normal matrix, de-normalisation, Zhang's two design rows
// or the pose decomposition) then runs lane-private: 32 views at once, nothing computed redundantly.
// Zhang's 6x6 Gram matrix is summed per camera in a fixed order (no atomics); the 6x6 eigenproblem
// and the 3x3 Cholesky that recover K run on the host (O(1) per camera).
//
// Numerical differences to the reference are confined to rounding: null vectors come from the
// normal matrices (inverse iteration for the 9x9 DLT system, Jacobi eigenvalues of Zhang's 6x6 Gram
// matrix) instead of Eigen::JacobiSVD of the design matrices, the orthogonal polar factor (= U V^T of the
// SVD whenever det > 0, which r3 = r1 x r2 guarantees) from a scaled Newton iteration.
#include <cmath>
#include <cstring>
#include <string>
#include <map>
#include <vector>

#include "../../include/calib_b200.h"
#include "dlt.cuh"
#include "seed_host.hpp"
#include "seed_kernels.cuh"

extern "C" void cal_set_last_error_(const char* msg);
extern "C" int cal_device_count(void);

namespace {

cal_status sfail(cal_status s, const std::string& m) { cal_set_last_error_(m.c_str()); return s; }
#define SCUDA(expr)                                                                                   \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) return sfail(CAL_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// device view of a caller array that may live on the host or already on the device
template <class T>
struct DevIn {
    const T* p = nullptr; T* owned = nullptr;
    cudaError_t init(const T* src, size_t n, cudaStream_t st) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeDevice) { p = src; return cudaSuccess; }
        cudaGetLastError();
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&owned), std::max<size_t>(n, 1) * sizeof(T));
        if (e != cudaSuccess) return e;
        p = owned;
        return cudaMemcpyAsync(owned, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
    }
    ~DevIn() { if (owned) cudaFree(owned); }
};
template <class T>
struct DevBuf {
    T* p = nullptr;
    cudaError_t alloc(size_t n) { return cudaMalloc(reinterpret_cast<void**>(&p), std::max<size_t>(n, 1) * sizeof(T)); }
    ~DevBuf() { if (p) cudaFree(p); }
};

struct SeedInputs {
    DevIn<int64_t> off; DevIn<int32_t> cam; DevIn<double> x, y, u, v;
};

cal_status check_views(int64_t n_views, const int64_t* view_offset, const int32_t* view_cam, int32_t n_cams, int64_t* n_obs) {
    if (n_views <= 0 || !view_offset || !view_cam || n_cams <= 0) return sfail(CAL_ERR_INVALID_ARGUMENT, "bad argument");
    if (view_offset[0] != 0) return sfail(CAL_ERR_INVALID_ARGUMENT, "view_offset must start at 0");
    for (int64_t k = 0; k < n_views; ++k) {
        if (view_offset[k + 1] < view_offset[k]) return sfail(CAL_ERR_INVALID_ARGUMENT, "view_offset must be non-decreasing");
        if (view_cam[k] < 0 || view_cam[k] >= n_cams) return sfail(CAL_ERR_INVALID_ARGUMENT, "view_cam out of range");
    }
    *n_obs = view_offset[n_views];
    return CAL_OK;
}

}  // namespace

extern "C" cal_status cal_ransac_homography_batch_dev(int64_t n_problems, int32_t n, const double* x_dev, const double* y_dev,
                                                      const double* u_dev, const double* v_dev, const cal_ransac_options* opts,
                                                      int seed_per_problem, cal_ransac_result* results_dev,
                                                      uint8_t* inlier_mask_dev, float* ms);

extern "C" cal_status cal_seed_intrinsics_ransac(int64_t n_views, const int64_t* view_offset, const int32_t* view_cam, const double* x,
                                                 const double* y, const double* u, const double* v, int32_t n_cams,
                                                 const cal_seed_options* opts, const cal_ransac_options* ransac, int device, double* kmtx,
                                                 int32_t* cam_success, int32_t* view_success, double* hmtx, double* sym_rms,
                                                 double* poses, uint8_t* inlier_mask) {
    if (!x || !y || !u || !v || !kmtx || !cam_success) return sfail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    int64_t n_obs = 0;
    if (cal_status s = check_views(n_views, view_offset, view_cam, n_cams, &n_obs)) return s;
    if (cal_device_count() <= device) return sfail(CAL_ERR_CUDA, "no CUDA device: calib_b200 has no CPU fallback");
    SCUDA(cudaSetDevice(device));
    cudaStream_t st; SCUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
    SeedInputs in;
    SCUDA(in.off.init(view_offset, n_views + 1, st)); SCUDA(in.cam.init(view_cam, n_views, st));
    SCUDA(in.x.init(x, n_obs, st)); SCUDA(in.y.init(y, n_obs, st)); SCUDA(in.u.init(u, n_obs, st)); SCUDA(in.v.init(v, n_obs, st));
    DevBuf<double> dH, drms, drows, dgram, dk, dposes; DevBuf<int32_t> dsucc, dcamok;
    SCUDA(dH.alloc(9 * n_views)); SCUDA(drms.alloc(n_views)); SCUDA(drows.alloc(12 * n_views)); SCUDA(dsucc.alloc(n_views));
    SCUDA(dgram.alloc(22 * n_cams)); SCUDA(dk.alloc(5 * n_cams)); SCUDA(dcamok.alloc(n_cams)); SCUDA(dposes.alloc(12 * n_views));
    const unsigned groups = (unsigned)((n_views + 31) / 32);
    DevBuf<cal_ransac_result> dres; DevBuf<uint8_t> dmask;
    if (ransac) {
        // IntrinsicsEstimOptions::homography_ransac (intrinsicsdlt.cpp:50-64): every view is one problem of the batched
        // RANSAC kernel, all with the seed of the options (each reference call constructs its own engine from opts.seed)
        const int64_t n = view_offset[1] - view_offset[0];
        bool equal = true;
        for (int64_t k = 0; k < n_views; ++k) {
            const int64_t nk = view_offset[k + 1] - view_offset[k];
            if (nk > 0x7fffffff) return sfail(CAL_ERR_INVALID_ARGUMENT, "homography_ransac: view too large");
            equal = equal && nk == n;
        }
        SCUDA(dres.alloc(n_views)); if (inlier_mask) SCUDA(dmask.alloc((size_t)n_obs));
        if (equal) {
            if (n <= 0) return sfail(CAL_ERR_INVALID_ARGUMENT, "homography_ransac: empty views");
            SCUDA(cudaStreamSynchronize(st));  // the inputs are in place before the RANSAC launch on the default stream
            if (cal_status rs = cal_ransac_homography_batch_dev(n_views, (int32_t)n, in.x.p, in.y.p, in.u.p, in.v.p, ransac, 0, dres.p,
                                                                inlier_mask ? dmask.p : nullptr, nullptr)) return rs;
        } else {
            // ragged views: the batched kernel takes problems of one size, so the views are grouped by size; each
            // group is gathered into a dense [group][n] batch, solved by one launch and scattered back (results per
            // view, inlier mask in the observations' own CSR layout).  Views with fewer than four points stay
            // unsuccessful (intrinsicsdlt.cpp:41-45).
            std::map<int64_t, std::vector<int64_t>> by_size;
            for (int64_t k = 0; k < n_views; ++k) by_size[view_offset[k + 1] - view_offset[k]].push_back(k);
            SCUDA(cudaMemsetAsync(dres.p, 0, (size_t)n_views * sizeof(cal_ransac_result), st));
            if (inlier_mask) SCUDA(cudaMemsetAsync(dmask.p, 0, (size_t)n_obs, st));
            for (const auto& [ng, ids] : by_size) {
                if (ng < 4) continue;
                const int64_t cnt = (int64_t)ids.size();
                DevBuf<int64_t> dids; DevBuf<double> g; DevBuf<cal_ransac_result> gres; DevBuf<uint8_t> gmask;
                SCUDA(dids.alloc(cnt)); SCUDA(g.alloc((size_t)4 * cnt * ng)); SCUDA(gres.alloc(cnt));
                if (inlier_mask) SCUDA(gmask.alloc((size_t)cnt * ng));
                SCUDA(cudaMemcpyAsync(dids.p, ids.data(), (size_t)cnt * sizeof(int64_t), cudaMemcpyHostToDevice, st));
                double *gx = g.p, *gy = gx + cnt * ng, *gu = gy + cnt * ng, *gv = gu + cnt * ng;
                const unsigned gb = (unsigned)std::min<int64_t>((cnt * ng + 255) / 256, 148 * 8);
                k_gather_views<<<gb, 256, 0, st>>>(cnt, (int)ng, dids.p, in.off.p, in.x.p, in.y.p, in.u.p, in.v.p, gx, gy, gu, gv);
                SCUDA(cudaStreamSynchronize(st));  // ids (a host vector) consumed; inputs in place before the default-stream launch
                if (cal_status rs = cal_ransac_homography_batch_dev(cnt, (int32_t)ng, gx, gy, gu, gv, ransac, 0, gres.p,
                                                                    inlier_mask ? gmask.p : nullptr, nullptr)) return rs;
                k_scatter_views<<<gb, 256, 0, st>>>(cnt, (int)ng, dids.p, in.off.p, gres.p, inlier_mask ? gmask.p : nullptr, dres.p,
                                                    inlier_mask ? dmask.p : nullptr);
                SCUDA(cudaStreamSynchronize(st));  // the group's buffers are released at the end of this iteration
            }
        }
        k_seed_from_ransac<<<(unsigned)((n_views + 127) / 128), 128, 0, st>>>(n_views, dres.p, dH.p, drms.p, drows.p, dsucc.p);
        if (inlier_mask) SCUDA(cudaMemcpyAsync(inlier_mask, dmask.p, (size_t)n_obs, cudaMemcpyDeviceToHost, st));
    } else {
        SeedArgs a{n_views, in.off.p, in.cam.p, in.x.p, in.y.p, in.u.p, in.v.p, nullptr, dH.p, drms.p, drows.p, dsucc.p, nullptr};
        k_view_dlt<false><<<(groups + 3) / 4, 128, 0, st>>>(a);
    }
    k_zhang_gram<<<n_cams, 256, 0, st>>>(n_views, in.cam.p, dsucc.p, drows.p, dgram.p);
    std::vector<double> gram(22 * (size_t)n_cams);
    SCUDA(cudaMemcpyAsync(gram.data(), dgram.p, gram.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    SCUDA(cudaStreamSynchronize(st));
    SCUDA(cudaGetLastError());
    std::vector<int32_t> camok(n_cams, 0);
    for (int c = 0; c < n_cams; ++c) {
        double* k5 = kmtx + 5 * c;
        for (int i = 0; i < 5; ++i) k5[i] = 0.0;
        const double* g = &gram[22 * (size_t)c];
        if (g[21] < 4.0) continue;  // "Zhang method requires at least 4 views" (zhang.cpp:152-155)
        // b = right singular vector of the smallest singular value of V = eigenvector of the smallest
        // eigenvalue of V^T V (the SAME minimiser of |V b| over |b| = 1 as zhang.cpp:193-196)
        double G[6][6], b[6];
        { int o = 0; for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { G[i][j] = G[j][i] = g[o]; ++o; } }
        smallest_eigvec6(G, b);
        if (!kmtx_from_conic(b, k5)) { for (int i = 0; i < 5; ++i) k5[i] = 0.0; continue; }
        if (opts && opts->use_bounds) sanitize(k5, *opts);
        camok[c] = 1;
    }
    for (int c = 0; c < n_cams; ++c) cam_success[c] = camok[c];
    SCUDA(cudaMemcpyAsync(dk.p, kmtx, sizeof(double) * 5 * n_cams, cudaMemcpyHostToDevice, st));
    SCUDA(cudaMemcpyAsync(dcamok.p, camok.data(), sizeof(int32_t) * n_cams, cudaMemcpyHostToDevice, st));
    if (poses) {
        k_pose_from_h<<<(unsigned)((n_views + 127) / 128), 128, 0, st>>>(n_views, in.cam.p, dsucc.p, dcamok.p, dk.p, dH.p, dposes.p);
        SCUDA(cudaMemcpyAsync(poses, dposes.p, sizeof(double) * 12 * n_views, cudaMemcpyDeviceToHost, st));
    }
    if (hmtx) SCUDA(cudaMemcpyAsync(hmtx, dH.p, sizeof(double) * 9 * n_views, cudaMemcpyDeviceToHost, st));
    if (sym_rms) SCUDA(cudaMemcpyAsync(sym_rms, drms.p, sizeof(double) * n_views, cudaMemcpyDeviceToHost, st));
    if (view_success) SCUDA(cudaMemcpyAsync(view_success, dsucc.p, sizeof(int32_t) * n_views, cudaMemcpyDeviceToHost, st));
    SCUDA(cudaStreamSynchronize(st));
    SCUDA(cudaGetLastError());
    return CAL_OK;
}

extern "C" cal_status cal_seed_intrinsics(int64_t n_views, const int64_t* view_offset, const int32_t* view_cam, const double* x,
                                          const double* y, const double* u, const double* v, int32_t n_cams,
                                          const cal_seed_options* opts, int device, double* kmtx, int32_t* cam_success,
                                          int32_t* view_success, double* hmtx, double* sym_rms, double* poses) {
    return cal_seed_intrinsics_ransac(n_views, view_offset, view_cam, x, y, u, v, n_cams, opts, nullptr, device, kmtx, cam_success,
                                      view_success, hmtx, sym_rms, poses, nullptr);
}

extern "C" cal_status cal_seed_planar_poses(int64_t n_views, const int64_t* view_offset, const int32_t* view_cam, const double* x,
                                            const double* y, const double* u, const double* v, int32_t n_cams, const double* kmtx,
                                            int device, double* poses, int32_t* view_success) {
    if (!x || !y || !u || !v || !kmtx || !poses) return sfail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    int64_t n_obs = 0;
    if (cal_status s = check_views(n_views, view_offset, view_cam, n_cams, &n_obs)) return s;
    if (cal_device_count() <= device) return sfail(CAL_ERR_CUDA, "no CUDA device: calib_b200 has no CPU fallback");
    SCUDA(cudaSetDevice(device));
    cudaStream_t st; SCUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
    SeedInputs in;
    SCUDA(in.off.init(view_offset, n_views + 1, st)); SCUDA(in.cam.init(view_cam, n_views, st));
    SCUDA(in.x.init(x, n_obs, st)); SCUDA(in.y.init(y, n_obs, st)); SCUDA(in.u.init(u, n_obs, st)); SCUDA(in.v.init(v, n_obs, st));
    DevBuf<double> dk, dposes; DevBuf<int32_t> dsucc;
    SCUDA(dk.alloc(5 * n_cams)); SCUDA(dposes.alloc(12 * n_views)); SCUDA(dsucc.alloc(n_views));
    SCUDA(cudaMemcpyAsync(dk.p, kmtx, sizeof(double) * 5 * n_cams, cudaMemcpyHostToDevice, st));
    SeedArgs a{n_views, in.off.p, in.cam.p, in.x.p, in.y.p, in.u.p, in.v.p, dk.p, nullptr, nullptr, nullptr, dsucc.p, dposes.p};
    const unsigned groups = (unsigned)((n_views + 31) / 32);
    k_view_dlt<true><<<(groups + 3) / 4, 128, 0, st>>>(a);
    SCUDA(cudaMemcpyAsync(poses, dposes.p, sizeof(double) * 12 * n_views, cudaMemcpyDeviceToHost, st));
    if (view_success) SCUDA(cudaMemcpyAsync(view_success, dsucc.p, sizeof(int32_t) * n_views, cudaMemcpyDeviceToHost, st));
    SCUDA(cudaStreamSynchronize(st));
    SCUDA(cudaGetLastError());
    return CAL_OK;
}
