// K3 generalised to the plane estimator: batched fit_plane_ransac
// (reference src/estimation/linear/planefit.cpp:86-104) = ransac<PlaneRansacEstimator>
// (include/calib/estimation/common/ransac.h:121-194) with the estimator of planefit.cpp:9-62 and
// fit_plane_svd (:66-84) as the refit.
//
// Same organisation as the homography kernel (ransac.cu): one warp per problem, the problem's points
// read from HBM once into shared memory, hypotheses processed in batches so that the O(n) work
// (scoring, the moment sums of the refit) is warp-cooperative and the O(1) work (three-point plane,
// the 3x3 eigenvector of the refit) runs one hypothesis per lane; the minimal-sample stream is the
// libstdc++ std::sample(.., 3, mt19937_64) stream generated on the device (ransac_sampler.cuh); the
// results are applied strictly in iteration order, statement by statement as ransac.h:147-190.
// Numerical differences to the reference are confined to rounding: the refit normal is the smallest
// eigenvector of the centred 3x3 scatter matrix (cyclic Jacobi) instead of the last right singular
// vector of the centred n x 3 matrix (Eigen::JacobiSVD).  That vector has no defined sign — the
// reference returns whatever JacobiSVD's V holds and its tests align signs before comparing
// (tests/unit/planefit_test.cpp:18-20); this kernel returns the sign that makes the normal component of
// largest magnitude positive.  Planes that come straight from a three-point sample keep the
// reference's orientation (v1 x v2).
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/calib_b200.h"
#include "dlt.cuh"
#include "plane_math.cuh"
#include "ransac_plane_kernel.cuh"
#include "ransac_sampler.cuh"

extern "C" void cal_set_last_error_(const char* msg);
extern "C" int cal_device_count(void);

namespace {

cal_status pfail(cal_status s, const std::string& m) { cal_set_last_error_(m.c_str()); return s; }
#define PCUDA(expr)                                                                                   \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) return pfail(CAL_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

cal_status launch_plane(int64_t n_problems, int n, const double* x, const double* y, const double* z, const cal_ransac_options& o,
                        int seed_per_problem, cal_plane_ransac_result* res, uint8_t* mask, cudaStream_t st, float* ms) {
    const size_t smem = kPlaneWarps * plane_per_warp_bytes(n);
    if (smem > 227 * 1024) return pfail(CAL_ERR_INVALID_ARGUMENT, "too many points per problem for the shared-memory plane RANSAC kernel (n <= ~2000)");
    PCUDA(cudaFuncSetAttribute(k_ransac_plane, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<int> table = build_niter_table(n, o, 3);
    int* dtable; PCUDA(cudaMalloc(reinterpret_cast<void**>(&dtable), table.size() * sizeof(int)));
    PCUDA(cudaMemcpyAsync(dtable, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    cudaEvent_t e0, e1; PCUDA(cudaEventCreate(&e0)); PCUDA(cudaEventCreate(&e1));
    PCUDA(cudaEventRecord(e0, st));
    const unsigned grid = (unsigned)((n_problems + kPlaneWarps - 1) / kPlaneWarps);
    k_ransac_plane<<<grid, 32 * kPlaneWarps, smem, st>>>(n_problems, n, x, y, z, o, seed_per_problem, dtable, res, mask);
    PCUDA(cudaEventRecord(e1, st));
    PCUDA(cudaEventSynchronize(e1));
    PCUDA(cudaGetLastError());
    float t = 0; PCUDA(cudaEventElapsedTime(&t, e0, e1));
    if (ms) *ms = t;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(dtable);
    return CAL_OK;
}

}  // namespace

extern "C" cal_status cal_ransac_plane_batch_dev(int64_t n_problems, int32_t n, const double* x_dev, const double* y_dev,
                                                 const double* z_dev, const cal_ransac_options* opts, int seed_per_problem,
                                                 cal_plane_ransac_result* results_dev, uint8_t* inlier_mask_dev, float* ms) {
    if (!opts || !results_dev || !x_dev || !y_dev || !z_dev || n_problems <= 0 || n <= 0) return pfail(CAL_ERR_INVALID_ARGUMENT, "bad argument");
    return launch_plane(n_problems, n, x_dev, y_dev, z_dev, *opts, seed_per_problem, results_dev, inlier_mask_dev, nullptr, ms);
}

extern "C" cal_status cal_ransac_plane_batch(int64_t n_problems, int32_t n, const double* x, const double* y, const double* z,
                                             const cal_ransac_options* opts, int seed_per_problem, int device,
                                             cal_plane_ransac_result* results, uint8_t* inlier_mask) {
    if (!opts || !results || !x || !y || !z || n_problems <= 0 || n <= 0) return pfail(CAL_ERR_INVALID_ARGUMENT, "bad argument");
    if (cal_device_count() <= device) return pfail(CAL_ERR_CUDA, "no CUDA device: calib_b200 has no CPU fallback");
    PCUDA(cudaSetDevice(device));
    const size_t nb = (size_t)n_problems * n * sizeof(double);
    double* dxyz = nullptr; cal_plane_ransac_result* dres = nullptr; uint8_t* dmask = nullptr;
    PCUDA(cudaMalloc(reinterpret_cast<void**>(&dxyz), 3 * nb));
    cal_status s = CAL_OK;
    auto done = [&](cal_status rc) { cudaFree(dxyz); cudaFree(dres); cudaFree(dmask); return rc; };
    if (cudaMalloc(reinterpret_cast<void**>(&dres), (size_t)n_problems * sizeof(cal_plane_ransac_result)) != cudaSuccess ||
        (inlier_mask && cudaMalloc(reinterpret_cast<void**>(&dmask), (size_t)n_problems * n) != cudaSuccess))
        return done(pfail(CAL_ERR_CUDA, "cudaMalloc failed"));
    double *dx = dxyz, *dy = dxyz + (size_t)n_problems * n, *dz = dy + (size_t)n_problems * n;
    if (cudaMemcpy(dx, x, nb, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(dy, y, nb, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(dz, z, nb, cudaMemcpyHostToDevice) != cudaSuccess)
        return done(pfail(CAL_ERR_CUDA, "host to device copy failed"));
    s = launch_plane(n_problems, n, dx, dy, dz, *opts, seed_per_problem, dres, dmask, nullptr, nullptr);
    if (s == CAL_OK) {
        if (cudaMemcpy(results, dres, (size_t)n_problems * sizeof(cal_plane_ransac_result), cudaMemcpyDeviceToHost) != cudaSuccess ||
            (inlier_mask && cudaMemcpy(inlier_mask, dmask, (size_t)n_problems * n, cudaMemcpyDeviceToHost) != cudaSuccess))
            s = pfail(CAL_ERR_CUDA, "device to host copy failed");
    }
    return done(s);
}
