// K3: batched RANSAC homography — estimate_homography(view, RansacOptions)
// (reference include/calib/estimation/linear/homography.h:22-24) =
// ransac<HomographyEstimator> (include/calib/estimation/common/ransac.h:121-194)
// with the estimator of src/estimation/linear/homographyestimator.cpp:16-174 and
// the wrapper / symmetric_rms_px of src/estimation/optim/homography.cpp:18-73.
//
// One warp per problem.  The problem's correspondences are read from HBM once
// (coalesced) into shared memory and every hypothesis is scored from there;
// hypotheses are processed in batches (see k_ransac) so that the small dense
// solves run one hypothesis per lane instead of 32 times redundantly.
// The reference's sequential semantics are replayed exactly per problem:
//   * the minimal-sample stream is the libstdc++ std::sample / std::mt19937_64
//     stream (ransac.h:135,144-145), generated ON THE DEVICE: the 312-word
//     twist and the tempering run lane-parallel, the selection-sampling walk
//     takes 32 (index, index+1) pairs per step with one Lemire multiply per
//     lane (integer arithmetic, bit-exact);
//   * degeneracy test, 4-point Hartley-normalised DLT, inlier scoring, refit
//     on the inliers, best-model update and the adaptive iteration bound
//     follow ransac.h:140-190 statement by statement.
// Numerical differences to the reference are confined to rounding: the DLT
// null vector comes from Householder QR of A^T (4-point) / inverse iteration on
// the 9x9 normal matrix (refit) instead of Eigen::JacobiSVD, and the inlier
// test compares squared errors.  Inlier sets are therefore identical whenever
// no residual lies within rounding distance of the threshold.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/calib_b200.h"
#include "dlt.cuh"
#include "ransac_kernel.cuh"
#include "ransac_sampler.cuh"

extern "C" void cal_set_last_error_(const char* msg);

namespace {

cal_status rfail(cal_status s, const std::string& m) { cal_set_last_error_(m.c_str()); return s; }
#define RCUDA(expr)                                                                                   \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) return rfail(CAL_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

size_t smem_per_cta(int n) {
    const int nw = (n + 31) / 32;
    const size_t per_warp = (size_t)4 * n * sizeof(double) + 312 * sizeof(unsigned long long) + (size_t)(1 + 2 * kBatch) * nw * sizeof(unsigned);
    return kWarpsPerCta * ((per_warp + 15) / 16 * 16);
}

cal_status launch(int64_t n_problems, int n, const double* x, const double* y, const double* u, const double* v,
                  const cal_ransac_options& o, int seed_per_problem, cal_ransac_result* res, uint8_t* mask, cudaStream_t st,
                  float* ms) {
    const size_t smem = smem_per_cta(n);
    if (smem > 227 * 1024) return rfail(CAL_ERR_INVALID_ARGUMENT, "too many correspondences per problem for the shared-memory RANSAC kernel (n <= ~1700)");
    RCUDA(cudaFuncSetAttribute(k_ransac, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<int> table = build_niter_table(n, o, 4);
    int* dtable; RCUDA(cudaMalloc(reinterpret_cast<void**>(&dtable), table.size() * sizeof(int)));
    RCUDA(cudaMemcpyAsync(dtable, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    cudaEvent_t e0, e1; RCUDA(cudaEventCreate(&e0)); RCUDA(cudaEventCreate(&e1));
    RCUDA(cudaEventRecord(e0, st));
    const unsigned grid = (unsigned)((n_problems + kWarpsPerCta - 1) / kWarpsPerCta);
    k_ransac<<<grid, 32 * kWarpsPerCta, smem, st>>>(n_problems, n, x, y, u, v, o, seed_per_problem, dtable, res, mask);
    RCUDA(cudaEventRecord(e1, st));
    RCUDA(cudaEventSynchronize(e1));
    RCUDA(cudaGetLastError());
    float t = 0; RCUDA(cudaEventElapsedTime(&t, e0, e1));
    if (ms) *ms = t;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(dtable);
    return CAL_OK;
}

}  // namespace

extern "C" int cal_device_count(void);

extern "C" cal_status cal_ransac_homography_batch_dev(int64_t n_problems, int32_t n, const double* x_dev, const double* y_dev,
                                                      const double* u_dev, const double* v_dev, const cal_ransac_options* opts,
                                                      int seed_per_problem, cal_ransac_result* results_dev,
                                                      uint8_t* inlier_mask_dev, float* ms) {
    if (!opts || !results_dev || n_problems <= 0 || n <= 0) return rfail(CAL_ERR_INVALID_ARGUMENT, "bad argument");
    return launch(n_problems, n, x_dev, y_dev, u_dev, v_dev, *opts, seed_per_problem, results_dev, inlier_mask_dev, nullptr, ms);
}

extern "C" cal_status cal_ransac_homography_batch(int64_t n_problems, int32_t n, const double* x, const double* y,
                                                  const double* u, const double* v, const cal_ransac_options* opts,
                                                  int seed_per_problem, int device, cal_ransac_result* results,
                                                  uint8_t* inlier_mask) {
    if (!opts || !results || !x || !y || !u || !v || n_problems <= 0 || n <= 0) return rfail(CAL_ERR_INVALID_ARGUMENT, "bad argument");
    if (cal_device_count() <= device) return rfail(CAL_ERR_CUDA, "no CUDA device: calib_b200 has no CPU fallback");
    RCUDA(cudaSetDevice(device));
    const size_t nb = (size_t)n_problems * n * sizeof(double);
    // every exit releases the device buffers (cudaFree(nullptr) is a no-op)
    double* dxyuv = nullptr; cal_ransac_result* dres = nullptr; uint8_t* dmask = nullptr;
    auto done = [&](cal_status rc) { cudaFree(dxyuv); cudaFree(dres); cudaFree(dmask); return rc; };
    if (cudaMalloc(reinterpret_cast<void**>(&dxyuv), 4 * nb) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&dres), (size_t)n_problems * sizeof(cal_ransac_result)) != cudaSuccess ||
        (inlier_mask && cudaMalloc(reinterpret_cast<void**>(&dmask), (size_t)n_problems * n) != cudaSuccess))
        return done(rfail(CAL_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(cudaGetLastError())));
    const size_t ne = (size_t)n_problems * n;
    double *dx = dxyuv, *dy = dx + ne, *du = dy + ne, *dv = du + ne;
    if (cudaMemcpy(dx, x, nb, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(dy, y, nb, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(du, u, nb, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(dv, v, nb, cudaMemcpyHostToDevice) != cudaSuccess)
        return done(rfail(CAL_ERR_CUDA, std::string("host to device copy: ") + cudaGetErrorString(cudaGetLastError())));
    cal_status s = launch(n_problems, n, dx, dy, du, dv, *opts, seed_per_problem, dres, dmask, nullptr, nullptr);
    if (s == CAL_OK) {
        if (cudaMemcpy(results, dres, (size_t)n_problems * sizeof(cal_ransac_result), cudaMemcpyDeviceToHost) != cudaSuccess ||
            (inlier_mask && cudaMemcpy(inlier_mask, dmask, (size_t)n_problems * n, cudaMemcpyDeviceToHost) != cudaSuccess))
            s = rfail(CAL_ERR_CUDA, std::string("device to host copy: ") + cudaGetErrorString(cudaGetLastError()));
    }
    return done(s);
}
