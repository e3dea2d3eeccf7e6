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
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
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

// device memory / events of one call, released on every exit
struct DevBuf { void* p = nullptr; ~DevBuf() { if (p) cudaFree(p); } template <class T> T* as() const { return static_cast<T*>(p); } };
struct Ev { cudaEvent_t e = nullptr; ~Ev() { if (e) cudaEventDestroy(e); } };
// stream-ordered buffer from the device's default memory pool, returned to it in stream order
struct PoolBuf { void* p = nullptr; cudaStream_t st = nullptr; ~PoolBuf() { if (p) cudaFreeAsync(p, st); } template <class T> T* as() const { return static_cast<T*>(p); } };
struct Stream { cudaStream_t s = nullptr; ~Stream() { if (s) cudaStreamDestroy(s); } };

// the adaptive-iteration table of calculate_iterations for this (n, options): built once per call, on the device
cal_status upload_niter_table(int n, const cal_ransac_options& o, DevBuf& dtable, cudaStream_t st) {
    const std::vector<int> table = build_niter_table(n, o, 4);
    RCUDA(cudaMalloc(&dtable.p, table.size() * sizeof(int)));
    RCUDA(cudaMemcpyAsync(dtable.p, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    RCUDA(cudaStreamSynchronize(st));   // the host vector goes out of scope
    return CAL_OK;
}

cal_status prepare_kernel(int n, size_t* smem_out) {
    const size_t smem = smem_per_cta(n);
    if (smem > 227 * 1024) return rfail(CAL_ERR_INVALID_ARGUMENT, "too many correspondences per problem for the shared-memory RANSAC kernel (n <= ~1700)");
    RCUDA(cudaFuncSetAttribute(k_ransac, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    *smem_out = smem;
    return CAL_OK;
}

// queue the kernel for problems [0, n_problems) of the given arrays; `o.seed` is the seed of problem 0
void enqueue(int64_t n_problems, int n, const double* x, const double* y, const double* u, const double* v, const cal_ransac_options& o,
             int seed_per_problem, const int* dtable, size_t smem, cal_ransac_result* res, uint8_t* mask, cudaStream_t st) {
    const unsigned grid = (unsigned)((n_problems + kWarpsPerCta - 1) / kWarpsPerCta);
    k_ransac<<<grid, 32 * kWarpsPerCta, smem, st>>>(n_problems, n, x, y, u, v, o, seed_per_problem, dtable, res, mask);
}

cal_status launch(int64_t n_problems, int n, const double* x, const double* y, const double* u, const double* v,
                  const cal_ransac_options& o, int seed_per_problem, cal_ransac_result* res, uint8_t* mask, cudaStream_t st,
                  float* ms) {
    size_t smem = 0;
    if (cal_status s = prepare_kernel(n, &smem)) return s;
    DevBuf dtable;
    if (cal_status s = upload_niter_table(n, o, dtable, st)) return s;
    Ev e0, e1; RCUDA(cudaEventCreate(&e0.e)); RCUDA(cudaEventCreate(&e1.e));
    RCUDA(cudaEventRecord(e0.e, st));
    enqueue(n_problems, n, x, y, u, v, o, seed_per_problem, dtable.as<int>(), smem, res, mask, st);
    RCUDA(cudaEventRecord(e1.e, st));
    RCUDA(cudaEventSynchronize(e1.e));
    RCUDA(cudaGetLastError());
    float t = 0; RCUDA(cudaEventElapsedTime(&t, e0.e, e1.e));
    if (ms) *ms = t;
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

// Host arrays in, host results out.  The problems are cut into chunks that travel through three streams — upload of
// chunk k + 1, kernel of chunk k, download of chunk k - 1 overlap (the arrays should be page-locked for that: a copy
// from pageable memory is staged by the driver and serialises) — so the call takes about max(PCIe time, kernel time),
// not their sum.  Seeds follow the problem index, so the chunking does not change any result.
extern "C" cal_status cal_ransac_homography_batch(int64_t n_problems, int32_t n, const double* x, const double* y,
                                                  const double* u, const double* v, const cal_ransac_options* opts,
                                                  int seed_per_problem, int device, cal_ransac_result* results,
                                                  uint8_t* inlier_mask) {
    if (!opts || !results || !x || !y || !u || !v || n_problems <= 0 || n <= 0) return rfail(CAL_ERR_INVALID_ARGUMENT, "bad argument");
    if (cal_device_count() <= device) return rfail(CAL_ERR_CUDA, "no CUDA device: calib_b200 has no CPU fallback");
    RCUDA(cudaSetDevice(device));
    size_t smem = 0;
    if (cal_status s = prepare_kernel(n, &smem)) return s;
    const size_t ne = (size_t)n_problems * n;
    Stream up, run, down;   // (declared before the buffers: the buffers are released in stream order on `run` first)
    RCUDA(cudaStreamCreateWithFlags(&up.s, cudaStreamNonBlocking)); RCUDA(cudaStreamCreateWithFlags(&run.s, cudaStreamNonBlocking));
    RCUDA(cudaStreamCreateWithFlags(&down.s, cudaStreamNonBlocking));
    {   // the gigabytes of a call come from the device's memory pool and go back to it: repeated calls reuse the same physical memory
        cudaMemPool_t pool; uint64_t thr = UINT64_MAX;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    PoolBuf dxyuv{nullptr, run.s}, dres{nullptr, run.s}, dmask{nullptr, run.s};
    DevBuf dtable;
    RCUDA(cudaMallocAsync(&dxyuv.p, 4 * ne * sizeof(double), run.s));
    RCUDA(cudaMallocAsync(&dres.p, (size_t)n_problems * sizeof(cal_ransac_result), run.s));
    if (inlier_mask) RCUDA(cudaMallocAsync(&dmask.p, ne, run.s));
    Ev ready; RCUDA(cudaEventCreateWithFlags(&ready.e, cudaEventDisableTiming)); RCUDA(cudaEventRecord(ready.e, run.s));
    RCUDA(cudaStreamWaitEvent(up.s, ready.e, 0));   // the upload stream may touch the buffers once they exist
    if (cal_status s = upload_niter_table(n, *opts, dtable, run.s)) return s;
    double* dx = dxyuv.as<double>(); double *dy = dx + ne, *du = dy + ne, *dv = du + ne;
    // few, large chunks: a problem's iteration count is data dependent, so a launch of a few thousand problems is
    // dominated by its slowest ones (measured: 16 chunks of 6 250 problems took 3.7x the time of one launch)
    int64_t chunk = 25000;
    if (const char* e = getenv("CALIB_B200_RANSAC_CHUNK")) chunk = std::max<int64_t>(1, atoll(e));   // tests: exercise the pipeline at small sizes
    const int64_t n_chunks = std::min<int64_t>(8, std::max<int64_t>(1, n_problems / chunk));
    const int64_t per = (n_problems + n_chunks - 1) / n_chunks;
    std::vector<Ev> uploaded((size_t)n_chunks), computed((size_t)n_chunks);
    for (int64_t k = 0; k < n_chunks; ++k) {   // uploads and kernels are queued first (nothing here blocks the host) ...
        const int64_t p0 = k * per, p1 = std::min(n_problems, p0 + per);
        if (p0 >= p1) break;
        const size_t off = (size_t)p0 * n, cnt = (size_t)(p1 - p0) * n * sizeof(double);
        RCUDA(cudaMemcpyAsync(dx + off, x + off, cnt, cudaMemcpyHostToDevice, up.s)); RCUDA(cudaMemcpyAsync(dy + off, y + off, cnt, cudaMemcpyHostToDevice, up.s));
        RCUDA(cudaMemcpyAsync(du + off, u + off, cnt, cudaMemcpyHostToDevice, up.s)); RCUDA(cudaMemcpyAsync(dv + off, v + off, cnt, cudaMemcpyHostToDevice, up.s));
        RCUDA(cudaEventCreateWithFlags(&uploaded[k].e, cudaEventDisableTiming)); RCUDA(cudaEventRecord(uploaded[k].e, up.s));
        RCUDA(cudaStreamWaitEvent(run.s, uploaded[k].e, 0));
        cal_ransac_options o = *opts;
        if (seed_per_problem) o.seed += (uint64_t)p0;
        enqueue(p1 - p0, n, dx + off, dy + off, du + off, dv + off, o, seed_per_problem, dtable.as<int>(), smem, dres.as<cal_ransac_result>() + p0,
                inlier_mask ? dmask.as<uint8_t>() + off : nullptr, run.s);
        RCUDA(cudaEventCreateWithFlags(&computed[k].e, cudaEventDisableTiming)); RCUDA(cudaEventRecord(computed[k].e, run.s));
    }
    for (int64_t k = 0; k < n_chunks; ++k) {   // ... then the downloads, chunk by chunk behind their kernels (a copy into pageable
        const int64_t p0 = k * per, p1 = std::min(n_problems, p0 + per);   // memory blocks the host until its kernel is done)
        if (p0 >= p1) break;
        const size_t off = (size_t)p0 * n;
        RCUDA(cudaStreamWaitEvent(down.s, computed[k].e, 0));
        RCUDA(cudaMemcpyAsync(results + p0, dres.as<cal_ransac_result>() + p0, (size_t)(p1 - p0) * sizeof(cal_ransac_result), cudaMemcpyDeviceToHost, down.s));
        if (inlier_mask) RCUDA(cudaMemcpyAsync(inlier_mask + off, dmask.as<uint8_t>() + off, (size_t)(p1 - p0) * n, cudaMemcpyDeviceToHost, down.s));
    }
    RCUDA(cudaStreamSynchronize(down.s));
    RCUDA(cudaStreamSynchronize(up.s));
    RCUDA(cudaStreamSynchronize(run.s));
    RCUDA(cudaGetLastError());
    return CAL_OK;
}

// Independent problems split over several devices of the box, no communication at all: device d takes a
// contiguous slice, its seeds continue the global problem index, one host thread per device.  devices may repeat.
extern "C" cal_status cal_ransac_homography_batch_multi(int64_t n_problems, int32_t n, const double* x, const double* y,
                                                        const double* u, const double* v, const cal_ransac_options* opts,
                                                        int seed_per_problem, int32_t n_devices, const int32_t* devices,
                                                        cal_ransac_result* results, uint8_t* inlier_mask) {
    if (!opts || !results || !x || !y || !u || !v || n_problems <= 0 || n <= 0 || n_devices <= 0 || !devices) return rfail(CAL_ERR_INVALID_ARGUMENT, "bad argument");
    const int nd = (int)std::min<int64_t>(n_devices, n_problems);
    const int64_t per = (n_problems + nd - 1) / nd;
    std::vector<cal_status> rc((size_t)nd, CAL_OK);
    std::vector<std::string> msg((size_t)nd);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; ++d)
        th.emplace_back([&, d] {
            const int64_t p0 = d * per, p1 = std::min(n_problems, p0 + per);
            if (p0 >= p1) return;
            const size_t off = (size_t)p0 * n;
            cal_ransac_options o = *opts;
            if (seed_per_problem) o.seed += (uint64_t)p0;
            rc[d] = cal_ransac_homography_batch(p1 - p0, n, x + off, y + off, u + off, v + off, &o, seed_per_problem, devices[d], results + p0,
                                                inlier_mask ? inlier_mask + off : nullptr);
            if (rc[d] != CAL_OK) msg[d] = cal_last_error();   // (thread-local in the callee's thread)
        });
    for (auto& t : th) t.join();
    for (int d = 0; d < nd; ++d) if (rc[d] != CAL_OK) return rfail(rc[d], msg[d]);
    return CAL_OK;
}
