// Device code of the batched RANSAC plane-fit kernel (see ransac_plane.cu for the description).  Kept in a header
// so that tests/host_emul can run this very source on the CPU under a lock-step SIMT shim (test-only).
#pragma once
#include "../../include/calib_b200.h"
#include "dlt.cuh"
#include "plane_math.cuh"
#include "ransac_sampler.cuh"

namespace {

constexpr int kPlaneWarps = 4;
constexpr int kPlaneBatch = 32;  // the per-hypothesis dense work is tiny: one hypothesis per lane, a full warp wide

struct PlaneMem {
    double *x, *y, *z;           // [n]
    unsigned long long* mt;      // [312]
    unsigned* cur;               // [kPlaneBatch][nw] inlier bit masks of the batch's hypotheses
    unsigned* ref;               // [kPlaneBatch][nw] after refit
    unsigned* best;              // [nw]
};

__host__ __device__ inline size_t plane_per_warp_bytes(int n) {
    const int nw = (n + 31) / 32;
    const size_t b = (size_t)3 * n * sizeof(double) + 312 * sizeof(unsigned long long) + (size_t)(1 + 2 * kPlaneBatch) * nw * sizeof(unsigned);
    return (b + 15) / 16 * 16;
}

__device__ __forceinline__ bool plane_from_sample(const PlaneMem& w, const int* s, double* P) {
    return calk::plane_from_points(w.x[s[0]], w.y[s[0]], w.z[s[0]], w.x[s[1]], w.y[s[1]], w.z[s[1]], w.x[s[2]], w.y[s[2]], w.z[s[2]], P);
}

// find_inliers (ransac.h:80-95) with PlaneRansacEstimator::residual (planefit.cpp:35-38).
// Writes the inlier bit mask, returns the count (warp-uniform).
__device__ __forceinline__ bool plane_inlier(const PlaneMem& w, int i, const double* P, double thresh) {
    return fabs(calk::plane_signed(P, w.x[i], w.y[i], w.z[i])) <= thresh;  // false for NaN
}
__device__ int plane_score(const PlaneMem& w, int n, int lane, const double* P, double thresh, unsigned* mask) {
    int cnt = 0;
    const int nfull = n & ~31;
    for (int base = 0; base < nfull; base += 32) {
        const unsigned bm = __ballot_sync(kFull, plane_inlier(w, base + lane, P, thresh));
        if (lane == 0) mask[base >> 5] = bm;
        cnt += __popc(bm);
    }
    if (nfull < n) {
        const int i = nfull + lane;
        const unsigned bm = __ballot_sync(kFull, i < n && plane_inlier(w, i < n ? i : n - 1, P, thresh));
        if (lane == 0) mask[nfull >> 5] = bm;
        cnt += __popc(bm);
    }
    __syncwarp();
    return cnt;
}
// sum of squared residuals over the inliers of `mask` (RansacResult::inlier_rms)
__device__ double plane_ssr(const PlaneMem& w, int n, int lane, const double* P, const unsigned* mask) {
    double s = 0.0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        if (i < n && ((mask[base >> 5] >> lane) & 1u)) {
            const double r = P[0] * w.x[i] + P[1] * w.y[i] + P[2] * w.z[i] + P[3];
            s = fma(r, r, s);
        }
    }
    return warp_sum(s);
}

// fit_plane_svd (planefit.cpp:66-84), O(n) part: centroid and the scatter matrix of the CENTRED inliers;
// the O(1) part is calk::plane_refit_solve (plane_math.cuh), lane-private.
using calk::PlaneSums;
__device__ void plane_sums(const PlaneMem& w, int n, int lane, const unsigned* mask, int cnt, PlaneSums& r) {
    double a0 = 0, a1 = 0, a2 = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        if (i < n && ((mask[base >> 5] >> lane) & 1u)) { a0 += w.x[i]; a1 += w.y[i]; a2 += w.z[i]; }
    }
    const double dn = (double)cnt;
    r.c[0] = warp_sum(a0) / dn; r.c[1] = warp_sum(a1) / dn; r.c[2] = warp_sum(a2) / dn;
    double m[6] = {0, 0, 0, 0, 0, 0};
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        if (i < n && ((mask[base >> 5] >> lane) & 1u)) {
            const double x = w.x[i] - r.c[0], y = w.y[i] - r.c[1], z = w.z[i] - r.c[2];
            m[0] = fma(x, x, m[0]); m[1] = fma(x, y, m[1]); m[2] = fma(x, z, m[2]);
            m[3] = fma(y, y, m[3]); m[4] = fma(y, z, m[4]); m[5] = fma(z, z, m[5]);
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) r.s[k] = warp_sum(m[k]);
}
__device__ __forceinline__ void bcast4(const double* mine, int src, double* out) {
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = __shfl_sync(kFull, mine[k], src);
}

__global__ void __launch_bounds__(32 * kPlaneWarps) k_ransac_plane(int64_t n_problems, int n, const double* __restrict__ gx,
                                                                   const double* __restrict__ gy, const double* __restrict__ gz,
                                                                   cal_ransac_options o, int seed_per_problem,
                                                                   const int* __restrict__ niter_table,
                                                                   cal_plane_ransac_result* __restrict__ results,
                                                                   uint8_t* __restrict__ gmask) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t prob = (int64_t)blockIdx.x * kPlaneWarps + wid;
    if (prob >= n_problems) return;
    const int nw = (n + 31) / 32;
    unsigned char* base = smem_raw + (size_t)wid * plane_per_warp_bytes(n);
    PlaneMem w;
    w.x = reinterpret_cast<double*>(base); w.y = w.x + n; w.z = w.y + n;
    w.mt = reinterpret_cast<unsigned long long*>(w.z + n);
    w.best = reinterpret_cast<unsigned*>(w.mt + 312); w.cur = w.best + nw; w.ref = w.cur + kPlaneBatch * nw;
    for (int i = lane; i < n; i += 32) { w.x[i] = gx[prob * n + i]; w.y[i] = gy[prob * n + i]; w.z[i] = gz[prob * n + i]; }
    for (int i = lane; i < nw; i += 32) w.best[i] = 0u;
    mt_seed(w.mt, o.seed + (seed_per_problem ? (unsigned long long)prob : 0ULL), lane);
    int pos = 312;
    bool has_best = false; int best_cnt = 0, best_iters = 0; double best_rms = INFINITY;
    double bestP[4] = {0, 0, 0, 0};  // PlaneRansacResult::plane = Zero (planefit.h:16)
    int dyn = o.max_iters, it = 0;
    while (n >= 3 && it < dyn) {
        const int B = min(kPlaneBatch, dyn - it);
        // ---- 1. minimal samples of the next B iterations (every iteration draws one, ransac.h:144-145) ----
        int my_idx[3] = {0, 0, 0};
        for (int h = 0; h < B; ++h) {
            int idx[3];
            sample_k<3>(w.mt, pos, lane, n, idx);
            if (lane == h) { my_idx[0] = idx[0]; my_idx[1] = idx[1]; my_idx[2] = idx[2]; }
        }
        // ---- 2. lane h: degeneracy test and three-point plane of hypothesis h ----
        double myP[4] = {0, 0, 0, 0};
        bool my_valid = false;
        if (lane < B) my_valid = plane_from_sample(w, my_idx, myP);
        __syncwarp();
        // ---- 3. score the hypotheses (warp-cooperative, one after the other) ----
        int my_cnt = 0;
        for (int h = 0; h < B; ++h) {
            if (!__shfl_sync(kFull, my_valid ? 1 : 0, h)) continue;
            double P[4]; bcast4(myP, h, P);
            const int cnt = plane_score(w, n, lane, P, o.thresh, w.cur + h * nw);
            if (lane == h) my_cnt = cnt;
        }
        // ---- 4.-6. refit on the inliers: sums cooperatively, eigenvector lane-private, rescore ----
        bool my_refit = false; int my_cnt2 = 0;
        double myP2[4] = {0, 0, 0, 0};
        if (o.refit_on_inliers) {
            PlaneSums mine{};
            for (int h = 0; h < B; ++h) {
                const int c = __shfl_sync(kFull, my_cnt, h);
                if (!__shfl_sync(kFull, my_valid ? 1 : 0, h) || c < o.min_inliers || c < 3) continue;
                PlaneSums r; plane_sums(w, n, lane, w.cur + h * nw, c, r);
                if (lane == h) mine = r;
            }
            if (lane < B && my_valid && my_cnt >= o.min_inliers && my_cnt >= 3) my_refit = calk::plane_refit_solve(mine, myP2);
            __syncwarp();
            for (int h = 0; h < B; ++h) {
                if (!__shfl_sync(kFull, my_refit ? 1 : 0, h)) continue;
                double P[4]; bcast4(myP2, h, P);
                const int cnt = plane_score(w, n, lane, P, o.thresh, w.ref + h * nw);
                if (lane == h) my_cnt2 = cnt;
            }
        }
        // ---- 7. apply the results in iteration order (ransac.h:147-190) ----
        for (int h = 0; h < B && it < dyn; ++h) {
            ++it;
            if (!__shfl_sync(kFull, my_valid ? 1 : 0, h)) continue;            // degenerate sample
            int cnt = __shfl_sync(kFull, my_cnt, h);
            if (cnt < o.min_inliers) continue;
            const bool refitted = __shfl_sync(kFull, my_refit ? 1 : 0, h) != 0;
            const unsigned* fin = w.cur + h * nw;
            double P[4];
            if (refitted) { bcast4(myP2, h, P); cnt = __shfl_sync(kFull, my_cnt2, h); fin = w.ref + h * nw; }
            else bcast4(myP, h, P);
            if (!has_best || cnt >= best_cnt) {  // is_better_model (ransac.h:113-117) needs the rms only on a tie or an improvement
                const double frms = cnt > 0 ? sqrt(plane_ssr(w, n, lane, P, fin) / (double)cnt) : INFINITY;
                if (!has_best || cnt > best_cnt || frms < best_rms) {
                    has_best = true; best_cnt = cnt; best_rms = frms; best_iters = it;
#pragma unroll
                    for (int k = 0; k < 4; ++k) bestP[k] = P[k];
                    for (int i = lane; i < nw; i += 32) w.best[i] = fin[i];
                    __syncwarp();
                }
            }
            // calculate_iterations (ransac.h:64-78) through the host-built table indexed by the inlier count
            const int niter = niter_table[cnt];
            int nd = niter == -1 ? o.max_iters : niter;
            if (nd < it) nd = it;
            if (nd > o.max_iters) nd = o.max_iters;
            dyn = nd;
        }
    }
    if (lane == 0) {
        cal_plane_ransac_result r;
        r.success = has_best ? 1 : 0; r.iters = best_iters; r.n_inliers = has_best ? best_cnt : 0; r.iters_run = it;
        for (int k = 0; k < 4; ++k) r.plane[k] = bestP[k];
        r.inlier_rms = best_rms; r.min_margin = 0.0;
        results[prob] = r;
    }
    if (gmask) for (int i = lane; i < n; i += 32) gmask[prob * n + i] = has_best ? (uint8_t)((w.best[i >> 5] >> (i & 31)) & 1u) : 0;
}

}  // namespace
