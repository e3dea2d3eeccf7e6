// Device code of the batched RANSAC homography kernel (see ransac.cu for the description).  Kept in a header so
// that tests/host_emul can run this very source on the CPU under a lock-step SIMT shim (test-only).
#pragma once
#include "../../include/calib_b200.h"
#include "dlt.cuh"
#include "ransac_sampler.cuh"

namespace {

constexpr int kWarpsPerCta = 4;

struct WarpMem {
    double *x, *y, *u, *v;       // [n]
    unsigned long long* mt;      // [312]
    unsigned* cur;               // [kBatch][nw] inlier bit masks of the batch's hypotheses
    unsigned* ref;               // [kBatch][nw] after refit
    unsigned* best;              // [nw]
};

// 4-point DLT (homographyestimator.cpp:45-78,123-143): null vector of the 8x9 matrix via Householder QR of A^T.
__device__ bool dlt4(const double* px, const double* py, const double* pu, const double* pv, double* H) {
    double scx = 0, scy = 0, dcx = 0, dcy = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { scx += px[i]; scy += py[i]; dcx += pu[i]; dcy += pv[i]; }
    scx *= 0.25; scy *= 0.25; dcx *= 0.25; dcy *= 0.25;
    double sm = 0, dm = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sm += sqrt((px[i] - scx) * (px[i] - scx) + (py[i] - scy) * (py[i] - scy));
        dm += sqrt((pu[i] - dcx) * (pu[i] - dcx) + (pv[i] - dcy) * (pv[i] - dcy));
    }
    sm *= 0.25; dm *= 0.25;
    const double ss = sm > 0 ? 1.4142135623730951 / sm : 1.0, ds = dm > 0 ? 1.4142135623730951 / dm : 1.0;
    // M = A^T (9 x 8), column c = row c of A
    double M[9][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double x = ss * (px[i] - scx), y = ss * (py[i] - scy), u = ds * (pu[i] - dcx), v = ds * (pv[i] - dcy);
        const int c0 = 2 * i, c1 = 2 * i + 1;
        M[0][c0] = -x; M[1][c0] = -y; M[2][c0] = -1.0; M[3][c0] = 0; M[4][c0] = 0; M[5][c0] = 0; M[6][c0] = u * x; M[7][c0] = u * y; M[8][c0] = u;
        M[0][c1] = 0; M[1][c1] = 0; M[2][c1] = 0; M[3][c1] = -x; M[4][c1] = -y; M[5][c1] = -1.0; M[6][c1] = v * x; M[7][c1] = v * y; M[8][c1] = v;
    }
    double beta[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double nrm2 = 0;
#pragma unroll
        for (int i = k; i < 9; ++i) nrm2 = fma(M[i][k], M[i][k], nrm2);
        const double nrm = sqrt(nrm2);
        const double alpha = M[k][k] > 0 ? -nrm : nrm;
        const double v0 = M[k][k] - alpha;
        double vtv = v0 * v0;
#pragma unroll
        for (int i = k + 1; i < 9; ++i) vtv = fma(M[i][k], M[i][k], vtv);
        beta[k] = vtv > 0 ? 2.0 / vtv : 0.0;
        M[k][k] = v0;  // column k now holds the Householder vector v_k (rows k..8)
#pragma unroll
        for (int j = k + 1; j < 8; ++j) {
            double d = 0;
#pragma unroll
            for (int i = k; i < 9; ++i) d = fma(M[i][k], M[i][j], d);
            d *= beta[k];
#pragma unroll
            for (int i = k; i < 9; ++i) M[i][j] = fma(-d, M[i][k], M[i][j]);
        }
    }
    // null vector = Q e9 = H_1 ... H_8 e9
    double z[9] = {0, 0, 0, 0, 0, 0, 0, 0, 1.0};
#pragma unroll
    for (int k = 7; k >= 0; --k) {
        double d = 0;
#pragma unroll
        for (int i = k; i < 9; ++i) d = fma(M[i][k], z[i], d);
        d *= beta[k];
#pragma unroll
        for (int i = k; i < 9; ++i) z[i] = fma(-d, M[i][k], z[i]);
    }
    double hn[9];
    const double ih = 1.0 / z[8];
#pragma unroll
    for (int i = 0; i < 9; ++i) hn[i] = z[i] * ih;
    denormalise(hn, ss, scx, scy, ds, dcx, dcy, H);
    return isfinite(H[0]);
}

// ---------------------------------------------------------------------------
// Hypotheses are processed in batches of up to kBatch: the O(n) work of a hypothesis (scoring, the
// monomial sums of the refit) is warp-cooperative, one hypothesis after the other, while the O(1)
// dense work (4-point DLT, the 9x9 null vector of the refit) runs lane-private, one hypothesis per
// lane — nothing is computed 32 times redundantly and the large unrolled blocks execute once per
// batch instead of once per hypothesis.  The minimal-sample stream depends only on (seed, n,
// iteration) (ransac.h:144-145), so drawing a batch ahead is exact; the results are then applied
// strictly in iteration order and the loop stops exactly where the sequential loop does (work on
// hypotheses past that point is discarded).
// ---------------------------------------------------------------------------
constexpr int kBatch = 16;

// find_inliers (ransac.h:80-95) with the symmetric transfer error of homographyestimator.cpp:80-93,
// written without divisions:  (e1 / qz^2 + e2 / pz^2) / 2 <= t^2  <=>  e1 pz^2 + e2 qz^2 <= 2 t^2 qz^2 pz^2.
// Writes the inlier bit mask, returns the count (warp-uniform).
__device__ __forceinline__ bool score_point(const WarpMem& w, int i, const double* H, const double* Hi, double two_thresh2) {
    const double x = w.x[i], y = w.y[i], u = w.u[i], v = w.v[i];
    const double qx = fma(H[0], x, fma(H[1], y, H[2])), qy = fma(H[3], x, fma(H[4], y, H[5])), qz = fma(H[6], x, fma(H[7], y, H[8]));
    const double px = fma(Hi[0], u, fma(Hi[1], v, Hi[2])), py = fma(Hi[3], u, fma(Hi[4], v, Hi[5])), pz = fma(Hi[6], u, fma(Hi[7], v, Hi[8]));
    const double a = fma(u, qz, -qx), b = fma(v, qz, -qy), c = fma(x, pz, -px), d = fma(y, pz, -py);
    const double e1 = fma(a, a, b * b), e2 = fma(c, c, d * d), qq = qz * qz, pp = pz * pz;
    return fma(e1, pp, e2 * qq) <= two_thresh2 * qq * pp && qq > 0.0 && pp > 0.0;  // false for NaN, like `r <= threshold`
}
__device__ int score_count(const WarpMem& w, int n, int lane, const double* H, double two_thresh2, unsigned* mask) {
    double Hi[9]; inv3(H, Hi);
    int cnt = 0;
    const int nfull = n & ~31;
    for (int base = 0; base < nfull; base += 32) {  // full rounds: no per-lane predicate
        const unsigned bm = __ballot_sync(kFull, score_point(w, base + lane, H, Hi, two_thresh2));
        if (lane == 0) mask[base >> 5] = bm;
        cnt += __popc(bm);
    }
    if (nfull < n) {
        const int i = nfull + lane;
        const unsigned bm = __ballot_sync(kFull, i < n && score_point(w, i < n ? i : n - 1, H, Hi, two_thresh2));
        if (lane == 0) mask[nfull >> 5] = bm;
        cnt += __popc(bm);
    }
    __syncwarp();
    return cnt;
}

// sum of squared residuals over the inliers of `mask` (for RansacResult::inlier_rms, ransac.h:97-111)
__device__ double inlier_ssr(const WarpMem& w, int n, int lane, const double* H, const unsigned* mask) {
    double Hi[9]; inv3(H, Hi);
    double s = 0.0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        if (i < n && ((mask[base >> 5] >> lane) & 1u)) {
            const double x = w.x[i], y = w.y[i], u = w.u[i], v = w.v[i];
            const double qx = H[0] * x + H[1] * y + H[2], qy = H[3] * x + H[4] * y + H[5], qz = H[6] * x + H[7] * y + H[8];
            const double iq = 1.0 / qz;
            const double du = u - qx * iq, dv = v - qy * iq;
            const double px = Hi[0] * u + Hi[1] * v + Hi[2], py = Hi[3] * u + Hi[4] * v + Hi[5], pz = Hi[6] * u + Hi[7] * v + Hi[8];
            const double ip = 1.0 / pz;
            const double dx = x - px * ip, dy = y - py * ip;
            s += 0.5 * (du * du + dv * dv + dx * dx + dy * dy);
        }
    }
    return warp_sum(s);
}

// refit, O(n) part: Hartley normalisation constants and the 24 monomial sums of A^T A over the inliers
// of `mask` (HomographyEstimator::refit, homographyestimator.cpp:146-166).  Warp-uniform results.
struct RefitSums { double scx, scy, dcx, dcy, ss, ds; double m[4][6]; };
__device__ void refit_sums(const WarpMem& w, int n, int lane, const unsigned* mask, int cnt, RefitSums& r) {
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        if (i < n && ((mask[base >> 5] >> lane) & 1u)) { a0 += w.x[i]; a1 += w.y[i]; a2 += w.u[i]; a3 += w.v[i]; }
    }
    const double inv = 1.0 / (double)cnt;
    r.scx = warp_sum(a0) * inv; r.scy = warp_sum(a1) * inv; r.dcx = warp_sum(a2) * inv; r.dcy = warp_sum(a3) * inv;
    // one more pass: mean distances and the monomial sums of the CENTRED coordinates; the Hartley
    // scales are powers of (ss, ds) per monomial and are applied to the sums afterwards
    a0 = a1 = 0;
    double m[4][6];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int e = 0; e < 6; ++e) m[k][e] = 0.0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        if (i < n && ((mask[base >> 5] >> lane) & 1u)) {
            const double x = w.x[i] - r.scx, y = w.y[i] - r.scy, u = w.u[i] - r.dcx, v = w.v[i] - r.dcy;
            a0 += sqrt(x * x + y * y); a1 += sqrt(u * u + v * v);
            const double pp[6] = {x * x, x * y, x, y * y, y, 1.0};
            const double wt[4] = {1.0, u, v, u * u + v * v};
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int e = 0; e < 6; ++e) m[k][e] = fma(wt[k], pp[e], m[k][e]);
        }
    }
    const double sm = warp_sum(a0) * inv, dm = warp_sum(a1) * inv;
    r.ss = sm > 0 ? 1.4142135623730951 / sm : 1.0; r.ds = dm > 0 ? 1.4142135623730951 / dm : 1.0;
    const double se[6] = {r.ss * r.ss, r.ss * r.ss, r.ss, r.ss * r.ss, r.ss, 1.0};
    const double sk[4] = {1.0, r.ds, r.ds, r.ds * r.ds};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int e = 0; e < 6; ++e) r.m[k][e] = warp_sum(m[k][e]) * (sk[k] * se[e]);
}
// refit, O(1) part (lane-private): null vector of A^T A, de-normalisation
__device__ bool refit_solve(const RefitSums& r, double* H) {
    double Gu[45];
    dlt_normal_matrix(r.m, Gu);
    double z[9];
    if (!smallest_eigvec9(Gu, z)) return false;
    double hn[9];
    const double ih = 1.0 / z[8];
#pragma unroll
    for (int i = 0; i < 9; ++i) hn[i] = z[i] * ih;
    denormalise(hn, r.ss, r.scx, r.scy, r.ds, r.dcx, r.dcy, H);
    return isfinite(H[0]);
}

__device__ __forceinline__ void bcast9(const double* mine, int src, double* out) {
#pragma unroll
    for (int k = 0; k < 9; ++k) out[k] = __shfl_sync(kFull, mine[k], src);
}

__global__ void __launch_bounds__(32 * kWarpsPerCta) k_ransac(int64_t n_problems, int n, const double* __restrict__ gx,
                                                              const double* __restrict__ gy, const double* __restrict__ gu,
                                                              const double* __restrict__ gv, cal_ransac_options o,
                                                              int seed_per_problem, const int* __restrict__ niter_table,
                                                              cal_ransac_result* __restrict__ results, uint8_t* __restrict__ gmask) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t prob = (int64_t)blockIdx.x * kWarpsPerCta + wid;
    if (prob >= n_problems) return;
    const int nw = (n + 31) / 32;
    const size_t per_warp = (size_t)4 * n * sizeof(double) + 312 * sizeof(unsigned long long) + (size_t)(1 + 2 * kBatch) * nw * sizeof(unsigned);
    unsigned char* base = smem_raw + (size_t)wid * ((per_warp + 15) / 16 * 16);
    WarpMem w;
    w.x = reinterpret_cast<double*>(base); w.y = w.x + n; w.u = w.y + n; w.v = w.u + n;
    w.mt = reinterpret_cast<unsigned long long*>(w.v + n);
    w.best = reinterpret_cast<unsigned*>(w.mt + 312); w.cur = w.best + nw; w.ref = w.cur + kBatch * nw;  // cur / ref: [kBatch][nw]
    for (int i = lane; i < n; i += 32) {
        w.x[i] = gx[prob * n + i]; w.y[i] = gy[prob * n + i]; w.u[i] = gu[prob * n + i]; w.v[i] = gv[prob * n + i];
    }
    for (int i = lane; i < nw; i += 32) w.best[i] = 0u;
    mt_seed(w.mt, o.seed + (seed_per_problem ? (unsigned long long)prob : 0ULL), lane);
    int pos = 312;
    const double two_thresh2 = 2.0 * o.thresh * o.thresh;
    bool has_best = false; int best_cnt = 0, best_iters = 0; double best_rms = INFINITY;
    double bestH[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    int dyn = o.max_iters, it = 0;
    while (n >= 4 && it < dyn) {
        const int B = min(kBatch, dyn - it);
        // ---- 1. minimal samples of the next B iterations (every iteration draws one, ransac.h:144-145) ----
        int my_idx[4] = {0, 0, 0, 0};
        for (int h = 0; h < B; ++h) {
            int idx[4];
            sample_k<4>(w.mt, pos, lane, n, idx);
            if (lane == h) { my_idx[0] = idx[0]; my_idx[1] = idx[1]; my_idx[2] = idx[2]; my_idx[3] = idx[3]; }
        }
        // ---- 2. lane h: degeneracy test and 4-point DLT of hypothesis h ----
        double myH[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        bool my_valid = false;
        if (lane < B) {
            double px[4], py[4], pu[4], pv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { px[k] = w.x[my_idx[k]]; py[k] = w.y[my_idx[k]]; pu[k] = w.u[my_idx[k]]; pv[k] = w.v[my_idx[k]]; }
            // has_near_collinear_triplet (homographyestimator.cpp:100-119), object coordinates
            bool degen = false;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = a + 1; b < 4; ++b)
#pragma unroll
                    for (int c = b + 1; c < 4; ++c)
                        degen |= fabs((px[b] - px[a]) * (py[c] - py[a]) - (py[b] - py[a]) * (px[c] - px[a])) < 1e-6;
            if (!degen) my_valid = dlt4(px, py, pu, pv, myH);
        }
        __syncwarp();
        // ---- 3. score the hypotheses (warp-cooperative, one after the other) ----
        int my_cnt = 0;
        for (int h = 0; h < B; ++h) {
            if (!__shfl_sync(kFull, my_valid ? 1 : 0, h)) continue;
            double H[9]; bcast9(myH, h, H);
            const int cnt = score_count(w, n, lane, H, two_thresh2, w.cur + h * nw);
            if (lane == h) my_cnt = cnt;
        }
        // ---- 4.-6. refit on the inliers: sums cooperatively, 9x9 null vector lane-private, rescore ----
        bool my_refit = false; int my_cnt2 = 0;
        double myH2[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (o.refit_on_inliers) {
            RefitSums mine{};
            for (int h = 0; h < B; ++h) {
                const int c = __shfl_sync(kFull, my_cnt, h);
                if (!__shfl_sync(kFull, my_valid ? 1 : 0, h) || c < o.min_inliers || c < 4) continue;
                RefitSums r; refit_sums(w, n, lane, w.cur + h * nw, c, r);
                if (lane == h) mine = r;
            }
            if (lane < B && my_valid && my_cnt >= o.min_inliers && my_cnt >= 4) my_refit = refit_solve(mine, myH2);
            __syncwarp();
            for (int h = 0; h < B; ++h) {
                if (!__shfl_sync(kFull, my_refit ? 1 : 0, h)) continue;
                double H[9]; bcast9(myH2, h, H);
                const int cnt = score_count(w, n, lane, H, two_thresh2, w.ref + h * nw);
                if (lane == h) my_cnt2 = cnt;
            }
        }
        // ---- 7. apply the results in iteration order (ransac.h:147-190) ----
        for (int h = 0; h < B && it < dyn; ++h) {
            ++it;
            if (!__shfl_sync(kFull, my_valid ? 1 : 0, h)) continue;            // degenerate sample or failed fit
            int cnt = __shfl_sync(kFull, my_cnt, h);
            if (cnt < o.min_inliers) continue;
            const bool refitted = __shfl_sync(kFull, my_refit ? 1 : 0, h) != 0;
            const unsigned* fin = w.cur + h * nw;
            double H[9];
            if (refitted) { bcast9(myH2, h, H); cnt = __shfl_sync(kFull, my_cnt2, h); fin = w.ref + h * nw; }
            else bcast9(myH, h, H);
            if (!has_best || cnt >= best_cnt) {  // is_better_model (ransac.h:113-117) needs the rms only on a tie or an improvement
                const double frms = cnt > 0 ? sqrt(inlier_ssr(w, n, lane, H, fin) / (double)cnt) : INFINITY;
                if (!has_best || cnt > best_cnt || frms < best_rms) {
                    has_best = true; best_cnt = cnt; best_rms = frms; best_iters = it;
#pragma unroll
                    for (int k = 0; k < 9; ++k) bestH[k] = H[k];
                    for (int i = lane; i < nw; i += 32) w.best[i] = fin[i];
                    __syncwarp();
                }
            }
            // calculate_iterations (ransac.h:64-78) through the host-built table indexed by the inlier count
            const int niter = niter_table[cnt];
            int nd = niter == -1 ? o.max_iters : niter;  // -1: the function returns max_iters before the clamp
            if (nd < it) nd = it;
            if (nd > o.max_iters) nd = o.max_iters;
            dyn = nd;
        }
    }
    // symmetric_rms_px (optim/homography.cpp:18-28): sqrt(sum of the (root) residuals / (2 n_inl))
    double sym = INFINITY;
    if (has_best && best_cnt > 0) {
        double Hi[9]; inv3(bestH, Hi);
        double s = 0.0;
        for (int basei = 0; basei < n; basei += 32) {
            const int i = basei + lane;
            if (i < n && ((w.best[basei >> 5] >> lane) & 1u)) {
                const double x = w.x[i], y = w.y[i], u = w.u[i], v = w.v[i];
                const double qz = bestH[6] * x + bestH[7] * y + bestH[8];
                const double du = u - (bestH[0] * x + bestH[1] * y + bestH[2]) / qz, dv = v - (bestH[3] * x + bestH[4] * y + bestH[5]) / qz;
                const double pz = Hi[6] * u + Hi[7] * v + Hi[8];
                const double dx = x - (Hi[0] * u + Hi[1] * v + Hi[2]) / pz, dy = y - (Hi[3] * u + Hi[4] * v + Hi[5]) / pz;
                s += sqrt(0.5 * (du * du + dv * dv + dx * dx + dy * dy));
            }
        }
        sym = sqrt(warp_sum(s) / (2.0 * (double)best_cnt));
    }
    if (lane == 0) {
        cal_ransac_result r;
        r.success = has_best ? 1 : 0; r.iters = best_iters; r.n_inliers = has_best ? best_cnt : 0; r.iters_run = it;
        for (int k = 0; k < 9; ++k) r.hmtx[k] = bestH[k];
        r.inlier_rms = best_rms; r.symmetric_rms_px = has_best ? sym : 0.0; r.min_margin = 0.0;
        results[prob] = r;
    }
    if (gmask) for (int i = lane; i < n; i += 32) gmask[prob * n + i] = has_best ? (uint8_t)((w.best[i >> 5] >> (i & 31)) & 1u) : 0;
}

}  // namespace
