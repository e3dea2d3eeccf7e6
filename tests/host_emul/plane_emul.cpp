// TEST INFRASTRUCTURE — CPU replay of the plane RANSAC kernel's per-problem logic, compiled with g++ (no GPU).
//
// It runs the PRODUCT's own per-hypothesis arithmetic (calibration_b200/csrc/plane_math.cuh: the three-point
// plane, the signed distance, the Jacobi eigenvector refit and its sign convention) and the PRODUCT's own
// iteration-bound table (ransac_iters.hpp) in the order k_ransac_plane applies them — batches of 32
// hypotheses drawn ahead, results applied in iteration order, work past the adaptive bound discarded — on a
// minimal-sample list the caller supplies (the oracle's std::sample stream).  Only the warp decomposition
// (ballots, shuffles, lane-partitioned sums) is not exercised here.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../calibration_b200/csrc/plane_math.cuh"
#include "../../calibration_b200/csrc/ransac_iters.hpp"

namespace {

int score(int n, const double* x, const double* y, const double* z, const double* P, double thresh, std::vector<uint8_t>& m) {
    int cnt = 0;
    for (int i = 0; i < n; ++i) { m[i] = std::fabs(calk::plane_signed(P, x[i], y[i], z[i])) <= thresh; cnt += m[i]; }
    return cnt;
}

}  // namespace

extern "C" int emul_ransac_plane(int32_t n, const double* x, const double* y, const double* z, const cal_ransac_options* op,
                                 const int32_t* sample_idx /* [max_iters][3] */, cal_plane_ransac_result* res, uint8_t* mask) {
    const cal_ransac_options& o = *op;
    constexpr int kBatch = 32;
    const std::vector<int> table = build_niter_table(n, o, 3);
    bool has_best = false; int best_cnt = 0, best_iters = 0; double best_rms = INFINITY;
    double bestP[4] = {0, 0, 0, 0};
    std::vector<uint8_t> best(n, 0);
    int dyn = o.max_iters, it = 0;
    struct Hyp { bool valid = false, refit = false; int cnt = 0, cnt2 = 0; double P[4] = {0, 0, 0, 0}, P2[4] = {0, 0, 0, 0}; std::vector<uint8_t> cur, ref; };
    while (n >= 3 && it < dyn) {
        const int B = std::min(kBatch, dyn - it);
        std::vector<Hyp> hyp(B);
        for (int h = 0; h < B; ++h) {
            Hyp& H = hyp[h];
            const int32_t* s = sample_idx + 3 * (it + h);
            H.cur.assign(n, 0); H.ref.assign(n, 0);
            H.valid = calk::plane_from_points(x[s[0]], y[s[0]], z[s[0]], x[s[1]], y[s[1]], z[s[1]], x[s[2]], y[s[2]], z[s[2]], H.P);
            if (!H.valid) continue;
            H.cnt = score(n, x, y, z, H.P, o.thresh, H.cur);
            if (!o.refit_on_inliers || H.cnt < o.min_inliers || H.cnt < 3) continue;
            calk::PlaneSums r{};
            for (int i = 0; i < n; ++i) if (H.cur[i]) { r.c[0] += x[i]; r.c[1] += y[i]; r.c[2] += z[i]; }
            for (int k = 0; k < 3; ++k) r.c[k] /= (double)H.cnt;
            for (int i = 0; i < n; ++i) if (H.cur[i]) {
                const double a = x[i] - r.c[0], b = y[i] - r.c[1], c = z[i] - r.c[2];
                r.s[0] += a * a; r.s[1] += a * b; r.s[2] += a * c; r.s[3] += b * b; r.s[4] += b * c; r.s[5] += c * c;
            }
            H.refit = calk::plane_refit_solve(r, H.P2);
            if (H.refit) H.cnt2 = score(n, x, y, z, H.P2, o.thresh, H.ref);
        }
        for (int h = 0; h < B && it < dyn; ++h) {
            ++it;
            const Hyp& H = hyp[h];
            if (!H.valid) continue;
            int cnt = H.cnt;
            if (cnt < o.min_inliers) continue;
            const double* P = H.P; const std::vector<uint8_t>* fin = &H.cur;
            if (H.refit) { P = H.P2; cnt = H.cnt2; fin = &H.ref; }
            if (!has_best || cnt >= best_cnt) {
                double ss = 0;
                for (int i = 0; i < n; ++i) if ((*fin)[i]) { const double r = P[0] * x[i] + P[1] * y[i] + P[2] * z[i] + P[3]; ss += r * r; }
                const double frms = cnt > 0 ? std::sqrt(ss / (double)cnt) : INFINITY;
                if (!has_best || cnt > best_cnt || frms < best_rms) {
                    has_best = true; best_cnt = cnt; best_rms = frms; best_iters = it;
                    std::memcpy(bestP, P, sizeof bestP); best = *fin;
                }
            }
            dyn = next_iteration_bound(table[cnt], it, o.max_iters);
        }
    }
    res->success = has_best; res->iters = best_iters; res->n_inliers = has_best ? best_cnt : 0; res->iters_run = it;
    std::memcpy(res->plane, bestP, sizeof bestP);
    res->inlier_rms = best_rms; res->min_margin = 0.0;
    for (int i = 0; i < n; ++i) mask[i] = has_best ? best[i] : 0;
    return 0;
}
