// ORACLE — test / benchmark infrastructure only (never loaded by the product).
//
// A CONSERVATIVE CPU BASELINE for the fused residual + Jacobian + J^T J pass of the bundle kind: hand-derived
// Jacobians instead of the forward-mode duals that the reference's ceres::AutoDiffCostFunction evaluates
// (src/estimation/residuals/bundleresidual.h:15-47, solve_problem src/estimation/detail/ceresutils.h:27-43), OpenMP over
// residual blocks, per-thread accumulation of the per-camera sums.  It answers "how fast could the reference's CPU
// path be with analytic derivatives": BASELINE.md promises that baseline next to the dual-number restatement
// (refine.cpp), which is the faithful one.  The per-observation arithmetic is the shared host/device header of the
// product (calibration_b200/csrc/k1_math.cuh: obs_rows, pose chains, chain-rule transforms) and the host model
// (refine_model.hpp: parameter blocks, assembly of the shared system) — compiled here for the CPU with g++ -O3;
// tests/test_oracle_analytic.py holds it to the dual-number restatement.
#include <omp.h>

#include <cstring>
#include <vector>

#include "../calibration_b200/csrc/refine_model.hpp"

using namespace calk;

namespace {

template <int MODEL, int IMODE>
void block_sums(const HostModel& M, const cal_problem_desc& d, const double* x, std::vector<double>& cam_sums, int threads) {
    using LT = Local<MODEL, IMODE>;
    const ProblemShape& S = M.S;
    constexpr int NL = LT::NL, NC = LT::NC, NE = LT::NE, PI = LT::PI;
    const size_t n_sums = (size_t)S.n_cams * S.NV;
    std::vector<std::vector<double>> part((size_t)threads, std::vector<double>(n_sums, 0.0));
#pragma omp parallel num_threads(threads)
    {
        std::vector<double>& mine = part[(size_t)omp_get_thread_num()];
#pragma omp for schedule(static)
        for (int64_t b = 0; b < d.n_blocks; ++b) {
            const int cam = d.block_cam[b];
            CamConst cc; cam_const_from_intr(x + S.off_intr + cam * S.P, S.model, cc);
            BlockPose bp;
            compose_bundle(x + S.off_viewq, x + S.off_viewt, x + S.off_camq + 4 * cam, x + S.off_camt + 3 * cam, d.block_b_se3_g + 12 * b, bp);
            double A[9], T[36];
            block_frame(bp, cc.Rs, A);
            view_transform(bp, cc.Rs, T);
            double N[NE];
            for (int e = 0; e < NE; ++e) N[e] = 0.0;
            const bool board = d.board_n > 0;
            for (int64_t i = d.block_offset[b]; i < d.block_offset[b + 1]; ++i) {
                double Ju[NL], Jv[NL];
                const int64_t k = board ? i - d.block_offset[b] : i;
                obs_rows<MODEL, IMODE>(cc, A, board ? d.board_x[k] : d.obj_x[k], board ? d.board_y[k] : d.obj_y[k], d.img_u[i], d.img_v[i], Ju, Jv);
                int e = 0;
                for (int a = 0; a < NL; ++a)
                    for (int c = a; c < NL; ++c, ++e) {
                        double v = N[e];
                        if (LT::has_u(a) && LT::has_u(c)) v += Ju[a] * Ju[c];
                        if (LT::has_v(a) && LT::has_v(c)) v += Jv[a] * Jv[c];
                        N[e] = v;
                    }
            }
            const double ssr = N[LT::idx(NC, NC)];
            double rho, w; huber_weight(S.huber_delta, ssr, rho, w);
            double* sums = &mine[(size_t)cam * S.NV];
            for (int e = 0; e < NE; ++e) sums[e] += w * N[e];
            sums[NE] += 0.5 * rho;
            if (!S.view_free_global) continue;
            auto Nxx = [&](int k, int j) { return N[k <= j ? LT::idx(k, j) : LT::idx(j, k)]; };
            double Q[36];
            for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += T[6 * k + i] * Nxx(k, j); Q[6 * i + j] = a * w; }
            double* Hvv = sums + NE + 1; double* gv = Hvv + 21; double* Qs = gv + 6; double* Evi = Qs + 36;
            int o = 0;
            for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += Q[6 * i + k] * T[6 * k + j]; Hvv[o++] += a; }
            for (int i = 0; i < 6; ++i) { double a = 0; for (int k = 0; k < 6; ++k) a += T[6 * k + i] * N[LT::idx(k, NC)]; gv[i] += a * w; }
            for (int i = 0; i < 36; ++i) Qs[i] += Q[i];
            for (int j = 0; j < PI; ++j) for (int i = 0; i < 6; ++i) { double a = 0; for (int k = 0; k < 6; ++k) a += T[6 * k + i] * N[LT::idx(k, 6 + j)]; Evi[PI * i + j] += a * w; }
        }
    }
    cam_sums.assign(n_sums, 0.0);
    for (const auto& p : part) for (size_t i = 0; i < n_sums; ++i) cam_sums[i] += p[i];   // thread order: fixed
}

}  // namespace

// cost, gradient and J^T J of the bundle kind in the canonical tangent order of the shared blocks (as orc_refine_eval
// returns them for this kind); g / H may be NULL (then only the per-camera sums and the cost are formed)
extern "C" int orc_analytic_bundle_eval(const cal_problem_desc* dp, const double* x, double* cost, double* g, double* H, int num_threads) {
    const cal_problem_desc& d = *dp;
    if (d.kind != CAL_KIND_BUNDLE) return 1;
    HostModel M; M.init_model(d);
    const ProblemShape& S = M.S;
    const int threads = num_threads > 0 ? num_threads : omp_get_max_threads();
    std::vector<double> cam_sums;
    if (S.model == 0 && S.imode == 0) block_sums<0, 0>(M, d, x, cam_sums, threads);
    else if (S.model == 0 && S.imode == 1) block_sums<0, 1>(M, d, x, cam_sums, threads);
    else if (S.model == 0 && S.imode == 2) block_sums<0, 2>(M, d, x, cam_sums, threads);
    else if (S.model == 1 && S.imode == 0) block_sums<1, 0>(M, d, x, cam_sums, threads);
    else if (S.model == 1 && S.imode == 1) block_sums<1, 1>(M, d, x, cam_sums, threads);
    else block_sums<1, 2>(M, d, x, cam_sums, threads);
    double c = 0; for (int k = 0; k < S.n_cams; ++k) c += cam_sums[(size_t)k * S.NV + S.NE];
    if (cost) *cost = c;
    if (g || H) {
        std::vector<double> Hss, gs;
        M.assemble_shared(cam_sums.data(), x, Hss, gs);
        if (g) std::memcpy(g, gs.data(), gs.size() * sizeof(double));
        if (H) std::memcpy(H, Hss.data(), Hss.size() * sizeof(double));
    }
    return 0;
}
