// TEST INFRASTRUCTURE — CPU emulation of the bundle-kind Jacobian pass, compiled with g++ (no GPU).
//
// It runs the PRODUCT's own per-observation math (calibration_b200/csrc/k1_math.cuh: obs_rows, the pose
// chains, frames and chain-rule transforms, the Huber weight) and the PRODUCT's own host assembly
// (refine_model.hpp: parameter blocks, assemble_shared) in the exact dataflow of K1's fused epilogue
// — per block: local system N, s_b -> w_b, Q = w T^T N_xixi, H_vv, g_v, E_vi; per camera: sums — so
// that tests/test_host_emul.py can compare (cost, J^T r, J^T J) with the oracle on the CPU suite.
// Only the CUDA kernels' parallel decomposition (roles, tiles, reductions) is not exercised here.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../calibration_b200/csrc/refine_model.hpp"

using namespace calk;

namespace {

template <int MODEL, int IMODE>
void block_sums(const HostModel& M, const cal_problem_desc& d, const double* x, std::vector<double>& cam_sums) {
    using LT = Local<MODEL, IMODE>;
    const ProblemShape& S = M.S;
    constexpr int NL = LT::NL, NC = LT::NC, NE = LT::NE, PI = LT::PI;
    for (int64_t b = 0; b < d.n_blocks; ++b) {
        const int cam = d.block_cam[b];
        CamConst cc; cam_const_from_intr(x + S.off_intr + cam * S.P, S.model, cc);
        BlockPose bp;
        compose_bundle(x + S.off_viewq, x + S.off_viewt, x + S.off_camq + 4 * cam, x + S.off_camt + 3 * cam, d.block_b_se3_g + 12 * b, bp);
        double A[9], T[36];
        block_frame(bp, cc.Rs, A);
        view_transform(bp, cc.Rs, T);
        double N[NE]; for (int e = 0; e < NE; ++e) N[e] = 0.0;
        for (int64_t i = d.block_offset[b]; i < d.block_offset[b + 1]; ++i) {
            double Ju[NL], Jv[NL];
            obs_rows<MODEL, IMODE>(cc, A, d.obj_x[i], d.obj_y[i], d.img_u[i], d.img_v[i], Ju, Jv);
            for (int a = 0; a < NL; ++a) for (int c = a; c < NL; ++c) {
                double v = N[LT::idx(a, c)];
                if (LT::has_u(a) && LT::has_u(c)) v += Ju[a] * Ju[c];
                if (LT::has_v(a) && LT::has_v(c)) v += Jv[a] * Jv[c];
                N[LT::idx(a, c)] = v;
            }
        }
        const double ssr = N[LT::idx(NC, NC)];
        double rho, w; huber_weight(S.huber_delta, ssr, rho, w);
        double* sums = &cam_sums[(size_t)cam * S.NV];
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

}  // namespace

extern "C" int emul_bundle_eval(const cal_problem_desc* dp, const double* x, double* cost, double* g, double* H) {
    const cal_problem_desc& d = *dp;
    if (d.kind != CAL_KIND_BUNDLE) return 1;
    HostModel M; M.init_model(d);
    const ProblemShape& S = M.S;
    std::vector<double> cam_sums((size_t)S.n_cams * S.NV, 0.0);
    if (S.model == 0 && S.imode == 0) block_sums<0, 0>(M, d, x, cam_sums);
    else if (S.model == 0 && S.imode == 1) block_sums<0, 1>(M, d, x, cam_sums);
    else if (S.model == 0 && S.imode == 2) block_sums<0, 2>(M, d, x, cam_sums);
    else if (S.model == 1 && S.imode == 0) block_sums<1, 0>(M, d, x, cam_sums);
    else if (S.model == 1 && S.imode == 1) block_sums<1, 1>(M, d, x, cam_sums);
    else block_sums<1, 2>(M, d, x, cam_sums);
    double c = 0; for (int k = 0; k < S.n_cams; ++k) c += cam_sums[(size_t)k * S.NV + S.NE];
    *cost = c;
    std::vector<double> Hss, gs;
    M.assemble_shared(cam_sums.data(), x, Hss, gs);
    std::memcpy(g, gs.data(), gs.size() * sizeof(double));
    std::memcpy(H, Hss.data(), Hss.size() * sizeof(double));
    return 0;
}
// ---- kinds with per-view pose blocks (intrinsics, extrinsics): K1's VIEW_STORE epilogue + k_view_gather ----
namespace {
template <int MODEL, int IMODE>
void view_kind_sums(const HostModel& M, const cal_problem_desc& d, const double* x, std::vector<double>& cam_sums, std::vector<double>& Hpp,
                    std::vector<double>& gp, std::vector<double>& Evc, std::vector<double>& Evi) {
    using LT = Local<MODEL, IMODE>;
    const ProblemShape& S = M.S;
    constexpr int NL = LT::NL, NC = LT::NC, NE = LT::NE, PI = LT::PI;
    const int64_t nb = d.n_blocks;
    for (int64_t b = 0; b < nb; ++b) {
        const int cam = d.block_cam[b];
        const int v = S.kind == CAL_KIND_INTRINSICS ? (int)b : d.block_view[b];
        CamConst cc; cam_const_from_intr(x + S.off_intr + (S.kind == CAL_KIND_INTRINSICS ? 0 : cam * S.P), S.model, cc);
        BlockPose bp;
        if (S.kind == CAL_KIND_INTRINSICS) compose_intrinsics(x + S.off_viewq + 4 * v, x + S.off_viewt + 3 * v, bp);
        else compose_extrinsics(x + S.off_camq + 4 * cam, x + S.off_camt + 3 * cam, x + S.off_viewq + 4 * v, x + S.off_viewt + 3 * v, bp);
        double A[9], T[36], Tc[36];
        block_frame(bp, cc.Rs, A);
        view_transform(bp, cc.Rs, T);
        for (int i = 0; i < 36; ++i) Tc[i] = 0.0;
        if (S.cam_pose_kind == 1) cam_transform_extrinsics(x + S.off_camt + 3 * cam, cc.Rs, Tc);
        double N[NE]; for (int e = 0; e < NE; ++e) N[e] = 0.0;
        for (int64_t i = d.block_offset[b]; i < d.block_offset[b + 1]; ++i) {
            double Ju[NL], Jv[NL];
            obs_rows<MODEL, IMODE>(cc, A, d.obj_x[i], d.obj_y[i], d.img_u[i], d.img_v[i], Ju, Jv);
            for (int a = 0; a < NL; ++a) for (int c = a; c < NL; ++c) {
                double t = N[LT::idx(a, c)];
                if (LT::has_u(a) && LT::has_u(c)) t += Ju[a] * Ju[c];
                if (LT::has_v(a) && LT::has_v(c)) t += Jv[a] * Jv[c];
                N[LT::idx(a, c)] = t;
            }
        }
        double rho, w; huber_weight(S.huber_delta, N[LT::idx(NC, NC)], rho, w);
        double* sums = &cam_sums[(size_t)cam * S.NV];
        for (int e = 0; e < NE; ++e) sums[e] += w * N[e];
        sums[NE] += 0.5 * rho;
        if (M.pbs[M.pb_viewq(v)].constant) continue;
        auto Nxx = [&](int k, int j) { return N[k <= j ? LT::idx(k, j) : LT::idx(j, k)]; };
        double Q[36];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += T[6 * k + i] * Nxx(k, j); Q[6 * i + j] = a * w; }
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += Q[6 * i + k] * T[6 * k + j]; Hpp[(size_t)v * 36 + 6 * i + j] += a; }
        for (int i = 0; i < 6; ++i) { double a = 0; for (int k = 0; k < 6; ++k) a += T[6 * k + i] * N[LT::idx(k, NC)]; gp[(size_t)v * 6 + i] += a * w; }
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += Q[6 * i + k] * Tc[6 * k + j]; Evc[(size_t)(6 * i + j) * nb + b] = a; }
        for (int j = 0; j < PI; ++j) for (int i = 0; i < 6; ++i) { double a = 0; for (int k = 0; k < 6; ++k) a += T[6 * k + i] * N[LT::idx(k, 6 + j)]; Evi[(size_t)(PI * i + j) * nb + b] = a * w; }
    }
}
}  // namespace

extern "C" int emul_views_eval(const cal_problem_desc* dp, const double* x, double* cost, double* g, double* H) {
    const cal_problem_desc& d = *dp;
    if (d.kind == CAL_KIND_BUNDLE) return 1;
    HostModel M; M.init_model(d);
    const ProblemShape& S = M.S;
    const int nv = S.n_views; const int64_t nb = d.n_blocks;
    std::vector<double> cam_sums((size_t)S.n_cams * S.NV, 0.0), Hpp((size_t)nv * 36, 0.0), gp((size_t)nv * 6, 0.0), Evc((size_t)36 * nb, 0.0),
        Evi((size_t)6 * std::max(S.PI, 1) * nb, 0.0);
    if (S.model == 0 && S.imode == 0) view_kind_sums<0, 0>(M, d, x, cam_sums, Hpp, gp, Evc, Evi);
    else if (S.model == 0 && S.imode == 1) view_kind_sums<0, 1>(M, d, x, cam_sums, Hpp, gp, Evc, Evi);
    else if (S.model == 0 && S.imode == 2) view_kind_sums<0, 2>(M, d, x, cam_sums, Hpp, gp, Evc, Evi);
    else if (S.model == 1 && S.imode == 0) view_kind_sums<1, 0>(M, d, x, cam_sums, Hpp, gp, Evc, Evi);
    else if (S.model == 1 && S.imode == 1) view_kind_sums<1, 1>(M, d, x, cam_sums, Hpp, gp, Evc, Evi);
    else view_kind_sums<1, 2>(M, d, x, cam_sums, Hpp, gp, Evc, Evi);
    double c = 0; for (int k = 0; k < S.n_cams; ++k) c += cam_sums[(size_t)k * S.NV + S.NE];
    *cost = c;
    std::vector<double> Hss, gs, Hd, gd;
    M.assemble_shared(cam_sums.data(), x, Hss, gs);
    // view CSR over the (original) blocks
    std::vector<int32_t> off(nv + 1, 0), idx(nb), cam(nb); std::vector<char> vfree(nv);
    for (int64_t b = 0; b < nb; ++b) { off[(S.kind == CAL_KIND_INTRINSICS ? (int)b : d.block_view[b]) + 1]++; cam[b] = d.block_cam[b]; }
    for (int v = 0; v < nv; ++v) { off[v + 1] += off[v]; vfree[v] = !M.pbs[M.pb_viewq(v)].constant; }
    { std::vector<int32_t> cur(off.begin(), off.end() - 1); for (int64_t b = 0; b < nb; ++b) idx[cur[S.kind == CAL_KIND_INTRINSICS ? (int)b : d.block_view[b]]++] = (int32_t)b; }
    M.assemble_dense(Hss, gs, Hpp.data(), gp.data(), Evc.data(), Evi.data(), nb, vfree.data(), off.data(), idx.data(), cam.data(), Hd, gd);
    std::memcpy(g, gd.data(), gd.size() * sizeof(double));
    std::memcpy(H, Hd.data(), Hd.size() * sizeof(double));
    return 0;
}

extern "C" int emul_tangent_count(const cal_problem_desc* dp) { HostModel M; M.init_model(*dp); return M.n_tan; }
