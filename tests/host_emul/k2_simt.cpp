// TEST INFRASTRUCTURE — the PRODUCT's K2 kernels (calibration_b200/csrc/refine_schur_kernels.cuh: k_view_gather,
// k_view_scale, k_schur_factor, k_schur_syrk, k_schur_reduce, k_backsub, k_reduce_views, and the covariance pair
// k_cov_view_prep / k_cov_vv) compiled by g++ and run on
// the CPU under the lock-step SIMT shim, launched in the order and with the grids of one LM iteration of
// cal_refine_solve (refine_host.cu) / launch_schur (refine_kernels.cu).  Input: the per-residual-block products K1's
// epilogue stores for the per-view kinds (H_vv, g_v, E_vc, E_vi) and the damped, Jacobi-scaled shared block; output:
// the Schur complement, the reduced solution, the per-view steps and the per-view terms of the model cost change.
#define SIMT_SHARED_STORAGE static
#include "simt_shim.hpp"

#include "../../calibration_b200/csrc/refine_schur_kernels.cuh"

using namespace calk;

namespace {
// launch_schur's grid (refine_kernels.cu: schur_num_ctas): up to 444 CTAs, at least 16 views each
int schur_ctas(int n_views) { return n_views < 64 ? 1 : (n_views < 444 * 16 ? (n_views + 15) / 16 : 444); }

// dense SPD solve (the product factorises the reduced system on the host with chol_host; any exact solver checks it)
bool solve_spd(std::vector<double> A, int n, double* b) {
    for (int j = 0; j < n; ++j) {
        double d = A[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
        if (!(d > 0)) return false;
        d = std::sqrt(d); A[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
            A[(size_t)i * n + j] = s / d;
        }
    }
    for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
    for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
    return true;
}
}  // namespace

extern "C" int simt_k2_step(int n_views, int n_cams, int PI, int64_t n_blk, const int32_t* blk_cam, const int32_t* blk_view,
                            const int32_t* view_free, const int32_t* cam_col_q, const int32_t* cam_col_t, const int32_t* cam_col_i,
                            const double* blk_Hvv, const double* blk_gv, const double* blk_Evc, const double* blk_Evi, int ns,
                            const double* s_shared, const double* Hss_scaled_damped, const double* gs_scaled, double radius,
                            double* C, double* c, double* y_shared, double* delta_p, double* red_out4, int32_t* fail, double* sp_out) {
    if (ns + 1 > kSyrkMaxN) return 2;
    ProblemShape S{};
    S.n_views = n_views; S.n_cams = n_cams; S.PI = PI;
    DevLayout L;
    L.n_blk = n_blk;
    std::vector<int32_t> bcam(blk_cam, blk_cam + n_blk), bview(blk_view, blk_view + n_blk);
    L.blk_cam = bcam.data(); L.blk_view = bview.data();
    EvalBuffers B;
    B.blk_Hvv = const_cast<double*>(blk_Hvv); B.blk_gv = const_cast<double*>(blk_gv);
    B.blk_Evc = const_cast<double*>(blk_Evc); B.blk_Evi = const_cast<double*>(blk_Evi);
    // view -> blocks CSR (cal_refine_create builds the same)
    std::vector<int32_t> off(n_views + 1, 0), idx(n_blk), vfree(view_free, view_free + n_views), cq(cam_col_q, cam_col_q + n_cams),
        ct(cam_col_t, cam_col_t + n_cams), ci(cam_col_i, cam_col_i + n_cams);
    for (int64_t b = 0; b < n_blk; ++b) off[blk_view[b] + 1]++;
    for (int v = 0; v < n_views; ++v) off[v + 1] += off[v];
    { std::vector<int32_t> cur(off.begin(), off.end() - 1); for (int64_t b = 0; b < n_blk; ++b) idx[cur[blk_view[b]]++] = (int32_t)b; }
    const int n_cta = schur_ctas(n_views), na = ns + 1, ncb = 6 + PI;
    std::vector<double> Hpp((size_t)n_views * 36), gp((size_t)n_views * 6), sp((size_t)n_views * 6, 0.0), dp((size_t)n_views * 6), Lp((size_t)n_views * 36), Linv((size_t)n_views * 6),
        view_f((size_t)n_views * 6), Fd((size_t)n_views * 6 * ((ns + 1 + kSyrkTile - 1) / kSyrkTile * kSyrkTile), 0.0), dlt((size_t)n_views * 6, 0.0), ss(s_shared, s_shared + ns), ys(std::max(ns, 1)),
        Cm((size_t)ns * ns), cv(std::max(ns, 1)), partialC((size_t)n_cta * na * na), red((size_t)n_views * 4, 0.0), ro(4, 0.0);
    int32_t failed = 0;
    ViewBuffers V;
    V.view_blk_off = off.data(); V.view_blk_idx = idx.data(); V.view_free = vfree.data();
    V.cam_col_q = cq.data(); V.cam_col_t = ct.data(); V.cam_col_i = ci.data();
    V.Hpp = Hpp.data(); V.gp = gp.data(); V.sp = sp.data(); V.dp = dp.data(); V.Lp = Lp.data(); V.Linv = Linv.data(); V.view_f = view_f.data(); V.Fd = Fd.data(); V.ns = ns; V.ncp = (ns + 1 + kSyrkTile - 1) / kSyrkTile * kSyrkTile;
    V.delta_p = dlt.data(); V.s_shared = ss.data(); V.y_shared = ys.data(); V.C = Cm.data(); V.c = cv.data(); V.partialC = partialC.data();
    std::vector<double> rpart((size_t)kReduceViewsCtas * 4, 0.0); unsigned rticket = 0u;
    V.red = red.data(); V.red_out = ro.data(); V.fail = &failed; V.red_part = rpart.data(); V.red_ticket = &rticket;
    // after the Jacobian pass: gather, Jacobi scaling once (compute_scale = 1); first LM iteration: the diagonal clamp
    simt::launch((unsigned)((n_views + 127) / 128), 128, [&] { k_view_gather(S, L, B, V); });
    simt::launch((unsigned)((n_views * 6 + 127) / 128), 128, [&] { k_view_scale(S, V, 1); });
    simt::launch((unsigned)((n_views * 6 + 127) / 128), 128, [&] { k_view_scale(S, V, 0); });
    // launch_schur
    simt::launch((unsigned)((S.n_views + 127) / 128), 128, [&] { k_view_chol(S, V, 1.0 / radius); });
    simt::launch((unsigned)(((int64_t)S.n_views * 32 + kFactorThreads - 1) / kFactorThreads), kFactorThreads, [&] { k_schur_factor(S, L, B, V); });
    const int per = (n_views + n_cta - 1) / n_cta, nt = (na + kSyrkTile - 1) / kSyrkTile, threads = (nt * (nt + 1) / 2 + 31) / 32 * 32;
    { int bt = 0, warps = 0; syrk_shape(ns, &bt, &warps); simt::launch((unsigned)n_cta, (unsigned)(32 * warps), [&] { k_schur_syrk_any(bt, S, V, ns, per); }); }
    simt::launch((unsigned)((na * na + kSchurReduceEntries - 1) / kSchurReduceEntries), 4 * kSchurReduceEntries, [&] { k_schur_reduce(V, n_cta, ns); });
    *fail = failed;
    std::memcpy(C, Cm.data(), sizeof(double) * ns * ns); std::memcpy(c, cv.data(), sizeof(double) * ns);
    std::memcpy(sp_out, sp.data(), sizeof(double) * 6 * n_views);
    if (failed) return 0;
    // host: (S Hss S + D / radius - C) y = S gs - c
    std::vector<double> Sm((size_t)ns * ns);
    for (int i = 0; i < ns; ++i) { ys[i] = gs_scaled[i] - cv[i]; for (int j = 0; j < ns; ++j) Sm[(size_t)i * ns + j] = Hss_scaled_damped[(size_t)i * ns + j] - Cm[(size_t)i * ns + j]; }
    if (ns > 0 && !solve_spd(Sm, ns, ys.data())) return 3;
    if (ns > 0 && ns <= kReducedMaxN) {   // the product's own reduced solve (k_reduced_solve) must agree with the dense host solve above
        std::vector<double> y_host(ys.begin(), ys.begin() + ns), Smd(Hss_scaled_damped, Hss_scaled_damped + (size_t)ns * ns), gsd(gs_scaled, gs_scaled + ns);
        int32_t info[2] = {0, 0};
        simt::launch(1, 256, [&] { k_reduced_solve(Smd.data(), gsd.data(), V, ns, info); });
        if (info[0]) return 4;
        double ref = 0.0, dev = 0.0;
        for (int i = 0; i < ns; ++i) { ref = std::max(ref, std::fabs(y_host[i])); dev = std::max(dev, std::fabs(y_host[i] - ys[i])); }
        if (!(dev <= 1e-9 * (ref + 1e-300))) return 5;
    }
    std::memcpy(y_shared, ys.data(), sizeof(double) * ns);
    simt::launch((unsigned)((n_views + kBacksubViews - 1) / kBacksubViews), 256, [&] { k_backsub(S, L, V, ns); });
    simt::launch(n_views >= 2048 ? 2u : 1u, 256, [&] { k_reduce_views(V, n_views); });
    std::memcpy(delta_p, dlt.data(), sizeof(double) * 6 * n_views);
    std::memcpy(red_out4, ro.data(), sizeof(double) * 4);
    return 0;
}

// Block-structured covariance of the per-view kinds (cal_refine_solve's covariance branch): launch_schur with an
// infinite radius, W = (S Hss S - C)^-1 on the host, k_cov_view_prep, k_cov_vv.  The parameter vector of this harness
// holds the view blocks only: x = [quat(4) x n_views | tran(3) x n_views], so cov_vv is (7 n_views)^2.
// Hss_scaled: the UNDAMPED scaled shared block.  Outputs: W [ns][ns], Z and G [n_views][6][ns], Ainv [n_views][36].
extern "C" int simt_k2_cov(int n_views, int n_cams, int PI, int64_t n_blk, const int32_t* blk_cam, const int32_t* blk_view,
                           const int32_t* view_free, const int32_t* cam_col_q, const int32_t* cam_col_t, const int32_t* cam_col_i,
                           const double* blk_Hvv, const double* blk_gv, const double* blk_Evc, const double* blk_Evi, int ns,
                           const double* s_shared, const double* Hss_scaled, const double* view_quats, double* W_out, double* Z_out,
                           double* G_out, double* Ainv_out, double* cov_vv) {
    if (ns + 1 > kSyrkMaxN) return 2;
    ProblemShape S{};
    S.n_views = n_views; S.n_cams = n_cams; S.PI = PI; S.off_viewq = 0; S.off_viewt = 4 * n_views;
    const int64_t na_amb = 7 * (int64_t)n_views;
    DevLayout L;
    L.n_blk = n_blk;
    std::vector<int32_t> bcam(blk_cam, blk_cam + n_blk), bview(blk_view, blk_view + n_blk);
    L.blk_cam = bcam.data(); L.blk_view = bview.data();
    EvalBuffers B;
    B.blk_Hvv = const_cast<double*>(blk_Hvv); B.blk_gv = const_cast<double*>(blk_gv);
    B.blk_Evc = const_cast<double*>(blk_Evc); B.blk_Evi = const_cast<double*>(blk_Evi);
    std::vector<int32_t> off(n_views + 1, 0), idx(n_blk), vfree(view_free, view_free + n_views), cq(cam_col_q, cam_col_q + n_cams),
        ct(cam_col_t, cam_col_t + n_cams), ci(cam_col_i, cam_col_i + n_cams);
    for (int64_t b = 0; b < n_blk; ++b) off[blk_view[b] + 1]++;
    for (int v = 0; v < n_views; ++v) off[v + 1] += off[v];
    { std::vector<int32_t> cur(off.begin(), off.end() - 1); for (int64_t b = 0; b < n_blk; ++b) idx[cur[blk_view[b]]++] = (int32_t)b; }
    const int n_cta = schur_ctas(n_views), na = ns + 1, ncb = 6 + PI;
    std::vector<double> Hpp((size_t)n_views * 36), gp((size_t)n_views * 6), sp((size_t)n_views * 6, 0.0), dp((size_t)n_views * 6), Lp((size_t)n_views * 36), Linv((size_t)n_views * 6),
        view_f((size_t)n_views * 6), Fd((size_t)n_views * 6 * ((ns + 1 + kSyrkTile - 1) / kSyrkTile * kSyrkTile), 0.0), ss(s_shared, s_shared + ns), Cm((size_t)ns * ns), cv(std::max(ns, 1)),
        partialC((size_t)n_cta * na * na), x(na_amb, 0.0);
    std::memcpy(x.data(), view_quats, sizeof(double) * 4 * n_views);
    int32_t failed = 0;
    ViewBuffers V;
    V.view_blk_off = off.data(); V.view_blk_idx = idx.data(); V.view_free = vfree.data();
    V.cam_col_q = cq.data(); V.cam_col_t = ct.data(); V.cam_col_i = ci.data();
    V.Hpp = Hpp.data(); V.gp = gp.data(); V.sp = sp.data(); V.dp = dp.data(); V.Lp = Lp.data(); V.Linv = Linv.data(); V.view_f = view_f.data(); V.Fd = Fd.data(); V.ns = ns; V.ncp = (ns + 1 + kSyrkTile - 1) / kSyrkTile * kSyrkTile;
    V.s_shared = ss.data(); V.C = Cm.data(); V.c = cv.data(); V.partialC = partialC.data(); V.fail = &failed;
    simt::launch((unsigned)((n_views + 127) / 128), 128, [&] { k_view_gather(S, L, B, V); });
    simt::launch((unsigned)((n_views * 6 + 127) / 128), 128, [&] { k_view_scale(S, V, 1); });
    // launch_schur(radius = infinity): undamped factors L_v and F_b = L_v^-1 E_b stay in V
    simt::launch((unsigned)((S.n_views + 127) / 128), 128, [&] { k_view_chol(S, V, 0.0); });
    simt::launch((unsigned)(((int64_t)S.n_views * 32 + kFactorThreads - 1) / kFactorThreads), kFactorThreads, [&] { k_schur_factor(S, L, B, V); });
    const int per = (n_views + n_cta - 1) / n_cta, nt = (na + kSyrkTile - 1) / kSyrkTile, threads = (nt * (nt + 1) / 2 + 31) / 32 * 32;
    { int bt = 0, warps = 0; syrk_shape(ns, &bt, &warps); simt::launch((unsigned)n_cta, (unsigned)(32 * warps), [&] { k_schur_syrk_any(bt, S, V, ns, per); }); }
    simt::launch((unsigned)((na * na + kSchurReduceEntries - 1) / kSchurReduceEntries), 4 * kSchurReduceEntries, [&] { k_schur_reduce(V, n_cta, ns); });
    if (failed) return 4;
    // host: W = (S Hss S - C)^-1, column by column (cal_refine_solve does the same with chol_host / chol_solve_host)
    std::vector<double> Sm((size_t)ns * ns), W((size_t)ns * ns), e(std::max(ns, 1));
    for (int i = 0; i < ns; ++i) for (int j = 0; j < ns; ++j) Sm[(size_t)i * ns + j] = Hss_scaled[(size_t)i * ns + j] - Cm[(size_t)i * ns + j];
    for (int j = 0; j < ns; ++j) {
        std::fill(e.begin(), e.end(), 0.0); e[j] = 1.0;
        if (!solve_spd(Sm, ns, e.data())) return 3;
        for (int i = 0; i < ns; ++i) W[(size_t)i * ns + j] = e[i];
    }
    std::vector<double> Z((size_t)n_views * 6 * std::max(ns, 1), 0.0), G((size_t)n_views * 6 * std::max(ns, 1), 0.0), Ainv((size_t)n_views * 36, 0.0),
        cov((size_t)na_amb * na_amb, 0.0);
    simt::launch((unsigned)n_views, 64, [&] { k_cov_view_prep(S, L, V, ns, W.data(), Z.data(), G.data(), Ainv.data()); });
    const unsigned t = (unsigned)((n_views + kCovTile - 1) / kCovTile);
    simt::launch(t, kCovTile * kCovTile, [&] { k_cov_vv(S, V, x.data(), ns, Z.data(), G.data(), Ainv.data(), cov.data(), na_amb); }, t);
    std::memcpy(W_out, W.data(), sizeof(double) * ns * ns);
    std::memcpy(Z_out, Z.data(), sizeof(double) * (size_t)n_views * 6 * ns);
    std::memcpy(G_out, G.data(), sizeof(double) * (size_t)n_views * 6 * ns);
    std::memcpy(Ainv_out, Ainv.data(), sizeof(double) * (size_t)n_views * 36);
    std::memcpy(cov_vv, cov.data(), sizeof(double) * na_amb * na_amb);
    return 0;
}

// k_view_plus (candidate x [+] t delta_p per view, |dx|^2) and k_view_norms (|x_v|^2, max |x_v - plus(x_v, -g_v)|) with
// k_reduce_views, as make_candidate / norms of cal_refine_solve launch them.  x = [quat(4) x n_views | tran(3) x n_views].
extern "C" int simt_k2_plus_norms(int n_views, const double* x, const double* delta_p, const double* gp, const int32_t* view_free, double t,
                                  double* x_cand, double* red_plus4, double* red_norms4) {
    ProblemShape S{};
    S.n_views = n_views; S.off_viewq = 0; S.off_viewt = 4 * n_views;
    std::vector<double> xs(x, x + 7 * (size_t)n_views), xc(7 * (size_t)n_views, 0.0), dl(delta_p, delta_p + 6 * (size_t)n_views),
        g(gp, gp + 6 * (size_t)n_views), red((size_t)n_views * 4, 0.0), ro(4, 0.0);
    std::vector<int32_t> vfree(view_free, view_free + n_views);
    EvalBuffers B; B.x = xs.data();
    ViewBuffers V;
    V.x_cand = xc.data(); V.delta_p = dl.data(); V.gp = g.data(); V.view_free = vfree.data(); V.red = red.data(); V.red_out = ro.data();
    std::vector<double> rpart((size_t)kReduceViewsCtas * 4, 0.0); unsigned rticket = 0u; V.red_part = rpart.data(); V.red_ticket = &rticket;
    simt::launch((unsigned)((n_views + 127) / 128), 128, [&] { k_view_plus(S, B, V, t); });
    simt::launch(n_views >= 2048 ? 2u : 1u, 256, [&] { k_reduce_views(V, n_views); });
    std::memcpy(red_plus4, ro.data(), 4 * sizeof(double));
    std::memcpy(x_cand, xc.data(), sizeof(double) * 7 * n_views);
    std::fill(red.begin(), red.end(), 0.0);
    simt::launch((unsigned)((n_views + 127) / 128), 128, [&] { k_view_norms(S, B, V); });
    simt::launch(n_views >= 2048 ? 2u : 1u, 256, [&] { k_reduce_views(V, n_views); });
    std::memcpy(red_norms4, ro.data(), 4 * sizeof(double));
    return 0;
}

// The product's reduced solve alone (k_reduced_solve, one CTA): (Sm - C) y = gss - c against the dense host solve.
// Returns 0 and the largest |y - y_ref| / max|y_ref| in err[0]; 1: the kernel flagged the matrix (info), 2: n too wide,
// 3: the host solver found the matrix not positive definite.
extern "C" int simt_k2_reduced_solve(int n, const double* Sm, const double* Cm, const double* gss, const double* cv, double* y_out, double* err) {
    if (n > kReducedMaxN) return 2;
    ViewBuffers V;
    std::vector<double> C(Cm, Cm + (size_t)n * n), c(cv, cv + n), y(std::max(n, 1), 0.0);
    V.C = C.data(); V.c = c.data(); V.y_shared = y.data();
    int32_t info[2] = {0, 0};
    simt::launch(1, 256, [&] { k_reduced_solve(Sm, gss, V, n, info); });
    std::memcpy(y_out, y.data(), sizeof(double) * n);
    if (info[0]) return 1;
    std::vector<double> A((size_t)n * n), b(n);
    for (int i = 0; i < n; ++i) { b[i] = gss[i] - cv[i]; for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = Sm[(size_t)i * n + j] - Cm[(size_t)i * n + j]; }
    if (!solve_spd(A, n, b.data())) return 3;
    double ref = 0.0, dev = 0.0;
    for (int i = 0; i < n; ++i) { ref = std::max(ref, std::fabs(b[i])); dev = std::max(dev, std::fabs(b[i] - y[i])); }
    err[0] = dev / (ref + 1e-300);
    return 0;
}
