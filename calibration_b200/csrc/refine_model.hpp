// Host-side model of a refinement problem: parameter blocks in the reference's order, their tangent
// layout (manifolds, constant blocks), the kernel's problem shape, and the assembly of the shared
// (per-camera / hand-eye / target-pose) normal-equation block from the per-camera sums the device
// produces.  Pure host code (no CUDA calls): used by refine_host.cu and, compiled with g++, by the CPU
// emulation test (tests/host_emul) that checks k1_math.cuh and this assembly against the oracle without
// a GPU.  Replaces the problem builders of the reference: build_problem (src/estimation/optim/
// intrinsics.cpp:63-90), set_residual_blocks / set_param_constraints (extrinsics.cpp:87-150),
// build_problem (bundle.cpp:84-133).
#pragma once
#include <algorithm>
#include <utility>
#include <vector>

#include "../../include/calib_b200.h"
#include "refine_kernels.cuh"

namespace calk {

enum PBType { PB_EUCLID = 0, PB_QUAT = 1, PB_INTR = 2 };
struct PB { int off, size, tsize, type; bool constant; int toff; };

struct HostModel {
    ProblemShape S{};
    std::vector<PB> pbs;
    int n_amb = 0, n_tan = 0, ns = 0;
    bool constrained = false;

    int pb_intr(int c) const { return S.kind == CAL_KIND_INTRINSICS ? 0 : c; }
    int pb_viewq(int v) const { return S.kind == CAL_KIND_INTRINSICS ? 1 + v : 3 * S.n_cams + v; }
    int pb_viewt(int v) const { return S.kind == CAL_KIND_INTRINSICS ? 1 + S.n_views + v : 3 * S.n_cams + S.n_views + v; }
    int pb_camq(int c) const { return S.n_cams + c; }
    int pb_camt(int c) const { return 2 * S.n_cams + c; }
    int pb_bq() const { return 3 * S.n_cams; }
    int pb_bt() const { return 3 * S.n_cams + 1; }

    void build_param_blocks(const cal_problem_desc& d) {
        const bool sk = d.optimize_skew != 0;
        const int P = S.P;
        auto add = [&](int size, int type, bool constant) {
            PB b; b.off = n_amb; b.size = size; b.type = type; b.constant = constant;
            b.tsize = type == PB_QUAT ? 3 : (type == PB_INTR && !sk ? size - 1 : size);
            b.toff = -1; n_amb += size; pbs.push_back(b);
        };
        if (d.kind == CAL_KIND_INTRINSICS) {
            add(P, PB_INTR, false);
            for (int v = 0; v < d.n_views; ++v) add(4, PB_QUAT, false);
            for (int v = 0; v < d.n_views; ++v) add(3, PB_EUCLID, false);
            constrained = true;  // lower bounds on fx, fy (intrinsics.cpp:81-82)
        } else if (d.kind == CAL_KIND_EXTRINSICS) {
            const bool oi = d.optimize_intrinsics, oe = d.optimize_extrinsics;  // extrinsics.cpp:110-150
            for (int c = 0; c < d.n_cams; ++c) add(P, PB_INTR, !oi);
            for (int c = 0; c < d.n_cams; ++c) add(4, PB_QUAT, !oe || c == 0);
            for (int c = 0; c < d.n_cams; ++c) add(3, PB_EUCLID, !oe || c == 0);
            for (int v = 0; v < d.n_views; ++v) add(4, PB_QUAT, oi && v + d.view_base == 0);
            for (int v = 0; v < d.n_views; ++v) add(3, PB_EUCLID, oi && v + d.view_base == 0);
            constrained = oi;
        } else {
            const bool oi = d.optimize_intrinsics, oh = d.optimize_hand_eye, ot = d.optimize_target_pose;  // bundle.cpp:98-131
            for (int c = 0; c < d.n_cams; ++c) add(P, PB_INTR, !oi);
            for (int c = 0; c < d.n_cams; ++c) add(4, PB_QUAT, !oh);
            for (int c = 0; c < d.n_cams; ++c) add(3, PB_EUCLID, !oh);
            add(4, PB_QUAT, !ot);
            add(3, PB_EUCLID, !ot);
            constrained = oi;
        }
        for (auto& b : pbs) if (!b.constant) { b.toff = n_tan; n_tan += b.tsize; }
        // shared block = everything that is not a per-view pose; for the per-view
        // kinds the view blocks come last in x, so shared tangent indices are 0..ns-1
        ns = n_tan;
        if (d.kind != CAL_KIND_BUNDLE) {
            ns = 0;
            const int first_view_pb = pb_viewq(0);
            for (int i = 0; i < first_view_pb; ++i) if (!pbs[i].constant) ns += pbs[i].tsize;
        }
    }

    // problem shape + parameter blocks from the caller's description
    void init_model(const cal_problem_desc& d) {
        S.kind = d.kind; S.model = d.model; S.n_cams = d.n_cams;
        S.n_views = d.kind == CAL_KIND_BUNDLE ? 0 : d.n_views;
        S.P = d.model == CAL_MODEL_SCHEIMPFLUG_BC5 ? 12 : 10;
        const bool intr_free = d.kind == CAL_KIND_INTRINSICS || d.optimize_intrinsics;
        S.imode = !intr_free ? INTR_NONE : (d.optimize_skew ? INTR_SKEW : INTR_NOSKEW);
        S.PI = S.imode == INTR_NONE ? 0 : (S.imode == INTR_NOSKEW ? S.P - 1 : S.P);
        S.NC = 6 + S.PI; S.NL = S.NC + 1; S.NE = S.NL * (S.NL + 1) / 2;
        S.huber_delta = d.huber_delta;
        S.cam_pose_kind = d.kind == CAL_KIND_EXTRINSICS ? 1 : (d.kind == CAL_KIND_BUNDLE ? 2 : 0);
        S.view_free_global = d.kind == CAL_KIND_BUNDLE ? (d.optimize_target_pose != 0) : 1;
        S.NV = S.NE + 1 + (d.kind == CAL_KIND_BUNDLE ? 63 + 6 * S.PI : 0);
        build_param_blocks(d);
        S.off_intr = 0;
        if (d.kind == CAL_KIND_INTRINSICS) { S.off_camq = S.off_camt = 0; S.off_viewq = S.P; S.off_viewt = S.P + 4 * d.n_views; }
        else if (d.kind == CAL_KIND_EXTRINSICS) {
            S.off_camq = S.P * d.n_cams; S.off_camt = S.off_camq + 4 * d.n_cams;
            S.off_viewq = S.off_camt + 3 * d.n_cams; S.off_viewt = S.off_viewq + 4 * d.n_views;
        } else {
            S.off_camq = S.P * d.n_cams; S.off_camt = S.off_camq + 4 * d.n_cams;
            S.off_viewq = S.off_camt + 3 * d.n_cams; S.off_viewt = S.off_viewq + 4;  // b_q_t, b_t_t
        }
    }

    // Shared block (tangent, canonical order) from the per-camera sums of one Jacobian pass:
    // cam_sums[c][NV] = [local system (NE) | cost | H_vv(21) g_v(6) Q(36) E_vi(6 PI)] (bundle view rows);
    // the per-camera constant chain-rule transforms T_c are applied here, once per camera.
    void assemble_shared(const double* cam_sums, const double* x_host, std::vector<double>& Hss, std::vector<double>& gs) const {
        const int NE = S.NE, NC = S.NC, PI = S.PI, NL = S.NL;
        auto idx = [NL](int a, int b) { if (a > b) std::swap(a, b); return a * NL - a * (a - 1) / 2 + (b - a); };
        Hss.assign((size_t)ns * ns, 0.0); gs.assign(ns, 0.0);
        for (int c = 0; c < S.n_cams; ++c) {
            const double* sums = &cam_sums[(size_t)c * S.NV];
            const PB& pi = pbs[pb_intr(c)];
            int pose_idx[6]; bool pose_free = false;
            double Tc[36];
            if (S.cam_pose_kind) {
                const PB& pq = pbs[pb_camq(c)]; const PB& pt = pbs[pb_camt(c)];
                pose_free = !pq.constant;
                for (int k = 0; k < 3; ++k) { pose_idx[k] = pq.toff + k; pose_idx[3 + k] = pt.toff + k; }
                CamConst cc; cam_const_from_intr(x_host + pi.off, S.model, cc);
                if (S.cam_pose_kind == 1) cam_transform_extrinsics(x_host + pt.off, cc.Rs, Tc);
                else cam_transform_bundle(x_host + pq.off, cc.Rs, Tc);
            }
            const bool intr_free = !pi.constant && PI > 0;
            if (pose_free) {
                double Q[36];  // T_c^T N_xixi
                for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += Tc[6 * k + i] * sums[idx(k, j)]; Q[6 * i + j] = a; }
                for (int i = 0; i < 6; ++i) {
                    for (int j = 0; j < 6; ++j) { double a = 0; for (int k = 0; k < 6; ++k) a += Q[6 * i + k] * Tc[6 * k + j]; Hss[(size_t)pose_idx[i] * ns + pose_idx[j]] += a; }
                    double g = 0; for (int k = 0; k < 6; ++k) g += Tc[6 * k + i] * sums[idx(k, NC)];
                    gs[pose_idx[i]] += g;
                    if (intr_free) for (int j = 0; j < PI; ++j) {
                        double a = 0; for (int k = 0; k < 6; ++k) a += Tc[6 * k + i] * sums[idx(k, 6 + j)];
                        Hss[(size_t)pose_idx[i] * ns + pi.toff + j] += a; Hss[(size_t)(pi.toff + j) * ns + pose_idx[i]] += a;
                    }
                }
            }
            if (intr_free) for (int i = 0; i < PI; ++i) {
                gs[pi.toff + i] += sums[idx(6 + i, NC)];
                for (int j = 0; j < PI; ++j) Hss[(size_t)(pi.toff + i) * ns + pi.toff + j] += sums[idx(6 + i, 6 + j)];
            }
            if (S.kind == CAL_KIND_BUNDLE && S.view_free_global) {
                const PB& bq = pbs[pb_bq()]; const PB& bt = pbs[pb_bt()];
                int bidx[6]; for (int k = 0; k < 3; ++k) { bidx[k] = bq.toff + k; bidx[3 + k] = bt.toff + k; }
                const double* Hvv = sums + NE + 1; const double* gv = Hvv + 21; const double* Qs = gv + 6; const double* Evi = Qs + 36;
                int o = 0;
                for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { const double a = Hvv[o++]; Hss[(size_t)bidx[i] * ns + bidx[j]] += a; if (i != j) Hss[(size_t)bidx[j] * ns + bidx[i]] += a; }
                for (int i = 0; i < 6; ++i) gs[bidx[i]] += gv[i];
                if (pose_free) for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
                    double a = 0; for (int k = 0; k < 6; ++k) a += Qs[6 * i + k] * Tc[6 * k + j];
                    Hss[(size_t)bidx[i] * ns + pose_idx[j]] += a; Hss[(size_t)pose_idx[j] * ns + bidx[i]] += a;
                }
                if (intr_free) for (int i = 0; i < 6; ++i) for (int j = 0; j < PI; ++j) {
                    const double a = Evi[PI * i + j];
                    Hss[(size_t)bidx[i] * ns + pi.toff + j] += a; Hss[(size_t)(pi.toff + j) * ns + bidx[i]] += a;
                }
            }
        }
    }

    // Dense tangent-space system (canonical order) = shared block + per-view 6x6 blocks + their couplings,
    // from the per-view sums H_pp [n_views][36], g_p [n_views][6] and the per-block couplings
    // E_vc [36][n_blk] (view x camera pose), E_vi [6 PI][n_blk] (view x intrinsics); views are given as a CSR
    // of device block ids.  Only used for cal_refine_eval's dense output and the dense covariance path.
    void assemble_dense(const std::vector<double>& Hss, const std::vector<double>& gs, const double* Hpp, const double* gp,
                        const double* Evc, const double* Evi, int64_t nblk, const char* view_free, const int32_t* view_blk_off,
                        const int32_t* view_blk_idx, const int32_t* blk_cam, std::vector<double>& H, std::vector<double>& g) const {
        const int n = n_tan, PI = S.PI;
        H.assign((size_t)n * n, 0.0); g.assign(n, 0.0);
        for (int i = 0; i < ns; ++i) { g[i] = gs[i]; for (int j = 0; j < ns; ++j) H[(size_t)i * n + j] = Hss[(size_t)i * ns + j]; }
        for (int v = 0; v < S.n_views; ++v) {
            if (!view_free[v]) continue;
            const PB& q = pbs[pb_viewq(v)]; const PB& t = pbs[pb_viewt(v)];
            int vi[6]; for (int k = 0; k < 3; ++k) { vi[k] = q.toff + k; vi[3 + k] = t.toff + k; }
            for (int i = 0; i < 6; ++i) { g[vi[i]] = gp[(size_t)v * 6 + i]; for (int j = 0; j < 6; ++j) H[(size_t)vi[i] * n + vi[j]] = Hpp[(size_t)v * 36 + 6 * i + j]; }
            for (int k = view_blk_off[v]; k < view_blk_off[v + 1]; ++k) {
                const int64_t b = view_blk_idx[k]; const int cam = blk_cam[b];
                const PB& pi = pbs[pb_intr(cam)];
                if (S.cam_pose_kind == 1 && !pbs[pb_camq(cam)].constant) {
                    const int cq = pbs[pb_camq(cam)].toff, ct = pbs[pb_camt(cam)].toff;
                    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
                        const int col = j < 3 ? cq + j : ct + j - 3; const double a = Evc[(size_t)(6 * i + j) * nblk + b];
                        H[(size_t)vi[i] * n + col] += a; H[(size_t)col * n + vi[i]] += a;
                    }
                }
                if (!pi.constant && PI > 0) for (int i = 0; i < 6; ++i) for (int j = 0; j < PI; ++j) {
                    const double a = Evi[(size_t)(PI * i + j) * nblk + b];
                    H[(size_t)vi[i] * n + pi.toff + j] += a; H[(size_t)(pi.toff + j) * n + vi[i]] += a;
                }
            }
        }
    }
};

}  // namespace calk
