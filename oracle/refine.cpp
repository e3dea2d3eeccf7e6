// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).
//
// CPU restatement of the reference's reprojection refinement problems:
//   optimize_intrinsics  src/estimation/optim/intrinsics.cpp:14-120
//   optimize_extrinsics  src/estimation/optim/extrinsics.cpp:16-196
//   optimize_bundle      src/estimation/optim/bundle.cpp:20-170
// with the residual functors of src/estimation/residuals/{intrinsicresidual,
// extrinsicsresidual,bundleresidual}.h evaluated on forward-mode duals (the
// arithmetic ceres::AutoDiffCostFunction performs), the per-block Huber loss
// and corrector of Ceres, QuaternionManifold / SubsetManifold
// tangent projection (B.4), the LM loop of lm.hpp (B.3) and the covariance
// assembly of src/estimation/detail/ceresutils.h:69-126 (B.5).
#include <omp.h>

#include <cstdio>
#include <memory>
#include <type_traits>

#include "lm.hpp"
#include "oracle_api.h"
#include "oracle_math.hpp"

namespace orc {

enum PBType { PB_EUCLID = 0, PB_QUAT = 1, PB_INTR = 2 };
struct PB {
    int off, size, tsize, type;
    bool constant;
    int toff;  // canonical tangent offset, -1 if constant
};

struct Problem {
    const orc_problem_desc* d = nullptr;
    int P = 10;       // intrinsic block size (pinhole.h:118, scheimpflug.h:235)
    int n_amb = 0, n_tan = 0;
    std::vector<PB> pbs;  // in get_param_blocks() order == layout of x == covariance order
    // internal order used by the LM: shared dims first, then 6 per view
    int ns = 0, nv = 0;
    std::vector<int> shared_can;  // shared idx -> canonical tangent idx
    std::vector<int> can_to_int;  // canonical tangent idx -> internal idx
    std::vector<char> view_free;
    std::vector<std::vector<int64_t>> view_blocks;
    bool constrained = false;

    int pb_intr(int c) const { return d->kind == ORC_KIND_INTRINSICS ? 0 : c; }
    // intrinsics.cpp:34-50 : [intr][quat_v...][tran_v...]
    // extrinsics.cpp:50-69 : [intr_c...][cam quat...][cam tran...][tgt quat...][tgt tran...]
    // bundle.cpp:48-68     : [intr_c...][g quat...][g tran...][b quat][b tran]
    int pb_viewq(int v) const { return d->kind == ORC_KIND_INTRINSICS ? 1 + v : 3 * d->n_cams + v; }
    int pb_viewt(int v) const { return d->kind == ORC_KIND_INTRINSICS ? 1 + d->n_views + v : 3 * d->n_cams + d->n_views + v; }
    int pb_camq(int c) const { return d->n_cams + c; }
    int pb_camt(int c) const { return 2 * d->n_cams + c; }
    int pb_bq() const { return 3 * d->n_cams; }
    int pb_bt() const { return 3 * d->n_cams + 1; }
    int n_int() const { return ns + 6 * nv; }
};

static void add_pb(Problem& p, int size, int type, bool constant, bool opt_skew) {
    PB b; b.off = p.n_amb; b.size = size; b.type = type; b.constant = constant;
    b.tsize = type == PB_QUAT ? 3 : (type == PB_INTR && !opt_skew ? size - 1 : size);
    b.toff = -1;
    p.n_amb += size;
    p.pbs.push_back(b);
}

static Problem build_problem(const orc_problem_desc* d) {
    Problem p; p.d = d;
    p.P = d->model == ORC_MODEL_SCHEIMPFLUG_BC5 ? 12 : 10;
    const bool sk = d->optimize_skew != 0;
    if (d->kind == ORC_KIND_INTRINSICS) {
        add_pb(p, p.P, PB_INTR, false, sk);
        for (int v = 0; v < d->n_views; ++v) add_pb(p, 4, PB_QUAT, false, sk);
        for (int v = 0; v < d->n_views; ++v) add_pb(p, 3, PB_EUCLID, false, sk);
        p.constrained = true;  // intrinsics.cpp:81-82 lower bounds on fx, fy
    } else if (d->kind == ORC_KIND_EXTRINSICS) {
        // extrinsics.cpp:110-150
        const bool oi = d->optimize_intrinsics, oe = d->optimize_extrinsics;
        for (int c = 0; c < d->n_cams; ++c) add_pb(p, p.P, PB_INTR, !oi, sk);
        for (int c = 0; c < d->n_cams; ++c) add_pb(p, 4, PB_QUAT, !oe || c == 0, sk);
        for (int c = 0; c < d->n_cams; ++c) add_pb(p, 3, PB_EUCLID, !oe || c == 0, sk);
        for (int v = 0; v < d->n_views; ++v) add_pb(p, 4, PB_QUAT, oi && v == 0, sk);
        for (int v = 0; v < d->n_views; ++v) add_pb(p, 3, PB_EUCLID, oi && v == 0, sk);
        p.constrained = oi;
    } else {
        // bundle.cpp:98-131
        const bool oi = d->optimize_intrinsics, oh = d->optimize_hand_eye, ot = d->optimize_target_pose;
        for (int c = 0; c < d->n_cams; ++c) add_pb(p, p.P, PB_INTR, !oi, sk);
        for (int c = 0; c < d->n_cams; ++c) add_pb(p, 4, PB_QUAT, !oh, sk);
        for (int c = 0; c < d->n_cams; ++c) add_pb(p, 3, PB_EUCLID, !oh, sk);
        add_pb(p, 4, PB_QUAT, !ot, sk);
        add_pb(p, 3, PB_EUCLID, !ot, sk);
        p.constrained = oi;
    }
    for (auto& b : p.pbs) if (!b.constant) { b.toff = p.n_tan; p.n_tan += b.tsize; }
    // internal order
    p.can_to_int.assign(p.n_tan, -1);
    const bool has_views = d->kind != ORC_KIND_BUNDLE;
    p.nv = has_views ? d->n_views : 0;
    std::vector<char> is_view_pb(p.pbs.size(), 0);
    if (has_views) for (int v = 0; v < d->n_views; ++v) { is_view_pb[p.pb_viewq(v)] = 1; is_view_pb[p.pb_viewt(v)] = 1; }
    for (size_t i = 0; i < p.pbs.size(); ++i) {
        const PB& b = p.pbs[i];
        if (b.constant || is_view_pb[i]) continue;
        for (int k = 0; k < b.tsize; ++k) { p.can_to_int[b.toff + k] = static_cast<int>(p.shared_can.size()); p.shared_can.push_back(b.toff + k); }
    }
    p.ns = static_cast<int>(p.shared_can.size());
    p.view_free.assign(p.nv, 0);
    for (int v = 0; v < p.nv; ++v) {
        const PB& q = p.pbs[p.pb_viewq(v)]; const PB& t = p.pbs[p.pb_viewt(v)];
        if (q.constant) continue;
        p.view_free[v] = 1;
        for (int k = 0; k < 3; ++k) { p.can_to_int[q.toff + k] = p.ns + 6 * v + k; p.can_to_int[t.toff + k] = p.ns + 6 * v + 3 + k; }
    }
    if (has_views) {
        p.view_blocks.assign(p.nv, {});
        for (int64_t b = 0; b < d->n_blocks; ++b) p.view_blocks[d->kind == ORC_KIND_INTRINSICS ? b : d->block_view[b]].push_back(b);
    }
    return p;
}

// ---------------------------------------------------------------------------
// Residual functors on duals.  Jet layout = AutoDiffCostFunction parameter
// order: intrinsics <4,3,P>, extrinsics <4,3,4,3,P>, bundle <4,3,4,3,P>.
// ---------------------------------------------------------------------------
template <class T, int MODEL>
static inline void project_model(const T* intr, const T* Pc, T& u, T& v) {
    if (MODEL == ORC_MODEL_SCHEIMPFLUG_BC5) scheimpflug_project(intr, Pc, u, v);
    else pinhole_project(intr, Pc, u, v);
}

// Evaluates residual block b: r (2n) and ambient Jacobian Jamb (2n x N) where
// N is the Jet width; pb_ids / jet_off describe the blocks in Jet order.
template <int KIND, int MODEL>
struct BlockEval {
    static constexpr int P = MODEL == ORC_MODEL_SCHEIMPFLUG_BC5 ? 12 : 10;
    static constexpr int NPOSE = KIND == ORC_KIND_INTRINSICS ? 7 : 14;
    static constexpr int N = NPOSE + P;
    using D = Dual<N>;
    static void run(const Problem& pr, const double* x, int64_t b, bool jac, std::vector<double>& r,
                    std::vector<double>& Jamb, int* pb_ids /*5*/, int* n_pb) {
        const orc_problem_desc* d = pr.d;
        const int cam = d->block_cam[b];
        const int64_t o0 = d->block_offset[b], o1 = d->block_offset[b + 1];
        const int n = static_cast<int>(o1 - o0);
        r.resize(2 * n);
        if (jac) Jamb.resize(static_cast<size_t>(2 * n) * N);
        if (KIND == ORC_KIND_INTRINSICS) {
            const int v = static_cast<int>(b);
            pb_ids[0] = pr.pb_viewq(v); pb_ids[1] = pr.pb_viewt(v); pb_ids[2] = 0; *n_pb = 3;
        } else if (KIND == ORC_KIND_EXTRINSICS) {
            const int v = d->block_view[b];
            pb_ids[0] = pr.pb_camq(cam); pb_ids[1] = pr.pb_camt(cam); pb_ids[2] = pr.pb_viewq(v); pb_ids[3] = pr.pb_viewt(v);
            pb_ids[4] = pr.pb_intr(cam); *n_pb = 5;
        } else {
            pb_ids[0] = pr.pb_bq(); pb_ids[1] = pr.pb_bt(); pb_ids[2] = pr.pb_camq(cam); pb_ids[3] = pr.pb_camt(cam);
            pb_ids[4] = pr.pb_intr(cam); *n_pb = 5;
        }
        if (jac) eval<D>(pr, x, b, cam, o0, n, pb_ids, *n_pb, r.data(), Jamb.data());
        else eval<double>(pr, x, b, cam, o0, n, pb_ids, *n_pb, r.data(), nullptr);
    }
    template <class T>
    static T mk(double v, int k) {
        if constexpr (std::is_same<T, double>::value) { (void)k; return v; } else { return T::var(v, k); }
    }
    template <class T>
    static void eval(const Problem& pr, const double* x, int64_t b, int cam, int64_t o0, int n, const int* pb_ids,
                     int n_pb, double* r, double* Jamb) {
        (void)cam;
        const orc_problem_desc* d = pr.d;
        T par[N]; int k = 0;
        for (int i = 0; i < n_pb; ++i) { const PB& pb = pr.pbs[pb_ids[i]]; for (int j = 0; j < pb.size; ++j, ++k) par[k] = mk<T>(x[pb.off + j], k); }
        T Rct[9], tct[3];
        const T* intr;
        if (KIND == ORC_KIND_INTRINSICS) {
            // intrinsicresidual.h:22-24
            quat_to_rotmat(par, Rct); tct[0] = par[4]; tct[1] = par[5]; tct[2] = par[6];
            intr = par + 7;
        } else if (KIND == ORC_KIND_EXTRINSICS) {
            // extrinsicsresidual.h:14-20,30-33: c_se3_t = c_se3_r * r_se3_t
            T Rcr[9], Rrt[9]; quat_to_rotmat(par, Rcr); quat_to_rotmat(par + 7, Rrt);
            se3_product(Rcr, par + 4, Rrt, par + 11, Rct, tct);
            intr = par + 14;
        } else {
            // bundleresidual.h:15-27,38-43: c_se3_t = inv(g_se3_c) * inv(b_se3_g) * b_se3_t
            T Rbt[9], Rgc[9]; quat_to_rotmat(par, Rbt); quat_to_rotmat(par + 7, Rgc);
            T Rbg[9], tbg[3];
            for (int i = 0; i < 9; ++i) Rbg[i] = T(d->block_b_se3_g[12 * b + i]);
            for (int i = 0; i < 3; ++i) tbg[i] = T(d->block_b_se3_g[12 * b + 9 + i]);
            T Rcg[9], tcg[3], Rgb[9], tgb[3], Rcb[9], tcb[3];
            invert_transform(Rgc, par + 11, Rcg, tcg);
            invert_transform(Rbg, tbg, Rgb, tgb);
            se3_product(Rcg, tcg, Rgb, tgb, Rcb, tcb);
            se3_product(Rcb, tcb, Rbt, par + 4, Rct, tct);
            intr = par + 14;
        }
        for (int i = 0; i < n; ++i) {
            const double X = d->obj_x[o0 + i], Y = d->obj_y[o0 + i];
            // point = R * (X, Y, 0) + t   (intrinsicresidual.h:28-29)
            T Pc[3];
            for (int a = 0; a < 3; ++a) Pc[a] = Rct[3 * a] * X + Rct[3 * a + 1] * Y + Rct[3 * a + 2] * 0.0 + tct[a];
            T u, v; project_model<T, MODEL>(intr, Pc, u, v);
            T ru = u - d->img_u[o0 + i], rv = v - d->img_v[o0 + i];
            r[2 * i] = scalar(ru); r[2 * i + 1] = scalar(rv);
            if constexpr (!std::is_same<T, double>::value) {
                for (int c = 0; c < N; ++c) { Jamb[static_cast<size_t>(2 * i) * N + c] = ru.d[c]; Jamb[static_cast<size_t>(2 * i + 1) * N + c] = rv.d[c]; }
            }
        }
    }
};

// QuaternionManifold::PlusJacobian (Ceres manifold.h), row-major 4x3
static inline void quat_plus_jacobian(const double* q, double* J) {
    J[0] = -q[1]; J[1] = -q[2]; J[2] = -q[3];
    J[3] = q[0];  J[4] = q[3];  J[5] = -q[2];
    J[6] = -q[3]; J[7] = q[0];  J[8] = q[1];
    J[9] = q[2];  J[10] = -q[1]; J[11] = q[0];
}
// QuaternionManifold::Plus
static inline void quat_plus(const double* q, const double* dl, double* out) {
    const double nd = std::sqrt(dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
    if (nd == 0.0) { for (int i = 0; i < 4; ++i) out[i] = q[i]; return; }
    const double sd = std::sin(nd) / nd;
    const double dq[4] = {std::cos(nd), sd * dl[0], sd * dl[1], sd * dl[2]};
    // QuaternionProduct(dq, q)
    out[0] = dq[0] * q[0] - dq[1] * q[1] - dq[2] * q[2] - dq[3] * q[3];
    out[1] = dq[0] * q[1] + dq[1] * q[0] + dq[2] * q[3] - dq[3] * q[2];
    out[2] = dq[0] * q[2] - dq[1] * q[3] + dq[2] * q[0] + dq[3] * q[1];
    out[3] = dq[0] * q[3] + dq[1] * q[2] - dq[2] * q[1] + dq[3] * q[0];
}

// HuberLoss::Evaluate (Ceres loss_function.cc)
static inline void huber(double a, double s, double* rho) {
    const double b = a * a;
    if (s > b) { const double r = std::sqrt(s); rho[0] = 2.0 * a * r - b; rho[1] = std::max(DBL_MIN, a / r); rho[2] = -rho[1] / (2.0 * s); }
    else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

struct NormalEq {
    std::vector<double> Hss, gs;       // ns x ns, ns
    std::vector<double> Hpp, Hps, gp;  // nv x 36, nv x 6 x ns, nv x 6
    double cost = 0;
};

// One residual block -> tangent-space, loss-corrected local system.
struct LocalSys {
    std::vector<int> cols;       // internal tangent indices
    std::vector<double> JtJ, Jtr;
    double rho0 = 0, ssr = 0;
};

template <int KIND, int MODEL>
static void block_local(const Problem& pr, const double* x, int64_t b, bool jac, LocalSys& ls,
                        std::vector<double>& r, std::vector<double>& Jamb, std::vector<double>& Jt) {
    using BE = BlockEval<KIND, MODEL>;
    int pb_ids[5], n_pb = 0;
    BE::run(pr, x, b, jac, r, Jamb, pb_ids, &n_pb);
    const int m = static_cast<int>(r.size());
    double s = 0; for (int i = 0; i < m; ++i) s += r[i] * r[i];
    ls.ssr = s;
    double rho[3] = {s, 1.0, 0.0};
    if (pr.d->huber_delta > 0) huber(pr.d->huber_delta, s, rho);  // intrinsics.cpp:70-71
    ls.rho0 = rho[0];
    if (!jac) return;
    // tangent projection (manifold plus-Jacobians), constant blocks dropped
    ls.cols.clear();
    for (int i = 0; i < n_pb; ++i) { const PB& pb = pr.pbs[pb_ids[i]]; if (pb.constant) continue; for (int k = 0; k < pb.tsize; ++k) ls.cols.push_back(pr.can_to_int[pb.toff + k]); }
    const int nc = static_cast<int>(ls.cols.size());
    Jt.assign(static_cast<size_t>(m) * nc, 0.0);
    const int N = BE::N;
    int joff = 0, tcol = 0;
    for (int i = 0; i < n_pb; ++i) {
        const PB& pb = pr.pbs[pb_ids[i]];
        if (!pb.constant) {
            if (pb.type == PB_QUAT) {
                double PJ[12]; quat_plus_jacobian(x + pb.off, PJ);
                for (int row = 0; row < m; ++row) for (int k = 0; k < 3; ++k) {
                    double a = 0; for (int j = 0; j < 4; ++j) a += Jamb[static_cast<size_t>(row) * N + joff + j] * PJ[3 * j + k];
                    Jt[static_cast<size_t>(row) * nc + tcol + k] = a;
                }
            } else if (pb.type == PB_INTR && pb.tsize == pb.size - 1) {
                // SubsetManifold(size, {idx_skew = 4})
                for (int row = 0; row < m; ++row) { int k = 0; for (int j = 0; j < pb.size; ++j) { if (j == 4) continue; Jt[static_cast<size_t>(row) * nc + tcol + k] = Jamb[static_cast<size_t>(row) * N + joff + j]; ++k; } }
            } else {
                for (int row = 0; row < m; ++row) for (int j = 0; j < pb.size; ++j) Jt[static_cast<size_t>(row) * nc + tcol + j] = Jamb[static_cast<size_t>(row) * N + joff + j];
            }
            tcol += pb.tsize;
        }
        joff += pb.size;
    }
    // Corrector with rho'' <= 0: r <- sqrt(rho') r, J <- sqrt(rho') J, i.e. weight rho'
    const double w = rho[1];
    ls.JtJ.assign(static_cast<size_t>(nc) * nc, 0.0); ls.Jtr.assign(nc, 0.0);
    for (int row = 0; row < m; ++row) {
        const double* jr = &Jt[static_cast<size_t>(row) * nc];
        for (int a = 0; a < nc; ++a) { const double ja = jr[a]; ls.Jtr[a] += ja * r[row]; for (int c = a; c < nc; ++c) ls.JtJ[static_cast<size_t>(a) * nc + c] += ja * jr[c]; }
    }
    for (int a = 0; a < nc; ++a) { ls.Jtr[a] *= w; for (int c = a; c < nc; ++c) { ls.JtJ[static_cast<size_t>(a) * nc + c] *= w; ls.JtJ[static_cast<size_t>(c) * nc + a] = ls.JtJ[static_cast<size_t>(a) * nc + c]; } }
}

static void block_local_dispatch(const Problem& pr, const double* x, int64_t b, bool jac, LocalSys& ls,
                                 std::vector<double>& r, std::vector<double>& Jamb, std::vector<double>& Jt) {
    const int k = pr.d->kind, m = pr.d->model;
#define ORC_CASE(K, M) if (k == K && m == M) { block_local<K, M>(pr, x, b, jac, ls, r, Jamb, Jt); return; }
    ORC_CASE(ORC_KIND_INTRINSICS, ORC_MODEL_PINHOLE_BC5)
    ORC_CASE(ORC_KIND_INTRINSICS, ORC_MODEL_SCHEIMPFLUG_BC5)
    ORC_CASE(ORC_KIND_EXTRINSICS, ORC_MODEL_PINHOLE_BC5)
    ORC_CASE(ORC_KIND_EXTRINSICS, ORC_MODEL_SCHEIMPFLUG_BC5)
    ORC_CASE(ORC_KIND_BUNDLE, ORC_MODEL_PINHOLE_BC5)
    ORC_CASE(ORC_KIND_BUNDLE, ORC_MODEL_SCHEIMPFLUG_BC5)
#undef ORC_CASE
}

static void scatter(const Problem& pr, const LocalSys& ls, std::vector<double>& Hss, std::vector<double>& gs, NormalEq& ne) {
    const int nc = static_cast<int>(ls.cols.size()), ns = pr.ns;
    for (int a = 0; a < nc; ++a) {
        const int ia = ls.cols[a];
        if (ia < ns) gs[ia] += ls.Jtr[a];
        else ne.gp[ia - ns] += ls.Jtr[a];
        for (int c = 0; c < nc; ++c) {
            const int ic = ls.cols[c];
            const double val = ls.JtJ[static_cast<size_t>(a) * nc + c];
            if (ia < ns && ic < ns) Hss[static_cast<size_t>(ia) * ns + ic] += val;
            else if (ia >= ns && ic >= ns) { const int v = (ia - ns) / 6; ne.Hpp[static_cast<size_t>(v) * 36 + ((ia - ns) % 6) * 6 + (ic - ns) % 6] += val; }
            else if (ia >= ns && ic < ns) { const int v = (ia - ns) / 6; ne.Hps[(static_cast<size_t>(v) * 6 + (ia - ns) % 6) * ns + ic] += val; }
        }
    }
}

static bool evaluate(const Problem& pr, const double* x, bool jac, NormalEq& ne, int num_threads) {
    const int ns = pr.ns, nv = pr.nv;
    if (jac) {
        ne.Hss.assign(static_cast<size_t>(ns) * ns, 0.0); ne.gs.assign(ns, 0.0);
        ne.Hpp.assign(static_cast<size_t>(nv) * 36, 0.0); ne.Hps.assign(static_cast<size_t>(nv) * 6 * ns, 0.0); ne.gp.assign(static_cast<size_t>(nv) * 6, 0.0);
    }
    const int nt = num_threads > 0 ? num_threads : omp_get_max_threads();
    std::vector<std::vector<double>> tH(nt), tg(nt);
    std::vector<double> tcost(nt, 0.0);
    const int64_t nb = pr.d->n_blocks;
#pragma omp parallel num_threads(nt)
    {
        const int tid = omp_get_thread_num();
        std::vector<double>& H = tH[tid]; std::vector<double>& g = tg[tid];
        if (jac) { H.assign(static_cast<size_t>(ns) * ns, 0.0); g.assign(ns, 0.0); }
        LocalSys ls; std::vector<double> r, Jamb, Jt;
        double c = 0;
        if (nv > 0) {
#pragma omp for schedule(static)
            for (int v = 0; v < nv; ++v)
                for (int64_t b : pr.view_blocks[v]) { block_local_dispatch(pr, x, b, jac, ls, r, Jamb, Jt); c += 0.5 * ls.rho0; if (jac) scatter(pr, ls, H, g, ne); }
        } else {
#pragma omp for schedule(static)
            for (int64_t b = 0; b < nb; ++b) { block_local_dispatch(pr, x, b, jac, ls, r, Jamb, Jt); c += 0.5 * ls.rho0; if (jac) scatter(pr, ls, H, g, ne); }
        }
        tcost[tid] = c;
    }
    ne.cost = 0;
    for (int t = 0; t < nt; ++t) {
        ne.cost += tcost[t];
        if (jac && !tH[t].empty()) { for (size_t i = 0; i < ne.Hss.size(); ++i) ne.Hss[i] += tH[t][i]; for (int i = 0; i < ns; ++i) ne.gs[i] += tg[t][i]; }
    }
    return std::isfinite(ne.cost);
}

// ---------------------------------------------------------------------------
// LMProblem adapter
// ---------------------------------------------------------------------------
struct ReprojLM final : LMProblem {
    Problem pr; NormalEq ne; int threads = 0; bool force_dense = false;
    int n_amb() const override { return pr.n_amb; }
    int n_int() const override { return pr.n_int(); }
    bool constrained() const override { return pr.constrained; }
    bool eval(const double* x, double* cost, bool jac) override {
        if (jac) { bool ok = evaluate(pr, x, true, ne, threads); *cost = ne.cost; return ok; }
        NormalEq tmp; bool ok = evaluate(pr, x, false, tmp, threads); *cost = tmp.cost; return ok;
    }
    void diag(double* out) const override {
        for (int i = 0; i < pr.ns; ++i) out[i] = ne.Hss[static_cast<size_t>(i) * pr.ns + i];
        for (int v = 0; v < pr.nv; ++v) for (int k = 0; k < 6; ++k) out[pr.ns + 6 * v + k] = ne.Hpp[static_cast<size_t>(v) * 36 + 7 * k];
    }
    void grad(double* out) const override {
        for (int i = 0; i < pr.ns; ++i) out[i] = ne.gs[i];
        for (int i = 0; i < 6 * pr.nv; ++i) out[pr.ns + i] = ne.gp[i];
    }
    bool solve_dense(const double* s, const double* D2, double* y) {
        const int n = n_int(), ns = pr.ns;
        std::vector<double> A(static_cast<size_t>(n) * n, 0.0);
        for (int i = 0; i < ns; ++i) for (int j = 0; j < ns; ++j) A[static_cast<size_t>(i) * n + j] = ne.Hss[static_cast<size_t>(i) * ns + j] * s[i] * s[j];
        for (int v = 0; v < pr.nv; ++v) {
            for (int a = 0; a < 6; ++a) {
                const int ia = ns + 6 * v + a;
                for (int c = 0; c < 6; ++c) A[static_cast<size_t>(ia) * n + ns + 6 * v + c] = ne.Hpp[static_cast<size_t>(v) * 36 + 6 * a + c] * s[ia] * s[ns + 6 * v + c];
                for (int j = 0; j < ns; ++j) { const double val = ne.Hps[(static_cast<size_t>(v) * 6 + a) * ns + j] * s[ia] * s[j]; A[static_cast<size_t>(ia) * n + j] = val; A[static_cast<size_t>(j) * n + ia] = val; }
            }
        }
        std::vector<double> g(n); grad(g.data());
        for (int i = 0; i < n; ++i) { A[static_cast<size_t>(i) * n + i] += D2[i]; y[i] = g[i] * s[i]; }
        if (!cholesky(A.data(), n)) return false;
        cholesky_solve(A.data(), n, y);
        return true;
    }
    bool solve(const double* s, const double* D2, double* y) override {
        if (force_dense) return solve_dense(s, D2, y);
        const int ns = pr.ns, nv = pr.nv;
        std::vector<double> S(static_cast<size_t>(ns) * ns), gr(ns);
        for (int i = 0; i < ns; ++i) { for (int j = 0; j < ns; ++j) S[static_cast<size_t>(i) * ns + j] = ne.Hss[static_cast<size_t>(i) * ns + j] * s[i] * s[j]; S[static_cast<size_t>(i) * ns + i] += D2[i]; gr[i] = ne.gs[i] * s[i]; }
        std::vector<double> Lv(static_cast<size_t>(nv) * 36), Bs(static_cast<size_t>(nv) * 6 * ns), gps(static_cast<size_t>(nv) * 6);
        bool ok = true;
        for (int v = 0; v < nv; ++v) {
            double* L = &Lv[static_cast<size_t>(v) * 36];
            const double* sv = s + ns + 6 * v;
            for (int a = 0; a < 6; ++a) { for (int c = 0; c < 6; ++c) L[6 * a + c] = ne.Hpp[static_cast<size_t>(v) * 36 + 6 * a + c] * sv[a] * sv[c]; L[7 * a] += D2[ns + 6 * v + a]; }
            if (!cholesky(L, 6)) { ok = false; break; }
            double* B = &Bs[static_cast<size_t>(v) * 6 * ns];
            for (int a = 0; a < 6; ++a) { for (int j = 0; j < ns; ++j) B[a * ns + j] = ne.Hps[(static_cast<size_t>(v) * 6 + a) * ns + j] * sv[a] * s[j]; gps[static_cast<size_t>(v) * 6 + a] = ne.gp[static_cast<size_t>(v) * 6 + a] * sv[a]; }
            if (!pr.view_free[v]) continue;
            // W = A^-1 B (column by column), S -= B^T W, gr -= B^T A^-1 gp
            double z[6]; for (int a = 0; a < 6; ++a) z[a] = gps[static_cast<size_t>(v) * 6 + a];
            cholesky_solve(L, 6, z);
            std::vector<int> nz; for (int j = 0; j < ns; ++j) { bool any = false; for (int a = 0; a < 6; ++a) any |= B[a * ns + j] != 0.0; if (any) nz.push_back(j); }
            std::vector<double> W(6 * nz.size());
            for (size_t jj = 0; jj < nz.size(); ++jj) { double col[6]; for (int a = 0; a < 6; ++a) col[a] = B[a * ns + nz[jj]]; cholesky_solve(L, 6, col); for (int a = 0; a < 6; ++a) W[a * nz.size() + jj] = col[a]; }
            for (size_t ii = 0; ii < nz.size(); ++ii) {
                const int i = nz[ii]; double gi = 0;
                for (int a = 0; a < 6; ++a) gi += B[a * ns + i] * z[a];
                gr[i] -= gi;
                for (size_t jj = 0; jj < nz.size(); ++jj) { double acc = 0; for (int a = 0; a < 6; ++a) acc += B[a * ns + i] * W[a * nz.size() + jj]; S[static_cast<size_t>(i) * ns + nz[jj]] -= acc; }
            }
        }
        if (!ok) return false;
        if (ns > 0) { if (!cholesky(S.data(), ns)) return false; cholesky_solve(S.data(), ns, gr.data()); }
        for (int i = 0; i < ns; ++i) y[i] = gr[i];
        for (int v = 0; v < nv; ++v) {
            double z[6]; const double* B = &Bs[static_cast<size_t>(v) * 6 * ns];
            for (int a = 0; a < 6; ++a) { double t = gps[static_cast<size_t>(v) * 6 + a]; for (int j = 0; j < ns; ++j) t -= B[a * ns + j] * y[j]; z[a] = t; }
            // constant views: H = 0, g = 0, damped diagonal -> y = 0
            cholesky_solve(&Lv[static_cast<size_t>(v) * 36], 6, z);
            for (int a = 0; a < 6; ++a) y[ns + 6 * v + a] = pr.view_free[v] ? z[a] : 0.0;
        }
        return true;
    }
    double quad(const double* s, const double* st) const override {
        const int ns = pr.ns; double q = 0;
        for (int i = 0; i < ns; ++i) { double row = 0; for (int j = 0; j < ns; ++j) row += ne.Hss[static_cast<size_t>(i) * ns + j] * s[j] * st[j]; q += row * s[i] * st[i]; }
        for (int v = 0; v < pr.nv; ++v) {
            const double* sv = s + ns + 6 * v; const double* tv = st + ns + 6 * v;
            for (int a = 0; a < 6; ++a) {
                double row = 0;
                for (int c = 0; c < 6; ++c) row += ne.Hpp[static_cast<size_t>(v) * 36 + 6 * a + c] * sv[c] * tv[c];
                double cross = 0;
                for (int j = 0; j < ns; ++j) cross += ne.Hps[(static_cast<size_t>(v) * 6 + a) * ns + j] * s[j] * st[j];
                q += (row + 2.0 * cross) * sv[a] * tv[a];
            }
        }
        return q;
    }
    void plus(const double* x, const double* delta, double* xp) const override {
        for (const PB& pb : pr.pbs) {
            if (pb.constant) { for (int j = 0; j < pb.size; ++j) xp[pb.off + j] = x[pb.off + j]; continue; }
            double dl[12];
            for (int k = 0; k < pb.tsize; ++k) dl[k] = delta[pr.can_to_int[pb.toff + k]];
            if (pb.type == PB_QUAT) quat_plus(x + pb.off, dl, xp + pb.off);
            else if (pb.type == PB_INTR) {
                int k = 0;
                for (int j = 0; j < pb.size; ++j) { if (pb.tsize == pb.size - 1 && j == 4) { xp[pb.off + j] = x[pb.off + j]; continue; } xp[pb.off + j] = x[pb.off + j] + dl[k++]; }
                // SetParameterLowerBound(fx, 0), (fy, 0): projection inside Program::Plus
                xp[pb.off + 0] = std::max(xp[pb.off + 0], 0.0); xp[pb.off + 1] = std::max(xp[pb.off + 1], 0.0);
            } else for (int j = 0; j < pb.size; ++j) xp[pb.off + j] = x[pb.off + j] + dl[j];
        }
    }
    // dense tangent-space H / g in canonical order
    void dense_canonical(std::vector<double>& H, std::vector<double>& g) const {
        const int n = pr.n_tan, ns = pr.ns;
        H.assign(static_cast<size_t>(n) * n, 0.0); g.assign(n, 0.0);
        std::vector<int> int_to_can(n_int(), -1);
        for (int c = 0; c < n; ++c) if (pr.can_to_int[c] >= 0) int_to_can[pr.can_to_int[c]] = c;
        for (int i = 0; i < ns; ++i) { g[int_to_can[i]] = ne.gs[i]; for (int j = 0; j < ns; ++j) H[static_cast<size_t>(int_to_can[i]) * n + int_to_can[j]] = ne.Hss[static_cast<size_t>(i) * ns + j]; }
        for (int v = 0; v < pr.nv; ++v) {
            if (!pr.view_free[v]) continue;
            for (int a = 0; a < 6; ++a) {
                const int ca = int_to_can[ns + 6 * v + a];
                g[ca] = ne.gp[static_cast<size_t>(v) * 6 + a];
                for (int c = 0; c < 6; ++c) H[static_cast<size_t>(ca) * n + int_to_can[ns + 6 * v + c]] = ne.Hpp[static_cast<size_t>(v) * 36 + 6 * a + c];
                for (int j = 0; j < ns; ++j) { const double val = ne.Hps[(static_cast<size_t>(v) * 6 + a) * ns + j]; H[static_cast<size_t>(ca) * n + int_to_can[j]] = val; H[static_cast<size_t>(int_to_can[j]) * n + ca] = val; }
            }
        }
    }
};

// compute_covariance (ceresutils.h:69-126) over ceres::Covariance
static bool covariance(ReprojLM& lm, const double* x, double* cov) {
    double c; if (!lm.eval(x, &c, true)) return false;
    const int n = lm.pr.n_tan, na = lm.pr.n_amb;
    std::vector<double> H, g; lm.dense_canonical(H, g);
    std::vector<double> L = H;
    if (n > 0 && !cholesky(L.data(), n)) return false;
    // C_tan = H^-1 column by column
    std::vector<double> C(static_cast<size_t>(n) * n, 0.0), e(n);
    for (int j = 0; j < n; ++j) { std::fill(e.begin(), e.end(), 0.0); e[j] = 1.0; cholesky_solve(L.data(), n, e.data()); for (int i = 0; i < n; ++i) C[static_cast<size_t>(i) * n + j] = e[i]; }
    // lift: cov_amb = P C P^T with P = blockdiag(plus jacobians)
    std::vector<double> Pm(static_cast<size_t>(na) * std::max(n, 1), 0.0);
    for (const PB& pb : lm.pr.pbs) {
        if (pb.constant) continue;
        if (pb.type == PB_QUAT) { double PJ[12]; quat_plus_jacobian(x + pb.off, PJ); for (int j = 0; j < 4; ++j) for (int k = 0; k < 3; ++k) Pm[static_cast<size_t>(pb.off + j) * n + pb.toff + k] = PJ[3 * j + k]; }
        else if (pb.type == PB_INTR && pb.tsize == pb.size - 1) { int k = 0; for (int j = 0; j < pb.size; ++j) { if (j == 4) continue; Pm[static_cast<size_t>(pb.off + j) * n + pb.toff + k] = 1.0; ++k; } }
        else for (int j = 0; j < pb.size; ++j) Pm[static_cast<size_t>(pb.off + j) * n + pb.toff + j] = 1.0;
    }
    std::vector<double> PC(static_cast<size_t>(na) * std::max(n, 1), 0.0);
    for (int i = 0; i < na; ++i) for (int k = 0; k < n; ++k) { const double pik = Pm[static_cast<size_t>(i) * n + k]; if (pik == 0.0) continue; for (int j = 0; j < n; ++j) PC[static_cast<size_t>(i) * n + j] += pik * C[static_cast<size_t>(k) * n + j]; }
    for (int i = 0; i < na; ++i) for (int j = 0; j < na; ++j) { double a = 0; for (int k = 0; k < n; ++k) { const double pjk = Pm[static_cast<size_t>(j) * n + k]; if (pjk != 0.0) a += PC[static_cast<size_t>(i) * n + k] * pjk; } cov[static_cast<size_t>(i) * na + j] = a; }
    return true;
}

}  // namespace orc

using namespace orc;

extern "C" {

int64_t orc_param_count(const orc_problem_desc* d) { return build_problem(d).n_amb; }
int64_t orc_tangent_count(const orc_problem_desc* d) { return build_problem(d).n_tan; }

int orc_refine_eval(const orc_problem_desc* d, const double* x, double* cost, double* g, double* H, int num_threads) {
    ReprojLM lm; lm.pr = build_problem(d); lm.threads = num_threads;
    const bool jac = g != nullptr || H != nullptr;
    double c = 0; const bool ok = lm.eval(x, &c, jac);
    if (cost) *cost = c;
    if (jac) {
        std::vector<double> Hd, gd; lm.dense_canonical(Hd, gd);
        if (g) std::memcpy(g, gd.data(), gd.size() * sizeof(double));
        if (H) std::memcpy(H, Hd.data(), Hd.size() * sizeof(double));
    }
    return ok ? 0 : 1;
}

int orc_block_ssr(const orc_problem_desc* d, const double* x, double* ssr, int num_threads) {
    Problem pr = build_problem(d);
    const int nt = num_threads > 0 ? num_threads : omp_get_max_threads();
#pragma omp parallel num_threads(nt)
    {
        LocalSys ls; std::vector<double> r, Jamb, Jt;
#pragma omp for schedule(static)
        for (int64_t b = 0; b < d->n_blocks; ++b) { block_local_dispatch(pr, x, b, false, ls, r, Jamb, Jt); ssr[b] = ls.ssr; }
    }
    return 0;
}

int orc_refine_solve(const orc_problem_desc* d, const orc_optim_options* o, double* x_inout, orc_optim_result* res,
                     double* cov, int force_dense) {
    ReprojLM lm; lm.pr = build_problem(d); lm.threads = o->num_threads; lm.force_dense = force_dense != 0;
    std::vector<double> x(x_inout, x_inout + lm.pr.n_amb);
    LMOptions lo; lo.max_iterations = o->max_iterations; lo.epsilon = o->epsilon; lo.verbose = o->verbose != 0;
    LMSummary s = lm_minimize(lm, lo, x);
    std::memcpy(x_inout, x.data(), x.size() * sizeof(double));
    std::memset(res, 0, sizeof *res);
    res->success = s.termination == 0; res->iterations = s.iterations; res->termination = s.termination;
    res->num_jac_evals = s.num_jac_evals; res->num_cost_evals = s.num_cost_evals;
    res->initial_cost = s.initial_cost; res->final_cost = s.final_cost;
    std::snprintf(res->report, sizeof res->report, "%s", brief_report(s).c_str());
    if (cov && o->compute_covariance) res->covariance_ok = covariance(lm, x.data(), cov) ? 1 : 0;
    return 0;
}

void orc_project(int model, const double* intr, const double* P, double* uv) {
    if (model == ORC_MODEL_SCHEIMPFLUG_BC5) scheimpflug_project(intr, P, uv[0], uv[1]);
    else pinhole_project(intr, P, uv[0], uv[1]);
}

}  // extern "C"
