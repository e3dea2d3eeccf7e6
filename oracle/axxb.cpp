// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).
//
// CPU restatement of optimize_handeye (reference src/estimation/optim/handeye.cpp:17-78):
// the AX = XB residual of src/estimation/residuals/handeyeresidual.h:18-54 on
// forward-mode duals, motion-pair construction of src/estimation/linear/
// handeyedlt.cpp:11-81 with the SO(3) helpers of include/calib/estimation/
// common/se3_utils.h:10-40, the per-pair Huber loss, and the LM loop of lm.hpp.
#include <omp.h>

#include <type_traits>

#include "lm.hpp"
#include "oracle_api.h"
#include "oracle_math.hpp"

namespace orc {

// Eigen::Quaternion<T>(Matrix3<T>) then Eigen::AngleAxis<T>(quaternion) — the
// conversion `Eigen::AngleAxis<T> axisangle(rot_s)` performs
// (handeyeresidual.h:33).
template <class T> static void rotmat_to_angle_axis(const T* R, T& angle, T* axis) {
    T q[4];
    T t = R[0] + R[4] + R[8];
    if (t > T(0.0)) {
        t = sqrt(t + 1.0); q[0] = 0.5 * t; t = 0.5 / t;
        q[1] = (R[7] - R[5]) * t; q[2] = (R[2] - R[6]) * t; q[3] = (R[3] - R[1]) * t;
    } else {
        int i = 0; if (R[4] > R[0]) i = 1; if (R[8] > R[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        q[1 + i] = 0.5 * t; t = 0.5 / t;
        q[0] = (R[3 * k + j] - R[3 * j + k]) * t;
        q[1 + j] = (R[3 * j + i] + R[3 * i + j]) * t;
        q[1 + k] = (R[3 * k + i] + R[3 * i + k]) * t;
    }
    T n = sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (scalar(n) != 0.0) {
        angle = 2.0 * atan2(n, abs(q[0]));
        if (q[0] < T(0.0)) n = -n;
        axis[0] = q[1] / n; axis[1] = q[2] / n; axis[2] = q[3] / n;
    } else {
        angle = T(0.0); axis[0] = T(1.0); axis[1] = T(0.0); axis[2] = T(0.0);
    }
}

// AxXbResidual::operator() (handeyeresidual.h:25-49)
template <class T> static void axxb_residual(const double* Ra, const double* Rb, const double* ta, const double* tb,
                                             const T* q, const T* t, T* res) {
    T Rx[9]; quat_to_rotmat(q, Rx);
    T A[9], B[9];
    for (int i = 0; i < 9; ++i) { A[i] = T(Ra[i]); B[i] = T(Rb[i]); }
    // rot_s = rot_a * rot_x * rot_b^T * rot_x^T
    T M1[9], M2[9], Rs[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M1[3 * i + j] = A[3 * i] * Rx[j] + A[3 * i + 1] * Rx[3 + j] + A[3 * i + 2] * Rx[6 + j];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M2[3 * i + j] = M1[3 * i] * B[3 * j] + M1[3 * i + 1] * B[3 * j + 1] + M1[3 * i + 2] * B[3 * j + 2];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rs[3 * i + j] = M2[3 * i] * Rx[3 * j] + M2[3 * i + 1] * Rx[3 * j + 1] + M2[3 * i + 2] * Rx[3 * j + 2];
    T angle, axis[3]; rotmat_to_angle_axis(Rs, angle, axis);
    // tra_e = (rot_a - I) tra_x - (rot_x tra_b - tra_a)
    for (int i = 0; i < 3; ++i) {
        T lhs = (A[3 * i] - (i == 0 ? 1.0 : 0.0)) * t[0] + (A[3 * i + 1] - (i == 1 ? 1.0 : 0.0)) * t[1] + (A[3 * i + 2] - (i == 2 ? 1.0 : 0.0)) * t[2];
        T rhs = Rx[3 * i] * tb[0] + Rx[3 * i + 1] * tb[1] + Rx[3 * i + 2] * tb[2] - ta[i];
        res[3 + i] = lhs - rhs;
    }
    res[0] = angle * axis[0]; res[1] = angle * axis[1]; res[2] = angle * axis[2];
}

static inline void quat_plus_jacobian7(const double* q, double* J) {
    J[0] = -q[1]; J[1] = -q[2]; J[2] = -q[3];
    J[3] = q[0];  J[4] = q[3];  J[5] = -q[2];
    J[6] = -q[3]; J[7] = q[0];  J[8] = q[1];
    J[9] = q[2];  J[10] = -q[1]; J[11] = q[0];
}

struct AxxbLM final : LMProblem {
    const orc_axxb_desc* d; int threads = 0;
    double H[36], g[6];
    int n_amb() const override { return 7; }
    int n_int() const override { return 6; }
    bool constrained() const override { return false; }
    bool eval(const double* x, double* cost, bool jac) override {
        const int nt = threads > 0 ? threads : omp_get_max_threads();
        std::vector<std::vector<double>> acc(nt, std::vector<double>(43, 0.0));
        double PJ[12]; quat_plus_jacobian7(x, PJ);
        const double hd = d->huber_delta;
#pragma omp parallel num_threads(nt)
        {
            std::vector<double>& a = acc[omp_get_thread_num()];
#pragma omp for schedule(static)
            for (int64_t p = 0; p < d->n_pairs; ++p) {
                const double* Ra = d->rot_a + 9 * p; const double* Rb = d->rot_b + 9 * p;
                const double* ta = d->tra_a + 3 * p; const double* tb = d->tra_b + 3 * p;
                double r[6], Jt[36];
                if (jac) {
                    using D = Dual<7>;
                    D par[7]; for (int k = 0; k < 7; ++k) par[k] = D::var(x[k], k);
                    D res[6]; axxb_residual<D>(Ra, Rb, ta, tb, par, par + 4, res);
                    for (int i = 0; i < 6; ++i) {
                        r[i] = res[i].v;
                        for (int k = 0; k < 3; ++k) { double s = 0; for (int j = 0; j < 4; ++j) s += res[i].d[j] * PJ[3 * j + k]; Jt[6 * i + k] = s; }
                        for (int k = 0; k < 3; ++k) Jt[6 * i + 3 + k] = res[i].d[4 + k];
                    }
                } else {
                    axxb_residual<double>(Ra, Rb, ta, tb, x, x + 4, r);
                }
                double s = 0; for (int i = 0; i < 6; ++i) s += r[i] * r[i];
                double rho0 = s, w = 1.0;
                if (hd > 0) { const double b = hd * hd; if (s > b) { const double rr = std::sqrt(s); rho0 = 2.0 * hd * rr - b; w = std::max(DBL_MIN, hd / rr); } }
                a[42] += 0.5 * rho0;
                if (jac) for (int i = 0; i < 6; ++i) for (int c = 0; c < 6; ++c) a[36 + c] += w * Jt[6 * i + c] * r[i];
                if (jac) for (int i = 0; i < 6; ++i) for (int c = 0; c < 6; ++c) for (int e = 0; e < 6; ++e) a[6 * c + e] += w * Jt[6 * i + c] * Jt[6 * i + e];
            }
        }
        double c = 0; if (jac) { for (double& v : H) v = 0; for (double& v : g) v = 0; }
        for (int t = 0; t < nt; ++t) { c += acc[t][42]; if (jac) { for (int i = 0; i < 36; ++i) H[i] += acc[t][i]; for (int i = 0; i < 6; ++i) g[i] += acc[t][36 + i]; } }
        *cost = c; return std::isfinite(c);
    }
    void diag(double* out) const override { for (int i = 0; i < 6; ++i) out[i] = H[7 * i]; }
    void grad(double* out) const override { for (int i = 0; i < 6; ++i) out[i] = g[i]; }
    bool solve(const double* s, const double* D2, double* y) override {
        double A[36]; for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) A[6 * i + j] = H[6 * i + j] * s[i] * s[j]; A[7 * i] += D2[i]; y[i] = g[i] * s[i]; }
        if (!cholesky(A, 6)) return false; cholesky_solve(A, 6, y); return true;
    }
    double quad(const double* s, const double* st) const override { double q = 0; for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) q += H[6 * i + j] * s[i] * st[i] * s[j] * st[j]; return q; }
    void plus(const double* x, const double* dl, double* xp) const override {
        const double nd = std::sqrt(dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
        if (nd == 0.0) { for (int i = 0; i < 4; ++i) xp[i] = x[i]; }
        else {
            const double sd = std::sin(nd) / nd; const double dq[4] = {std::cos(nd), sd * dl[0], sd * dl[1], sd * dl[2]};
            xp[0] = dq[0] * x[0] - dq[1] * x[1] - dq[2] * x[2] - dq[3] * x[3];
            xp[1] = dq[0] * x[1] + dq[1] * x[0] + dq[2] * x[3] - dq[3] * x[2];
            xp[2] = dq[0] * x[2] - dq[1] * x[3] + dq[2] * x[0] + dq[3] * x[1];
            xp[3] = dq[0] * x[3] + dq[1] * x[2] - dq[2] * x[1] + dq[3] * x[0];
        }
        for (int i = 0; i < 3; ++i) xp[4 + i] = x[4 + i] + dl[3 + i];
    }
};

// project_to_so3 (se3_utils.h:10-20): U diag(1,1,det(U V^T)) V^T of the SVD
static void project_to_so3(const double* R, double* out) {
    std::vector<double> A(R, R + 9), V, sv; jacobi_svd(A, 3, 3, V, sv);
    double U[9];
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) U[3 * i + j] = sv[j] > 0 ? A[3 * i + j] / sv[j] : 0.0;
    double Vt[9]; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Vt[3 * i + j] = V[3 * j + i];
    double UVt[9]; mat3_mul(U, Vt, UVt);
    if (mat3_det(UVt) < 0.0) {
        // flip the direction of the smallest singular value
        int m = 0; for (int j = 1; j < 3; ++j) if (sv[j] < sv[m]) m = j;
        for (int i = 0; i < 3; ++i) U[3 * i + m] = -U[3 * i + m];
        mat3_mul(U, Vt, UVt);
    }
    std::memcpy(out, UVt, sizeof UVt);
}
// log_so3 (se3_utils.h:28-42)
static void log_so3(const double* Rin, double* w) {
    double R[9]; project_to_so3(Rin, R);
    double c = (R[0] + R[4] + R[8] - 1.0) * 0.5; c = std::min(1.0, std::max(-1.0, c));
    const double th = std::acos(c);
    if (th < 1e-12) { w[0] = w[1] = w[2] = 0; return; }
    const double k = 0.5 / std::sin(th) * th;
    w[0] = (R[7] - R[5]) * k; w[1] = (R[2] - R[6]) * k; w[2] = (R[3] - R[1]) * k;
}

}  // namespace orc
using namespace orc;

extern "C" {

// known-answer hooks for tests/unit/se3_utils_test.cpp:10-30
void orc_project_to_so3(const double* R9, double* out9) { project_to_so3(R9, out9); }
void orc_log_so3(const double* R9, double* w3) { log_so3(R9, w3); }

int orc_axxb_eval(const orc_axxb_desc* d, const double* x7, double* cost, double* g6, double* H36, int num_threads) {
    AxxbLM lm; lm.d = d; lm.threads = num_threads;
    double c; const bool ok = lm.eval(x7, &c, g6 || H36);
    if (cost) *cost = c;
    if (g6) std::memcpy(g6, lm.g, sizeof lm.g);
    if (H36) std::memcpy(H36, lm.H, sizeof lm.H);
    return ok ? 0 : 1;
}

int orc_axxb_solve(const orc_axxb_desc* d, const orc_optim_options* o, double* x7, orc_optim_result* res, double* cov49) {
    AxxbLM lm; lm.d = d; lm.threads = o->num_threads;
    std::vector<double> x(x7, x7 + 7);
    LMOptions lo; lo.max_iterations = o->max_iterations; lo.epsilon = o->epsilon; lo.verbose = o->verbose != 0;
    LMSummary s = lm_minimize(lm, lo, x);
    std::memcpy(x7, x.data(), 7 * sizeof(double));
    std::memset(res, 0, sizeof *res);
    res->success = s.termination == 0; res->iterations = s.iterations; res->termination = s.termination;
    res->num_jac_evals = s.num_jac_evals; res->num_cost_evals = s.num_cost_evals;
    res->initial_cost = s.initial_cost; res->final_cost = s.final_cost;
    std::snprintf(res->report, sizeof res->report, "%s", brief_report(s).c_str());
    if (cov49 && o->compute_covariance) {
        double c; lm.eval(x.data(), &c, true);
        double L[36]; std::memcpy(L, lm.H, sizeof L);
        if (cholesky(L, 6)) {
            double C[36]; for (int j = 0; j < 6; ++j) { double e[6] = {0, 0, 0, 0, 0, 0}; e[j] = 1; cholesky_solve(L, 6, e); for (int i = 0; i < 6; ++i) C[6 * i + j] = e[i]; }
            double Pm[42] = {0}; double PJ[12]; quat_plus_jacobian7(x.data(), PJ);
            for (int j = 0; j < 4; ++j) for (int k = 0; k < 3; ++k) Pm[6 * j + k] = PJ[3 * j + k];
            for (int j = 0; j < 3; ++j) Pm[6 * (4 + j) + 3 + j] = 1.0;
            for (int i = 0; i < 7; ++i) for (int j = 0; j < 7; ++j) { double a = 0; for (int k = 0; k < 6; ++k) for (int l = 0; l < 6; ++l) a += Pm[6 * i + k] * C[6 * k + l] * Pm[6 * j + l]; cov49[7 * i + j] = a; }
            res->covariance_ok = 1;
        }
    }
    return 0;
}

int64_t orc_build_all_pairs(int64_t n, const double* bg, const double* ct, double min_angle_deg, double* rot_a,
                            double* rot_b, double* tra_a, double* tra_b) {
    // build_all_pairs(base, cam, min_angle_deg, reject_axis_parallel = true, eps = 1e-3)
    const double min_angle = min_angle_deg * M_PI / 180.0;
    int64_t cnt = 0;
    for (int64_t i = 0; i + 1 < n; ++i) for (int64_t j = i + 1; j < n; ++j) {
        // make_motion_pair (handeyedlt.cpp:11-23): A = bTg_i^-1 * bTg_j ; B = cTt_i * cTt_j^-1
        const double *Ri = bg + 12 * i, *ti = Ri + 9, *Rj = bg + 12 * j, *tj = Rj + 9;
        double Rii[9], tii[3]; invert_transform(Ri, ti, Rii, tii);
        double RA[9], tA[3]; se3_product(Rii, tii, Rj, tj, RA, tA);
        const double *Ci = ct + 12 * i, *ci = Ci + 9, *Cj = ct + 12 * j, *cj = Cj + 9;
        double Cji[9], cji[3]; invert_transform(Cj, cj, Cji, cji);
        double RB[9], tB[3]; se3_product(Ci, ci, Cji, cji, RB, tB);
        double PA[9], PB[9]; project_to_so3(RA, PA); project_to_so3(RB, PB);
        // is_good_pair (handeyedlt.cpp:25-49)
        double al[3], be[3]; log_so3(PA, al); log_so3(PB, be);
        const double na = std::sqrt(al[0] * al[0] + al[1] * al[1] + al[2] * al[2]);
        const double nb = std::sqrt(be[0] * be[0] + be[1] * be[1] + be[2] * be[2]);
        if (std::min(na, nb) < min_angle) continue;
        if (na >= 1e-9 && nb >= 1e-9) {
            const double a[3] = {al[0] / na, al[1] / na, al[2] / na}, b[3] = {be[0] / nb, be[1] / nb, be[2] / nb};
            const double cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2], cz = a[0] * b[1] - a[1] * b[0];
            if (std::sqrt(cx * cx + cy * cy + cz * cz) < 1e-3) continue;
        }
        if (rot_a) { std::memcpy(rot_a + 9 * cnt, PA, sizeof PA); std::memcpy(rot_b + 9 * cnt, PB, sizeof PB); std::memcpy(tra_a + 3 * cnt, tA, sizeof tA); std::memcpy(tra_b + 3 * cnt, tB, sizeof tB); }
        ++cnt;
    }
    return cnt;
}

}  // extern "C"
