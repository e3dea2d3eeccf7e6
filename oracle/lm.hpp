// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).
//
// Restatement of the Ceres 2.2 trust-region Levenberg–Marquardt loop that
// `solve_problem` (reference src/estimation/detail/ceresutils.h:27-43) runs.
// Ceres itself is a third-party dependency that is NOT vendored in the
// reference and NOT installable here (cmake/Dependencies.cmake:1, unpinned;
// de-facto Ceres 2.2.0).  What follows restates its published algorithm
// (trust_region_minimizer.cc, levenberg_marquardt_strategy.cc) with the
// options the reference sets: function = gradient = parameter tolerance =
// epsilon, max_num_iterations, everything else default.
#pragma once
#include <cmath>
#include <cstdio>
#include <limits>
#include <string>
#include <vector>

namespace orc {

struct LMProblem {
    virtual ~LMProblem() = default;
    virtual int n_amb() const = 0;  // ambient parameter count
    virtual int n_int() const = 0;  // tangent dims in the problem's internal order
    // Evaluate at x. Always sets cost; if jac, also refreshes the normal
    // equations held by the problem (H = J^T J, g = J^T r, loss applied).
    virtual bool eval(const double* x, double* cost, bool jac) = 0;
    virtual void diag(double* out) const = 0;   // H_ii (squared column norms of J)
    virtual void grad(double* out) const = 0;   // g, unscaled
    // Solve (S H S + diag(D2)) y = S g with S = diag(s). false on failure.
    virtual bool solve(const double* s, const double* D2, double* y) = 0;
    // step^T (S H S) step
    virtual double quad(const double* s, const double* step) const = 0;
    // x_plus = x [+] delta (manifold plus per block, then projection onto bounds)
    virtual void plus(const double* x, const double* delta, double* xp) const = 0;
    virtual bool constrained() const = 0;
};

struct LMOptions {
    int max_iterations = 1000;
    double epsilon = 1e-9;
    bool verbose = false;
};
struct LMSummary {
    int termination = 1;  // 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE
    int iterations = 0;
    int num_jac_evals = 0, num_cost_evals = 0;
    double initial_cost = 0, final_cost = 0;
    std::string message;
};

inline double vec_dot(const std::vector<double>& a, const std::vector<double>& b) {
    double s = 0; for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i]; return s; }

inline void lm_gradient_norms(const LMProblem& p, const std::vector<double>& x, double* gmax) {
    // trust_region_minimizer.cc EvaluateGradientAndJacobian: the norm is of
    // x - Plus(x, -g) in ambient coordinates.
    const int n = p.n_int(), na = p.n_amb();
    std::vector<double> g(n), ng(n), xp(na);
    p.grad(g.data());
    for (int i = 0; i < n; ++i) ng[i] = -g[i];
    p.plus(x.data(), ng.data(), xp.data());
    double m = 0; for (int i = 0; i < na; ++i) m = std::max(m, std::fabs(x[i] - xp[i]));
    *gmax = m;
}

inline LMSummary lm_minimize(LMProblem& p, const LMOptions& o, std::vector<double>& x) {
    LMSummary sum;
    const int n = p.n_int(), na = p.n_amb();
    const double min_relative_decrease = 1e-3, min_diag = 1e-6, max_diag = 1e32;
    const double max_radius = 1e16, min_radius = 1e-32;
    double radius = 1e4, decrease_factor = 2.0;
    bool reuse_diag = false;
    std::vector<double> s(n, 1.0), diag(n), D2(n), y(n), step(n), delta(n), g(n), gs(n), xp(na), hd(n);

    if (p.constrained()) {  // IterationZero: project the start onto the feasible set
        std::vector<double> z(n, 0.0);
        p.plus(x.data(), z.data(), xp.data());
        x = xp;
    }
    double cost = 0;
    if (!p.eval(x.data(), &cost, true)) { sum.termination = 2; sum.message = "initial evaluation failed"; return sum; }
    sum.num_jac_evals++;
    sum.initial_cost = cost;
    p.diag(hd.data());
    for (int i = 0; i < n; ++i) s[i] = 1.0 / (1.0 + std::sqrt(hd[i]));  // jacobi_scaling, once
    double gmax = 0; lm_gradient_norms(p, x, &gmax);
    double x_norm = 0; for (double v : x) x_norm += v * v; x_norm = std::sqrt(x_norm);
    int iter = 0, n_invalid = 0;
    if (o.verbose) std::printf("iter      cost      cost_change  |gradient|   tr_radius\n%4d % .6e %.2e %.2e %.2e\n", 0, cost, 0.0, gmax, radius);

    for (;;) {
        // FinalizeIterationAndCheckIfMinimizerCanContinue
        if (iter >= o.max_iterations) { sum.termination = 1; sum.message = "Maximum number of iterations reached."; break; }
        if (gmax <= o.epsilon) { sum.termination = 0; sum.message = "Gradient tolerance reached."; break; }
        if (radius <= min_radius) { sum.termination = 0; sum.message = "Minimum trust region radius reached."; break; }
        ++iter;
        // LevenbergMarquardtStrategy::ComputeStep
        if (!reuse_diag) {
            p.diag(hd.data());
            for (int i = 0; i < n; ++i) diag[i] = std::min(std::max(hd[i] * s[i] * s[i], min_diag), max_diag);
        }
        for (int i = 0; i < n; ++i) D2[i] = diag[i] / radius;
        bool ok = p.solve(s.data(), D2.data(), y.data());
        reuse_diag = true;
        if (ok) for (int i = 0; i < n; ++i) if (!std::isfinite(y[i])) { ok = false; break; }
        double model_cost_change = 0;
        if (ok) {
            p.grad(g.data());
            double sg = 0;
            for (int i = 0; i < n; ++i) { step[i] = -y[i]; gs[i] = g[i] * s[i]; sg += step[i] * gs[i]; }
            // -(J step)'(r + J step / 2) = -(step'g + step'H step / 2)
            model_cost_change = -(sg + 0.5 * p.quad(s.data(), step.data()));
            ok = model_cost_change > 0.0;
        }
        if (!ok) {  // HandleInvalidStep
            if (++n_invalid >= 5) { sum.termination = 2; sum.message = "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps: 5"; break; }
            radius /= decrease_factor; decrease_factor *= 2.0; reuse_diag = true;
            continue;
        }
        n_invalid = 0;
        for (int i = 0; i < n; ++i) delta[i] = step[i] * s[i];
        double cand_cost = 0; bool have_cand = false;
        if (p.constrained()) {
            // DoLineSearch: Armijo projected line search along delta starting at
            // step size 1 (sufficient decrease 1e-4).  Accepted at 1 in the
            // normal case, leaving delta unchanged.  Backtracking restated with
            // value-only quadratic interpolation clamped to [1e-3, 0.6] x the
            // current step (Ceres interpolates cubically with gradients: the
            // accepted-at-1 path is identical, the backtracking path is not
            // pinned — see DESIGN.md).
            double g0 = 0; for (int i = 0; i < n; ++i) g0 += g[i] * delta[i];
            double t = 1.0; std::vector<double> dt(n);
            for (int ls = 0; ls < 20; ++ls) {
                for (int i = 0; i < n; ++i) dt[i] = t * delta[i];
                p.plus(x.data(), dt.data(), xp.data());
                double c = 0; bool v = p.eval(xp.data(), &c, false) && std::isfinite(c);
                sum.num_cost_evals++;
                if (v && c <= cost + 1e-4 * g0 * t) { cand_cost = c; have_cand = true; break; }
                double tn = 0.5 * t;
                if (v) { const double denom = 2.0 * (c - cost - g0 * t); if (denom > 0) tn = -g0 * t * t / denom; }
                tn = std::min(std::max(tn, 1e-3 * t), 0.6 * t);
                t = tn;
            }
            if (have_cand && t != 1.0) for (int i = 0; i < n; ++i) delta[i] *= t;
            have_cand = have_cand && true;
        }
        // ComputeCandidatePointAndEvaluateCost
        p.plus(x.data(), delta.data(), xp.data());
        if (!have_cand) {
            bool v = p.eval(xp.data(), &cand_cost, false);
            sum.num_cost_evals++;
            if (!v || !std::isfinite(cand_cost)) cand_cost = std::numeric_limits<double>::max();
        }
        // ParameterToleranceReached
        double sn = 0; for (int i = 0; i < na; ++i) { const double d = x[i] - xp[i]; sn += d * d; } sn = std::sqrt(sn);
        if (sn <= o.epsilon * (x_norm + o.epsilon)) { sum.termination = 0; sum.message = "Parameter tolerance reached."; break; }
        // FunctionToleranceReached
        const double cost_change = cost - cand_cost;
        if (std::fabs(cost_change) <= o.epsilon * cost) { sum.termination = 0; sum.message = "Function tolerance reached."; break; }
        const double rho = cost_change / model_cost_change;
        if (rho > min_relative_decrease) {  // HandleSuccessfulStep
            x = xp; x_norm = 0; for (double v : x) x_norm += v * v; x_norm = std::sqrt(x_norm);
            if (!p.eval(x.data(), &cost, true)) { sum.termination = 2; sum.message = "evaluation failed"; break; }
            sum.num_jac_evals++;
            lm_gradient_norms(p, x, &gmax);
            radius = std::min(max_radius, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
            decrease_factor = 2.0; reuse_diag = false;
        } else {  // HandleUnsuccessfulStep
            radius /= decrease_factor; decrease_factor *= 2.0; reuse_diag = true;
        }
        if (o.verbose) std::printf("%4d % .6e %.2e %.2e %.2e rho=%.2e\n", iter, cost, cost_change, gmax, radius, rho);
    }
    sum.iterations = iter;
    sum.final_cost = cost;
    return sum;
}

inline std::string brief_report(const LMSummary& s) {
    // ceres::Solver::Summary::BriefReport()
    static const char* names[] = {"CONVERGENCE", "NO_CONVERGENCE", "FAILURE"};
    char buf[256];
    std::snprintf(buf, sizeof buf, "Ceres Solver Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s",
                  s.iterations, s.initial_cost, s.final_cost, names[s.termination]);
    return buf;
}

}  // namespace orc
