// AX = XB hand-eye refinement: replaces optimize_handeye's Ceres problem
// (reference src/estimation/optim/handeye.cpp:45-78 over the AxXbResidual of
// src/estimation/residuals/handeyeresidual.h:18-54).
//
// One residual block per motion pair (6 residuals, 6 tangent dofs, per-pair
// Huber loss).  The kernel reads each pair once (24 doubles = 192 B, SoA so the
// loads are coalesced), forms the residual and its ANALYTIC Jacobian
//     r_rot = Log(R_A R_X R_B^T R_X^T)      J = [ 2 Jl^-1(r_rot) (R_A - R_S)    0      ]
//     r_t   = (R_A - I) t_X - (R_X t_B - t_A)    [ 2 [R_X t_B]x               R_A - I ]
// in registers (tangent = QuaternionManifold increments, rotation by 2|delta|)
// and accumulates the 6x6 normal equations with fixed-order reductions.
// HBM-bound: ~350 flop per 192 B.  LM driver on the host.
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/calib_b200.h"
#include "k1_math.cuh"

using namespace calk;

namespace {

thread_local std::string g_axxb_err;

constexpr int kAcc = 28;  // 21 (H upper) + 6 (g) + 1 (cost)

// Eigen::AngleAxis(Matrix3) = Quaternion(Matrix3) -> angle * axis
__device__ __forceinline__ void log_so3_eigen(const double* R, double* phi) {
    double q[4];
    double t = R[0] + R[4] + R[8];
    if (t > 0.0) {
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
    double n = sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (n != 0.0) {
        const double angle = 2.0 * atan2(n, fabs(q[0]));
        if (q[0] < 0.0) n = -n;
        const double s = angle / n;
        phi[0] = q[1] * s; phi[1] = q[2] * s; phi[2] = q[3] * s;
    } else { phi[0] = phi[1] = phi[2] = 0.0; }
}

// inverse left Jacobian of SO(3): d Log(Exp(eps) R) / d eps at eps = 0, phi = Log(R)
__device__ __forceinline__ void so3_left_jacobian_inv(const double* phi, double* J) {
    const double th2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
    double c;
    if (th2 < 1e-8) c = 1.0 / 12.0 + th2 / 720.0 + th2 * th2 / 30240.0;
    else { const double th = sqrt(th2); c = 1.0 / th2 - (1.0 + cos(th)) / (2.0 * th * sin(th)); }
    const double x = phi[0], y = phi[1], z = phi[2];
    // I - 1/2 [phi]x + c [phi]x^2 ; [phi]x^2 = phi phi^T - th2 I
    J[0] = 1.0 + c * (x * x - th2); J[1] = 0.5 * z + c * x * y;      J[2] = -0.5 * y + c * x * z;
    J[3] = -0.5 * z + c * x * y;    J[4] = 1.0 + c * (y * y - th2); J[5] = 0.5 * x + c * y * z;
    J[6] = 0.5 * y + c * x * z;     J[7] = -0.5 * x + c * y * z;    J[8] = 1.0 + c * (z * z - th2);
}

template <int JAC>
__global__ void __launch_bounds__(256) k_axxb(const double* __restrict__ pairs /*[24][n]*/, int64_t n, const double* __restrict__ x7,
                                              double huber, double* __restrict__ partial) {
    __shared__ double sm[8][kAcc];
    double q[4], tx[3];
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = x7[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) tx[i] = x7[4 + i];
    double Rx[9]; quat_to_R(q, Rx);
    double acc[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) acc[i] = 0.0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        double Ra[9], Rb[9], ta[3], tb[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) { Ra[i] = pairs[(int64_t)i * n + p]; Rb[i] = pairs[(int64_t)(9 + i) * n + p]; }
#pragma unroll
        for (int i = 0; i < 3; ++i) { ta[i] = pairs[(int64_t)(18 + i) * n + p]; tb[i] = pairs[(int64_t)(21 + i) * n + p]; }
        // rot_s = rot_a * rot_x * rot_b^T * rot_x^T (handeyeresidual.h:32)
        double M1[9], M2[9], Rs[9];
        mat3_mul(Ra, Rx, M1);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) M2[3 * i + j] = M1[3 * i] * Rb[3 * j] + M1[3 * i + 1] * Rb[3 * j + 1] + M1[3 * i + 2] * Rb[3 * j + 2];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) Rs[3 * i + j] = M2[3 * i] * Rx[3 * j] + M2[3 * i + 1] * Rx[3 * j + 1] + M2[3 * i + 2] * Rx[3 * j + 2];
        double r[6];
        log_so3_eigen(Rs, r);
        double rtb[3]; mat3_vec(Rx, tb, rtb);
#pragma unroll
        for (int i = 0; i < 3; ++i)
            r[3 + i] = (Ra[3 * i] - (i == 0 ? 1.0 : 0.0)) * tx[0] + (Ra[3 * i + 1] - (i == 1 ? 1.0 : 0.0)) * tx[1] +
                       (Ra[3 * i + 2] - (i == 2 ? 1.0 : 0.0)) * tx[2] - (rtb[i] - ta[i]);
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) s = fma(r[i], r[i], s);
        double rho, w; huber_weight(huber, s, rho, w);
        acc[27] += 0.5 * rho;
        if (JAC) {
            double J[36];
            double Jl[9]; so3_left_jacobian_inv(r, Jl);
            double D[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) D[i] = Ra[i] - Rs[i];
            double JD[9]; mat3_mul(Jl, D, JD);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    J[6 * i + j] = 2.0 * JD[3 * i + j];
                    J[6 * i + 3 + j] = 0.0;
                    J[6 * (3 + i) + 3 + j] = Ra[3 * i + j] - (i == j ? 1.0 : 0.0);
                }
            // d(-R_X t_B)/d delta = 2 [R_X t_B]x
            J[18] = 0.0;             J[19] = -2.0 * rtb[2];  J[20] = 2.0 * rtb[1];
            J[24] = 2.0 * rtb[2];    J[25] = 0.0;            J[26] = -2.0 * rtb[0];
            J[30] = -2.0 * rtb[1];   J[31] = 2.0 * rtb[0];   J[32] = 0.0;
            int o = 0;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = a; b < 6; ++b) {
                    double h = 0.0;
#pragma unroll
                    for (int i = 0; i < 6; ++i) h = fma(J[6 * i + a], J[6 * i + b], h);
                    acc[o] = fma(w, h, acc[o]); ++o;
                }
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double g = 0.0;
#pragma unroll
                for (int i = 0; i < 6; ++i) g = fma(J[6 * i + a], r[i], g);
                acc[21 + a] = fma(w, g, acc[21 + a]);
            }
        }
    }
    // fixed-order reduction: shuffle tree, then shared memory, one partial row per CTA
#pragma unroll
    for (int i = 0; i < kAcc; ++i) {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < kAcc) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
        partial[(int64_t)blockIdx.x * kAcc + threadIdx.x] = t;
    }
}

__global__ void k_axxb_final(const double* __restrict__ partial, int n_cta, double* __restrict__ out) {
    if (threadIdx.x >= kAcc) return;
    double t = 0.0;
    for (int c = 0; c < n_cta; ++c) t += partial[(int64_t)c * kAcc + threadIdx.x];
    out[threadIdx.x] = t;
}

// AoS [n][w] -> SoA rows [row0 + j][n]
__global__ void k_axxb_transpose(const double* __restrict__ src, int w, int64_t n, double* __restrict__ dst, int row0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * w) return;
    const int64_t p = i / w; const int j = (int)(i % w);
    dst[(int64_t)(row0 + j) * n + p] = src[i];
}

bool chol6h(double* A) { return chol6(A); }

}  // namespace

struct cal_axxb_handle {
    int device = 0;
    cudaStream_t st = nullptr;
    int64_t n = 0;
    double huber = 0;
    double *pairs = nullptr, *x = nullptr, *partial = nullptr, *out = nullptr;
    int n_cta = 1;
    int64_t launches = 0;
    ~cal_axxb_handle() {
        cudaFree(pairs); cudaFree(x); cudaFree(partial); cudaFree(out);
        if (st) cudaStreamDestroy(st);
    }
};

namespace {
cal_status afail(cal_status s, const std::string& m);
#define ACUDA(expr)                                                                                   \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) return afail(CAL_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

cal_status axxb_pass(cal_axxb_handle& h, const double* x7, bool jac, double* cost, double* g6, double* H36) {
    ACUDA(cudaMemcpyAsync(h.x, x7, 7 * sizeof(double), cudaMemcpyHostToDevice, h.st));
    if (jac) k_axxb<1><<<h.n_cta, 256, 0, h.st>>>(h.pairs, h.n, h.x, h.huber, h.partial);
    else k_axxb<0><<<h.n_cta, 256, 0, h.st>>>(h.pairs, h.n, h.x, h.huber, h.partial);
    k_axxb_final<<<1, 32, 0, h.st>>>(h.partial, h.n_cta, h.out);
    h.launches += 2;
    double o[kAcc];
    ACUDA(cudaMemcpyAsync(o, h.out, sizeof o, cudaMemcpyDeviceToHost, h.st));
    ACUDA(cudaStreamSynchronize(h.st));
    ACUDA(cudaGetLastError());
    if (cost) *cost = o[27];
    if (jac) {
        int k = 0;
        for (int a = 0; a < 6; ++a) for (int b = a; b < 6; ++b) { if (H36) { H36[6 * a + b] = o[k]; H36[6 * b + a] = o[k]; } ++k; }
        if (g6) for (int a = 0; a < 6; ++a) g6[a] = o[21 + a];
    }
    return CAL_OK;
}
}  // namespace

// error string shared with refine_host.cu's cal_last_error through a setter
extern "C" void cal_set_last_error_(const char* msg);
namespace { cal_status afail(cal_status s, const std::string& m) { cal_set_last_error_(m.c_str()); return s; } }

extern "C" cal_status cal_axxb_create(const cal_axxb_desc* d, int device, cal_axxb_handle** out) {
    if (!d || !out) return afail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    // build_all_pairs throws std::runtime_error when no pair survives (handeyedlt.cpp:76-79)
    if (d->n_pairs <= 0) return afail(CAL_ERR_RUNTIME, "No valid motion pairs after filtering. Increase motion or relax thresholds.");
    if (!d->rot_a || !d->rot_b || !d->tra_a || !d->tra_b) return afail(CAL_ERR_INVALID_ARGUMENT, "null pair arrays");
    if (cal_device_count() <= device) return afail(CAL_ERR_CUDA, "no CUDA device: calib_b200 has no CPU fallback");
    ACUDA(cudaSetDevice(device));
    cal_axxb_handle* h = new cal_axxb_handle;
    h->device = device; h->n = d->n_pairs; h->huber = d->huber_delta;
    const int64_t n = h->n;
    cudaError_t e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->pairs), sizeof(double) * 24 * n);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->x), sizeof(double) * 8);
    h->n_cta = (int)std::min<int64_t>(148 * 8, (n + 255) / 256);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->partial), sizeof(double) * kAcc * h->n_cta);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->out), sizeof(double) * kAcc);
    double* stage = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&stage), sizeof(double) * 9 * n);
    if (e != cudaSuccess) { delete h; cudaFree(stage); return afail(CAL_ERR_CUDA, cudaGetErrorString(e)); }
    const double* srcs[4] = {d->rot_a, d->rot_b, d->tra_a, d->tra_b};
    const int widths[4] = {9, 9, 3, 3}; const int rows[4] = {0, 9, 18, 21};
    for (int k = 0; k < 4 && e == cudaSuccess; ++k) {
        e = cudaMemcpyAsync(stage, srcs[k], sizeof(double) * widths[k] * n, cudaMemcpyHostToDevice, h->st);
        const int64_t tot = n * widths[k];
        k_axxb_transpose<<<(unsigned)((tot + 255) / 256), 256, 0, h->st>>>(stage, widths[k], n, h->pairs, rows[k]);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    }
    cudaFree(stage);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { delete h; return afail(CAL_ERR_CUDA, cudaGetErrorString(e)); }
    *out = h;
    return CAL_OK;
}

extern "C" void cal_axxb_destroy(cal_axxb_handle* h) { if (h) { cudaSetDevice(h->device); delete h; } }

extern "C" cal_status cal_axxb_eval(cal_axxb_handle* h, const double* x7, double* cost, double* g6, double* H36) {
    if (!h || !x7) return afail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    ACUDA(cudaSetDevice(h->device));
    return axxb_pass(*h, x7, g6 || H36, cost, g6, H36);
}

// LM with Ceres 2.2 trust-region semantics on the dense 6x6 system.
extern "C" cal_status cal_axxb_solve(cal_axxb_handle* hp, const cal_optim_options* o, double* x7, cal_optim_result* res,
                                     double* cov49) {
    if (!hp || !o || !x7 || !res) return afail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    cal_axxb_handle& h = *hp;
    ACUDA(cudaSetDevice(h.device));
    std::memset(res, 0, sizeof *res);
    const double eps = o->epsilon;
    double x[7], xp[7], H[36], g[6], cost = 0;
    std::memcpy(x, x7, sizeof x);
    if (cal_status s = axxb_pass(h, x, true, &cost, g, H)) return s;
    int jev = 1, cev = 0, iter = 0, n_invalid = 0, term = CAL_TERM_NO_CONVERGENCE;
    res->initial_cost = cost;
    double sc[6], diag[6], y[6], step[6], delta[6];
    for (int i = 0; i < 6; ++i) sc[i] = 1.0 / (1.0 + std::sqrt(H[7 * i]));
    auto plus = [](const double* a, const double* d, double* out) { quat_plus(a, d, out); for (int i = 0; i < 3; ++i) out[4 + i] = a[4 + i] + d[3 + i]; };
    auto gnorm = [&]() { double ng[6], t[7], m = 0; for (int i = 0; i < 6; ++i) ng[i] = -g[i]; plus(x, ng, t); for (int i = 0; i < 7; ++i) m = std::max(m, std::fabs(x[i] - t[i])); return m; };
    auto xnorm = [&]() { double s = 0; for (double v : x) s += v * v; return std::sqrt(s); };
    double gmax = gnorm(), x_norm = xnorm(), radius = 1e4, dec = 2.0;
    bool reuse = false;
    for (;;) {
        if (iter >= o->max_iterations) { term = CAL_TERM_NO_CONVERGENCE; break; }
        if (gmax <= eps) { term = CAL_TERM_CONVERGENCE; break; }
        if (radius <= 1e-32) { term = CAL_TERM_CONVERGENCE; break; }
        ++iter;
        if (!reuse) for (int i = 0; i < 6; ++i) diag[i] = std::min(std::max(H[7 * i] * sc[i] * sc[i], 1e-6), 1e32);
        double A[36], Hs[36];
        for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) { Hs[6 * i + j] = H[6 * i + j] * sc[i] * sc[j]; A[6 * i + j] = Hs[6 * i + j]; } A[7 * i] += diag[i] / radius; y[i] = g[i] * sc[i]; }
        bool ok = chol6h(A); if (ok) chol6_solve(A, y);
        reuse = true;
        double mcc = 0;
        if (ok) {
            double sg = 0, quad = 0;
            for (int i = 0; i < 6; ++i) { if (!std::isfinite(y[i])) ok = false; step[i] = -y[i]; sg += step[i] * g[i] * sc[i]; }
            for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) quad += Hs[6 * i + j] * step[i] * step[j];
            mcc = -(sg + 0.5 * quad); ok = ok && mcc > 0.0;
        }
        if (!ok) { if (++n_invalid >= 5) { term = CAL_TERM_FAILURE; break; } radius /= dec; dec *= 2.0; continue; }
        n_invalid = 0;
        for (int i = 0; i < 6; ++i) delta[i] = step[i] * sc[i];
        plus(x, delta, xp);
        double cc = 0;
        if (cal_status s = axxb_pass(h, xp, false, &cc, nullptr, nullptr)) return s;
        ++cev;
        if (!std::isfinite(cc)) cc = std::numeric_limits<double>::max();
        double sn = 0; for (int i = 0; i < 7; ++i) sn += (x[i] - xp[i]) * (x[i] - xp[i]);
        if (std::sqrt(sn) <= eps * (x_norm + eps)) { term = CAL_TERM_CONVERGENCE; break; }
        const double change = cost - cc;
        if (std::fabs(change) <= eps * cost) { term = CAL_TERM_CONVERGENCE; break; }
        const double rho = change / mcc;
        if (rho > 1e-3) {
            std::memcpy(x, xp, sizeof x);
            if (cal_status s = axxb_pass(h, x, true, &cost, g, H)) return s;
            ++jev; gmax = gnorm(); x_norm = xnorm();
            radius = std::min(1e16, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
            dec = 2.0; reuse = false;
        } else { radius /= dec; dec *= 2.0; }
    }
    std::memcpy(x7, x, sizeof x);
    res->success = term == CAL_TERM_CONVERGENCE; res->iterations = iter; res->termination = term;
    res->num_jac_evals = jev; res->num_cost_evals = cev; res->final_cost = cost;
    static const char* names[] = {"CONVERGENCE", "NO_CONVERGENCE", "FAILURE"};
    std::snprintf(res->report, sizeof res->report, "Ceres Solver Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s",
                  iter, res->initial_cost, cost, names[term]);
    if (cov49 && o->compute_covariance) {
        double L[36]; std::memcpy(L, H, sizeof L);
        if (chol6h(L)) {
            double C[36];
            for (int j = 0; j < 6; ++j) { double e[6] = {0, 0, 0, 0, 0, 0}; e[j] = 1.0; chol6_solve(L, e); for (int i = 0; i < 6; ++i) C[6 * i + j] = e[i]; }
            double P[42] = {0};
            const double* q = x;
            const double PJ[12] = {-q[1], -q[2], -q[3], q[0], q[3], -q[2], -q[3], q[0], q[1], q[2], -q[1], q[0]};
            for (int j = 0; j < 4; ++j) for (int k = 0; k < 3; ++k) P[6 * j + k] = PJ[3 * j + k];
            for (int j = 0; j < 3; ++j) P[6 * (4 + j) + 3 + j] = 1.0;
            for (int i = 0; i < 7; ++i) for (int j = 0; j < 7; ++j) { double a = 0; for (int k = 0; k < 6; ++k) for (int l = 0; l < 6; ++l) a += P[6 * i + k] * C[6 * k + l] * P[6 * j + l]; cov49[7 * i + j] = a; }
            res->covariance_ok = 1;
        }
    }
    return CAL_OK;
}
