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
#include <algorithm>

#include "axxb_kernels.cuh"
#include "comm.h"
#include "k1_math.cuh"

using namespace calk;

namespace {

thread_local std::string g_axxb_err;

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
    // pairs formed on the fly from the poses (cal_axxb_create_from_poses)
    bool from_poses = false;
    PairTiles tiles{};
    // sharded over the ranks of a communicator (cal_axxb_attach_comm): this rank evaluates the tiles [tile0, tile0 + n_cta)
    // and the 28 sums (upper 6x6, gradient, cost) are all-reduced after every pass
    int64_t tile0 = 0;
    calcomm::Comm* comm = nullptr;   // not owned
    double *dG = nullptr, *dC = nullptr;
    ~cal_axxb_handle() {
        cudaFree(pairs); cudaFree(x); cudaFree(partial); cudaFree(out); cudaFree(dG); cudaFree(dC); cudaFree(tiles.mask);
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
    if (h.from_poses) {
        PairTiles t = h.tiles; t.tile_base = h.tile0;
        if (jac) k_axxb_otf<1><<<h.n_cta, 256, 0, h.st>>>(t, h.x, h.huber, h.partial);
        else k_axxb_otf<0><<<h.n_cta, 256, 0, h.st>>>(t, h.x, h.huber, h.partial);
    } else if (jac) k_axxb<1><<<h.n_cta, 256, 0, h.st>>>(h.pairs, h.n, h.x, h.huber, h.partial);
    else k_axxb<0><<<h.n_cta, 256, 0, h.st>>>(h.pairs, h.n, h.x, h.huber, h.partial);
    k_axxb_final<<<1, 32 * kAcc, 0, h.st>>>(h.partial, h.n_cta, h.out);
    h.launches += 2;
    if (h.comm && !h.comm->allreduce_sum(h.out, kAcc, h.st)) return afail(CAL_ERR_COMM, h.comm->error());
    double o[kAcc];
    ACUDA(cudaMemcpyAsync(o, h.out, sizeof o, cudaMemcpyDeviceToHost, h.st));
    ACUDA(cudaStreamSynchronize(h.st));
    ACUDA(cudaGetLastError());
    if (h.comm && !h.comm->check_timeout()) return afail(CAL_ERR_COMM, h.comm->error());
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

extern "C" cal_status cal_axxb_create_from_poses(int64_t n_poses, const double* base_se3_gripper, const double* cam_se3_target,
                                                 double min_angle_deg, int reject_axis_parallel, double axis_parallel_eps,
                                                 double huber_delta, int device, cal_axxb_handle** out, int64_t* n_pairs_kept) {
    if (!out) return afail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    // build_all_pairs: std::runtime_error on inconsistent sizes (handeyedlt.cpp:56-58)
    if (n_poses < 2 || !base_se3_gripper || !cam_se3_target) return afail(CAL_ERR_RUNTIME, "Inconsistent hand-eye input sizes");
    if (cal_device_count() <= device) return afail(CAL_ERR_CUDA, "no CUDA device: calib_b200 has no CPU fallback");
    ACUDA(cudaSetDevice(device));
    const int T = (int)((n_poses + kTile - 1) / kTile);
    const int64_t n_tiles = (int64_t)T * (T + 1) / 2;
    if (n_tiles > 0x7fffffff) return afail(CAL_ERR_INVALID_ARGUMENT, "too many poses for the pair tiling");
    cal_axxb_handle* h = new cal_axxb_handle;
    h->device = device; h->huber = huber_delta; h->from_poses = true; h->n_cta = (int)n_tiles;
    unsigned long long* dcount = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->dG), sizeof(double) * 12 * n_poses);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->dC), sizeof(double) * 12 * n_poses);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->tiles.mask), sizeof(unsigned) * 32 * n_tiles);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->x), sizeof(double) * 8);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->partial), sizeof(double) * kAcc * n_tiles);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->out), sizeof(double) * kAcc);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&dcount), sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->dG, base_se3_gripper, sizeof(double) * 12 * n_poses, cudaMemcpyDefault, h->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->dC, cam_se3_target, sizeof(double) * 12 * n_poses, cudaMemcpyDefault, h->st);
    if (e == cudaSuccess) e = cudaMemsetAsync(dcount, 0, sizeof(unsigned long long), h->st);
    unsigned long long kept = 0;
    if (e == cudaSuccess) {
        h->tiles.n_poses = n_poses; h->tiles.n_tiles_1d = T; h->tiles.G = h->dG; h->tiles.C = h->dC;
        const double kPi = 3.14159265358979323846;
        k_pair_mask<<<(unsigned)n_tiles, 256, 0, h->st>>>(h->tiles, min_angle_deg * kPi / 180.0, reject_axis_parallel, axis_parallel_eps, dcount);
        e = cudaMemcpyAsync(&kept, dcount, sizeof kept, cudaMemcpyDeviceToHost, h->st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(dcount);
    if (e != cudaSuccess) { delete h; return afail(CAL_ERR_CUDA, cudaGetErrorString(e)); }
    h->n = (int64_t)kept;
    if (n_pairs_kept) *n_pairs_kept = h->n;
    if (kept == 0) { delete h; return afail(CAL_ERR_RUNTIME, "No valid motion pairs after filtering. Increase motion or relax thresholds."); }
    *out = h;
    return CAL_OK;
}

extern "C" void cal_axxb_destroy(cal_axxb_handle* h) { if (h) { cudaSetDevice(h->device); delete h; } }

extern "C" cal_status cal_axxb_eval(cal_axxb_handle* h, const double* x7, double* cost, double* g6, double* H36) {
    if (!h || !x7) return afail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    ACUDA(cudaSetDevice(h->device));
    return axxb_pass(*h, x7, g6 || H36, cost, g6, H36);
}

// Shard a from-poses handle over the ranks of a communicator: every rank holds all poses (2 x 96 B each) and
// the pair mask, evaluates a contiguous slice of the 32 x 32 pair tiles, and the 28 sums are all-reduced after every
// pass — the LM then runs replicated.  comm == NULL detaches (the handle evaluates every tile again).
extern "C" cal_status cal_axxb_attach_comm(cal_axxb_handle* h, cal_comm* c) {
    if (!h) return afail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    if (c && c->c && !h->from_poses) return afail(CAL_ERR_INVALID_ARGUMENT, "only handles created from poses can be sharded");
    const int T = h->tiles.n_tiles_1d;
    const int64_t n_tiles = (int64_t)T * (T + 1) / 2;
    if (!c || !c->c) { h->comm = nullptr; h->tile0 = 0; if (h->from_poses) h->n_cta = (int)n_tiles; return CAL_OK; }
    const int rank = c->c->rank(), world = c->c->world();
    if (n_tiles < world) return afail(CAL_ERR_INVALID_ARGUMENT, "fewer pair tiles than ranks");
    const int64_t per = n_tiles / world, rem = n_tiles % world;
    h->tile0 = rank * per + std::min<int64_t>(rank, rem);
    h->n_cta = (int)(per + (rank < rem ? 1 : 0));
    h->comm = c->c;
    return CAL_OK;
}

// benchmark hook: `reps` Jacobian passes back to back on the handle's stream, timed with CUDA events on that stream
extern "C" cal_status cal_axxb_bench_pass(cal_axxb_handle* h, const double* x7, int reps, float* ms_total) {
    if (!h || !x7 || reps <= 0 || !ms_total) return afail(CAL_ERR_INVALID_ARGUMENT, "bad argument");
    ACUDA(cudaSetDevice(h->device));
    ACUDA(cudaMemcpyAsync(h->x, x7, 7 * sizeof(double), cudaMemcpyHostToDevice, h->st));
    cudaEvent_t e0, e1; ACUDA(cudaEventCreate(&e0)); ACUDA(cudaEventCreate(&e1));
    ACUDA(cudaStreamSynchronize(h->st));
    ACUDA(cudaEventRecord(e0, h->st));
    for (int r = 0; r < reps; ++r) {
        PairTiles t = h->tiles; t.tile_base = h->tile0;
        if (h->from_poses) k_axxb_otf<1><<<h->n_cta, 256, 0, h->st>>>(t, h->x, h->huber, h->partial);
        else k_axxb<1><<<h->n_cta, 256, 0, h->st>>>(h->pairs, h->n, h->x, h->huber, h->partial);
        k_axxb_final<<<1, 32 * kAcc, 0, h->st>>>(h->partial, h->n_cta, h->out);
        h->launches += 2;
        if (h->comm && !h->comm->allreduce_sum(h->out, kAcc, h->st)) return afail(CAL_ERR_COMM, h->comm->error());
    }
    ACUDA(cudaEventRecord(e1, h->st));
    ACUDA(cudaEventSynchronize(e1));
    ACUDA(cudaGetLastError());
    ACUDA(cudaEventElapsedTime(ms_total, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return CAL_OK;
}
extern "C" int64_t cal_axxb_launch_count(const cal_axxb_handle* h) { return h ? h->launches : 0; }

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
