// Device code of the AX = XB kernels (see axxb.cu for the description).  Kept in a header so that
// tests/host_emul can run this very source on the CPU under a lock-step SIMT shim (test-only).
#pragma once
#include <stdint.h>

#include "k1_math.cuh"

using namespace calk;

namespace {

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

// One motion pair: residual (handeyeresidual.h:25-49), per-pair Huber weight, analytic Jacobian,
// accumulation of the upper triangle of J^T J (21), J^T r (6) and the cost (1).
template <int JAC>
__device__ __forceinline__ void axxb_pair(const double* Ra, const double* Rb, const double* ta, const double* tb, const double* Rx,
                                          const double* tx, double huber, double* acc) {
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

// fixed-order reduction of the 28 accumulators of a 256-thread CTA: shuffle tree, then shared memory,
// one partial row per CTA
__device__ __forceinline__ void cta_reduce_store(const double* acc, double (*sm)[kAcc], double* __restrict__ partial) {
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

// CTAs per SM of the two pair kernels (see k_axxb_otf)
#ifndef CALK_AXXB_MINB
#define CALK_AXXB_MINB 2
#endif
template <int JAC>
__global__ void __launch_bounds__(256, CALK_AXXB_MINB) k_axxb(const double* __restrict__ pairs /*[24][n]*/, int64_t n, const double* __restrict__ x7,
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
        axxb_pair<JAC>(Ra, Rb, ta, tb, Rx, tx, huber, acc);
    }
    cta_reduce_store(acc, sm, partial);
}

// ---------------------------------------------------------------------------
// Motion pairs formed on the fly from the poses: build_all_pairs
// (src/estimation/linear/handeyedlt.cpp:11-81) is O(N^2) pairs of 192 B each — 2.4 GB at 5000
// poses — so nothing is materialised.  The index space i < j is cut into 32 x 32 tiles; a CTA
// stages the 2 x 32 poses of its tile in shared memory and every thread forms 4 pairs,
//     A = b_T_g,i^-1 b_T_g,j ,  B = c_T_t,i c_T_t,j^-1 ,  rotations projected onto SO(3)
// (make_motion_pair :11-23), in registers.  is_good_pair (:25-49) is evaluated once at create
// and kept as one bit per pair (32 words per tile).
// ---------------------------------------------------------------------------
constexpr int kTile = 32;

// project_to_so3 (se3_utils.h:10-19): orthogonal polar factor by Newton's iteration, stopped when
// stationary — one or two steps for the near-orthonormal products of rotation matrices.
__device__ __forceinline__ void project_so3(double* X) {
    for (int it = 0; it < 20; ++it) {
        const double c00 = X[4] * X[8] - X[5] * X[7], c01 = X[5] * X[6] - X[3] * X[8], c02 = X[3] * X[7] - X[4] * X[6];
        const double det = X[0] * c00 + X[1] * c01 + X[2] * c02;
        const double id = 1.0 / det;
        double Y[9];  // X^-T = cofactors / det
        Y[0] = c00 * id; Y[1] = c01 * id; Y[2] = c02 * id;
        Y[3] = (X[2] * X[7] - X[1] * X[8]) * id; Y[4] = (X[0] * X[8] - X[2] * X[6]) * id; Y[5] = (X[1] * X[6] - X[0] * X[7]) * id;
        Y[6] = (X[1] * X[5] - X[2] * X[4]) * id; Y[7] = (X[2] * X[3] - X[0] * X[5]) * id; Y[8] = (X[0] * X[4] - X[1] * X[3]) * id;
        double g = 1.0;
        if (it < 3 && fabs(fabs(det) - 1.0) > 0.1) {  // far from orthogonal: Frobenius-norm scaling speeds up the first steps
            double nx = 0, ny = 0;
#pragma unroll
            for (int i = 0; i < 9; ++i) { nx = fma(X[i], X[i], nx); ny = fma(Y[i], Y[i], ny); }
            g = sqrt(sqrt(ny / nx));
        }
        const double a = 0.5 * g, b = 0.5 / g;
        double d = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) { const double xn = a * X[i] + b * Y[i]; d = fmax(d, fabs(xn - X[i])); X[i] = xn; }
        if (!(d > 4e-16)) break;
    }
}

// log_so3 (se3_utils.h:27-42; its own project_to_so3 is the identity on an already projected matrix)
__device__ __forceinline__ void log_so3_ref(const double* R, double* w) {
    double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
    c = fmin(1.0, fmax(-1.0, c));
    const double th = acos(c);
    if (th < 1e-12) { w[0] = w[1] = w[2] = 0.0; return; }
    const double k = 0.5 / sin(th) * th;
    w[0] = (R[7] - R[5]) * k; w[1] = (R[2] - R[6]) * k; w[2] = (R[3] - R[1]) * k;
}

// pose layout in shared memory: [12] = R row-major, t
__device__ __forceinline__ void make_pair(const double* Gi, const double* Gj, const double* Ci, const double* Cj, double* Ra, double* Rb,
                                          double* ta, double* tb) {
    // A = Gi^-1 Gj : R = Ri^T Rj, t = Ri^T (tj - ti)
    const double d[3] = {Gj[9] - Gi[9], Gj[10] - Gi[10], Gj[11] - Gi[11]};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) Ra[3 * r + c] = Gi[r] * Gj[c] + Gi[3 + r] * Gj[3 + c] + Gi[6 + r] * Gj[6 + c];
        ta[r] = Gi[r] * d[0] + Gi[3 + r] * d[1] + Gi[6 + r] * d[2];
    }
    // B = Ci Cj^-1 : R = Ri Rj^T, t = ti - R tj  (with the unprojected R, as Eigen's Isometry product does)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) Rb[3 * r + c] = Ci[3 * r] * Cj[3 * c] + Ci[3 * r + 1] * Cj[3 * c + 1] + Ci[3 * r + 2] * Cj[3 * c + 2];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) tb[r] = Ci[9 + r] - (Rb[3 * r] * Cj[9] + Rb[3 * r + 1] * Cj[10] + Rb[3 * r + 2] * Cj[11]);
    project_so3(Ra); project_so3(Rb);
}

struct PairTiles {
    int64_t n_poses;
    int n_tiles_1d;
    const double* G;   // [n][12] base_se3_gripper
    const double* C;   // [n][12] cam_se3_target
    unsigned* mask;    // [tile][32] one bit per pair
    int64_t tile_base; // first tile of this launch (a rank of a sharded handle evaluates a slice of the tile space)
};

__device__ __forceinline__ void tile_coords(int64_t t, int T, int& bi, int& bj) {
    // tiles of the upper triangle, row-major: row bi holds T - bi tiles
    int row = 0; int64_t rem = t;
    while (rem >= T - row) { rem -= T - row; ++row; }
    bi = row; bj = row + (int)rem;
}

__device__ __forceinline__ void stage_tile(const PairTiles& P, int bi, int bj, double (*sG)[12], double (*sC)[12]) {
    // rows 0..31: poses of the tile's i range, rows 32..63: poses of its j range
    for (int k = threadIdx.x; k < 64 * 12; k += blockDim.x) {
        const int r = k / 12, e = k % 12;
        const int64_t pose = (int64_t)(r < 32 ? bi : bj) * kTile + (r & 31);
        const bool ok = pose < P.n_poses;
        sG[r][e] = ok ? P.G[pose * 12 + e] : (e < 9 && e % 4 == 0 ? 1.0 : 0.0);
        sC[r][e] = ok ? P.C[pose * 12 + e] : (e < 9 && e % 4 == 0 ? 1.0 : 0.0);
    }
    __syncthreads();
}

// is_good_pair (handeyedlt.cpp:25-49) for every pair, once
__global__ void __launch_bounds__(256) k_pair_mask(PairTiles P, double min_angle, int reject_parallel, double parallel_eps,
                                                  unsigned long long* __restrict__ count) {
    __shared__ double sG[64][12], sC[64][12];
    int bi, bj; tile_coords(blockIdx.x, P.n_tiles_1d, bi, bj);
    stage_tile(P, bi, bj, sG, sC);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int kept = 0;
    for (int pass = 0; pass < 4; ++pass) {
        const int li = pass * 8 + warp, lj = lane;
        const int64_t i = (int64_t)bi * kTile + li, j = (int64_t)bj * kTile + lj;
        bool good = false;
        if (i < j && j < P.n_poses) {
            double Ra[9], Rb[9], ta[3], tb[3];
            make_pair(sG[li], sG[32 + lj], sC[li], sC[32 + lj], Ra, Rb, ta, tb);
            double al[3], be[3]; log_so3_ref(Ra, al); log_so3_ref(Rb, be);
            const double na = sqrt(al[0] * al[0] + al[1] * al[1] + al[2] * al[2]), nb = sqrt(be[0] * be[0] + be[1] * be[1] + be[2] * be[2]);
            good = !(fmin(na, nb) < min_angle);
            if (good && reject_parallel && !(na < 1e-9) && !(nb < 1e-9)) {
                const double a[3] = {al[0] / na, al[1] / na, al[2] / na}, b[3] = {be[0] / nb, be[1] / nb, be[2] / nb};
                const double cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2], cz = a[0] * b[1] - a[1] * b[0];
                if (sqrt(cx * cx + cy * cy + cz * cz) < parallel_eps) good = false;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, good);
        if (lane == 0) { P.mask[(int64_t)blockIdx.x * 32 + li] = m; kept += __popc(m); }
    }
    if (lane == 0 && kept) atomicAdd(count, (unsigned long long)kept);  // integer count: order-independent
}

// Two CTAs per SM (128 registers; ptxas spills 152 bytes in the Jacobian instance, 4 in the cost instance): the pair arithmetic
// is one long dependent FP64 chain per thread, and with 186 registers only 8 warps per SM were there to interleave
// (round-2 capture: FP64 pipe 38 % active, 2.1 "wait" stalls per issue).
template <int JAC>
__global__ void __launch_bounds__(256, CALK_AXXB_MINB) k_axxb_otf(PairTiles P, const double* __restrict__ x7, double huber, double* __restrict__ partial) {
    __shared__ double sG[64][12], sC[64][12];
    __shared__ double sm[8][kAcc];
    const int64_t tile = P.tile_base + blockIdx.x;
    int bi, bj; tile_coords(tile, P.n_tiles_1d, bi, bj);
    stage_tile(P, bi, bj, sG, sC);
    double q[4], tx[3];
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = x7[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) tx[i] = x7[4 + i];
    double Rx[9]; quat_to_R(q, Rx);
    double acc[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) acc[i] = 0.0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int pass = 0; pass < 4; ++pass) {
        const int li = pass * 8 + warp;
        const unsigned m = P.mask[tile * 32 + li];
        if ((m >> lane) & 1u) {
            double Ra[9], Rb[9], ta[3], tb[3];
            make_pair(sG[li], sG[32 + lane], sC[li], sC[32 + lane], Ra, Rb, ta, tb);
            axxb_pair<JAC>(Ra, Rb, ta, tb, Rx, tx, huber, acc);
        }
    }
    cta_reduce_store(acc, sm, partial);
}

// out[e] = sum over CTAs of partial[cta][e]: one warp per entry, lanes stride over the CTAs, fixed shuffle tree
__global__ void __launch_bounds__(32 * kAcc) k_axxb_final(const double* __restrict__ partial, int n_cta, double* __restrict__ out) {
    const int e = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double t = 0.0;
    for (int c = lane; c < n_cta; c += 32) t += partial[(int64_t)c * kAcc + e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane == 0) out[e] = t;
}

// AoS [n][w] -> SoA rows [row0 + j][n]
__global__ void k_axxb_transpose(const double* __restrict__ src, int w, int64_t n, double* __restrict__ dst, int row0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * w) return;
    const int64_t p = i / w; const int j = (int)(i % w);
    dst[(int64_t)(row0 + j) * n + p] = src[i];
}

}  // namespace
