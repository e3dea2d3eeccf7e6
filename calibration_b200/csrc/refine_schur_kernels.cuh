// Device code of K2 — the per-view pose blocks of the intrinsics / extrinsics kinds: gather, Jacobi scaling, batched
// 6x6 Cholesky, Schur complement as a tiled SYRK, back-substitution, per-view plus / norms / reductions (see
// refine_kernels.cu for the launchers).  Kept in a header so that tests/host_emul can run this very source on the
// CPU under the lock-step SIMT shim (test-only); nothing here differs between the two builds.
#pragma once
#include "refine_kernels.cuh"

namespace calk {

// ---------------------------------------------------------------------------
// K2: per-view pose blocks — gather, scaling, batched Cholesky, Schur, back-substitution
// ---------------------------------------------------------------------------
__device__ __forceinline__ int sym6(int i, int j) {
    return i <= j ? i * 6 - i * (i - 1) / 2 + (j - i) : j * 6 - j * (j - 1) / 2 + (i - j);
}
__device__ __forceinline__ int shared_col(const ViewBuffers& V, int cam, int j) {
    // column j of a block's coupling E = [cam pose quat(3) | cam pose tran(3) | intr(PI)]
    const int base = j < 3 ? V.cam_col_q[cam] : (j < 6 ? V.cam_col_t[cam] : V.cam_col_i[cam]);
    return base < 0 ? -1 : base + (j < 3 ? j : (j < 6 ? j - 3 : j - 6));
}

__global__ void k_view_gather(ProblemShape S, DevLayout L, EvalBuffers B, ViewBuffers V) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= S.n_views) return;
    double H[21], g[6];
    for (int i = 0; i < 21; ++i) H[i] = 0.0;
    for (int i = 0; i < 6; ++i) g[i] = 0.0;
    if (V.view_free[v])
        for (int k = V.view_blk_off[v]; k < V.view_blk_off[v + 1]; ++k) {
            const int64_t b = V.view_blk_idx[k];
            for (int i = 0; i < 21; ++i) H[i] += B.blk_Hvv[(int64_t)i * L.n_blk + b];
            for (int i = 0; i < 6; ++i) g[i] += B.blk_gv[(int64_t)i * L.n_blk + b];
        }
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) V.Hpp[(int64_t)v * 36 + 6 * i + j] = H[sym6(i, j)];
        V.gp[(int64_t)v * 6 + i] = g[i];
    }
}

__global__ void k_view_scale(ProblemShape S, ViewBuffers V, int compute_scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n_views * 6) return;
    const int v = i / 6, k = i % 6;
    const double h = V.Hpp[(int64_t)v * 36 + 7 * k];
    double s = V.sp[i];
    if (compute_scale) { s = 1.0 / (1.0 + sqrt(h)); V.sp[i] = s; }  // jacobi_scaling, once
    V.dp[i] = fmin(fmax(h * s * s, 1e-6), 1e32);                      // LM diagonal clamp (B.3-1)
}

// L_v = chol(sp Hpp sp + dp / radius), f_v = L_v^-1 (sp o gp), and for every residual block of the view
// F_b = L_v^-1 (diag(sp) E_b diag(s_shared)), so that E^T A^-1 E = F^T F and E^T A^-1 g = F^T f.
// F goes to the DENSE per-view rows
// Fd[view][6][ncp] (column = shared tangent column, column ns = f_v, zeros where a camera does not see the view — the
// buffer is cleared once at create and the sparsity never changes): 6 x ncp contiguous doubles per view, which is what
// the SYRK stages with one bulk copy and what the back-substitution reads as plain dot products.
// Two kernels.  k_view_chol — one THREAD per view: the 6x6 factorisation is a serial chain of about 600 dependent
// instructions (6 square roots, 6 divisions); one thread per view has every view of the problem in flight at once,
// where the earlier one-warp-per-view kernel repeated that chain in front of every warp's loads with 16 warps per SM
// (0.88 ms at 100 k views x 8 cameras, 25 % issue utilisation, long scoreboard 9.4 per issue).  It publishes L_v, the
// reciprocals of its diagonal, f_v and column ns of the dense rows; a view without blocks publishes nothing (it has no
// rows in the Schur complement).  k_schur_factor — one WARP per view, one lane per (residual block, coupling column):
// L_v and the per-camera column tables are staged in shared memory, the lanes walk the view's columns 32 at a time
// (two trips' gathers in flight), forward-substitute with broadcast reads of L_v and store F into the dense rows with
// consecutive lanes on consecutive columns — full-line writes; the reads of E_b ([entry][n_blk]) are 8-byte gathers
// whose sectors are shared by the warps of the neighbouring views (device blocks are ordered by camera, then view).  No factorisation in this kernel: 64 registers,
// 32 warps per SM.  What bounds it now is the L1 data pipe (70 % of its wavefront rate: a warp-wide gather touches 11
// cache lines) and instruction issue, at 2.7 TB/s of DRAM traffic.  A variant that transposes E through shared memory
// (one plane per warp-wide load, a group of 32 / n_cams views per CTA) was measured on the B200 and is 40 % SLOWER
// (588 against 418 us at 100 k views x 8 cameras): the staging phases and their barriers cost more instructions than
// the coalesced loads save.  So is a block-fastest item order with the rows assembled in a per-warp shared-memory tile
// and written out whole: 632 us.  Both assumed a view's blocks to be neighbours in memory; they are n_views apart (the
// device order is camera-major for K1's per-camera reductions), so only lanes on CONSECUTIVE VIEWS of one camera read
// consecutive doubles — a one-lane-per-view organisation, which in turn needs L_v plane-major and scatters the stores.
// (A warp-per-view variant with the matrix spread over the lanes — one coalesced load, the factorisation by shuffles, no
// local-memory frame — produces the same bits and takes the same time on the B200: 82 against 78 us at 100 k views.)
__global__ void __launch_bounds__(128) k_view_chol(ProblemShape S, ViewBuffers V, double inv_radius) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= S.n_views || !V.view_free[v]) return;
    if (V.view_blk_off[v + 1] == V.view_blk_off[v]) return;
    double A[36], sp[6];
    for (int i = 0; i < 6; ++i) sp[i] = V.sp[(int64_t)v * 6 + i];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) A[6 * i + j] = V.Hpp[(int64_t)v * 36 + 6 * i + j] * sp[i] * sp[j];
    for (int i = 0; i < 6; ++i) A[7 * i] += V.dp[(int64_t)v * 6 + i] * inv_radius;
    if (!chol6(A)) { atomicExch(V.fail, 1); return; }
    for (int i = 0; i < 36; ++i) V.Lp[(int64_t)v * 36 + i] = A[i];
    for (int i = 0; i < 6; ++i) V.Linv[(int64_t)v * 6 + i] = 1.0 / A[7 * i];
    double f[6];
    for (int i = 0; i < 6; ++i) f[i] = V.gp[(int64_t)v * 6 + i] * sp[i];
    for (int i = 0; i < 6; ++i) { double s = f[i]; for (int k = 0; k < i; ++k) s -= A[6 * i + k] * f[k]; f[i] = s / A[7 * i]; }
    for (int i = 0; i < 6; ++i) { V.view_f[(int64_t)v * 6 + i] = f[i]; V.Fd[((int64_t)v * 6 + i) * V.ncp + V.ns] = f[i]; }
}

constexpr int kFactorThreads = 256;
constexpr int kFactorCamTable = 64;    // cameras whose column tables fit the shared-memory copy (more: read from global memory)
__global__ void __launch_bounds__(kFactorThreads) k_schur_factor(ProblemShape S, DevLayout L, EvalBuffers B, ViewBuffers V) {
    __shared__ double sL[kFactorThreads / 32][28];                    // per warp: L_v (21, packed rows) | 1 / diag (6)
    __shared__ double s_scale[kSyrkMaxN];                             // s_shared
    __shared__ int s_col[3][kFactorCamTable];                         // cam_col_q / _t / _i
    for (int i = threadIdx.x; i < V.ns; i += kFactorThreads) s_scale[i] = V.s_shared[i];
    for (int i = threadIdx.x; i < min(S.n_cams, kFactorCamTable); i += kFactorThreads) {
        s_col[0][i] = V.cam_col_q[i]; s_col[1][i] = V.cam_col_t[i]; s_col[2][i] = V.cam_col_i[i];
    }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int v = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (v >= S.n_views || !V.view_free[v]) return;
    const int k0 = V.view_blk_off[v], nb = V.view_blk_off[v + 1] - k0;
    if (nb == 0) return;
    {   // lane l < 21: entry (ri, rj) of the lower triangle; lanes 21..26: the reciprocal diagonal
        const int ri = (lane >= 1) + (lane >= 3) + (lane >= 6) + (lane >= 10) + (lane >= 15), rj = lane - ri * (ri + 1) / 2;
        if (lane < 21) sL[w][lane] = V.Lp[(int64_t)v * 36 + 6 * ri + rj];
        else if (lane < 27) sL[w][lane] = V.Linv[(int64_t)v * 6 + lane - 21];
    }
    double sp[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) sp[i] = V.sp[(int64_t)v * 6 + i];
    __syncwarp();
    const double* const Lw = sL[w];
    const int ncb = 6 + S.PI, n_items = nb * ncb;
    const bool cam_table = S.n_cams <= kFactorCamTable;
    double* const Fv = V.Fd + (int64_t)v * 6 * V.ncp;
    for (int t0 = lane; t0 < n_items; t0 += 64) {
        double e[2][6];
        int col[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int t = t0 + 32 * u;
            const bool live = t < n_items;
            const int k = live ? t / ncb : 0, j = live ? t - k * ncb : 0;
            const int64_t b = V.view_blk_idx[k0 + k];
#pragma unroll
            for (int i = 0; i < 6; ++i)
                e[u][i] = j < 6 ? B.blk_Evc[(int64_t)(6 * i + j) * L.n_blk + b] : B.blk_Evi[(int64_t)(S.PI * i + j - 6) * L.n_blk + b];
            const int cam = L.blk_cam[b];
            int c;
            if (cam_table) { const int base = s_col[j < 3 ? 0 : (j < 6 ? 1 : 2)][cam]; c = base < 0 ? -1 : base + (j < 3 ? j : (j < 6 ? j - 3 : j - 6)); }
            else c = shared_col(V, cam, j);
            col[u] = live ? c : -1;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (col[u] < 0) continue;
            const double sc = s_scale[col[u]];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double s = e[u][i] * sp[i] * sc;
#pragma unroll
                for (int kk = 0; kk < i; ++kk) s -= Lw[i * (i + 1) / 2 + kk] * e[u][kk];
                e[u][i] = s * Lw[21 + i];
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) Fv[i * V.ncp + col[u]] = e[u][i];
        }
    }
}

// The reduced (shared-block) system of one LM iteration, solved where its inputs are: one CTA forms
//     S = Sm - C,  rhs = gss - c      (Sm: damped Jacobi-scaled H_ss from the host, [ns][ns]; C, c: the Schur complement)
// in shared memory, factors it by a BLOCKED right-looking Cholesky and solves for y (k_backsub reads y_shared next, in
// stream order: no round trip to the host between the Schur complement and the back-substitution).  info[0] = 1 if the
// matrix was not positive definite or y is not finite.  n <= kReducedMaxN.
// The right-hand side rides along as row n of the matrix, so the forward substitution is part of the factorisation
// (row n of the factor is L^-1 rhs).  Per panel of kRedNB columns: (1) one THREAD per row keeps its kRedNB panel entries
// in registers and the columns are finished left-looking, the row of the diagonal block broadcast through shared memory
// — one barrier per column; (2) the trailing matrix gets the rank-kRedNB update from the transposed panel in interleaved
// 8x8 register tiles, one per thread (the k_schur_syrk mapping: conflict-free rows), one barrier per panel.  The
// column-at-a-time version this replaces spent 3 barriers and a dependent load-FMA-store chain per column: 229 us at
// n = 114 (round-2 capture), barrier stalls 8 per issue.  Back substitution: panels in reverse, the diagonal block by
// one warp with shuffles, the rest one thread per unknown.
// Shared memory: (n + 1) ld + kRedNB kRedPtLd + kRedNB^2 + kRedNB doubles, ld = n | 1 (odd: column walks hit distinct banks).
constexpr int kRedNB = 16;
constexpr int kRedPtLd = (kReducedMaxN + 1 - 1 + 7) / 8 * 8;   // the widest trailing block (n rows after a one-column panel), padded to the tile
__host__ __device__ inline size_t reduced_solve_smem_doubles(int n) {
    return (size_t)(n + 1) * (n | 1) + (size_t)kRedNB * kRedPtLd + kRedNB * kRedNB + kRedNB;
}
__global__ void __launch_bounds__(256) k_reduced_solve(const double* __restrict__ Sm, const double* __restrict__ gss, ViewBuffers V, int n,
                                                       int32_t* __restrict__ info) {
#if defined(__CUDACC__)
    extern __shared__ __align__(16) double red_sm[];
    __shared__ int bad;
#else   // host build of this source (tests/host_emul): one CTA runs at a time
    static double red_sm[(kReducedMaxN + 1) * (kReducedMaxN | 1) + kRedNB * kRedPtLd + kRedNB * kRedNB + kRedNB];
    static int bad;
#endif
    const int ld = n | 1;
    double* const A = red_sm;                             // [n + 1][ld]: lower triangle of S, row n = right-hand side
    double* const Pt = A + (size_t)(n + 1) * ld;          // [kRedNB][kRedPtLd]: the panel below its diagonal block, transposed
    double* const dg = Pt + (size_t)kRedNB * kRedPtLd;    // [kRedNB][kRedNB]: rows of the panel's diagonal block
    double* const dgi = dg + kRedNB * kRedNB;             // [kRedNB]: reciprocals of its diagonal
    double* const z = A + (size_t)n * ld;                 // row n
    const int tid = threadIdx.x;
    if (tid == 0) bad = 0;
    for (int i = tid; i < n * n; i += 256) { const int r = i / n, c = i - r * n; if (c <= r) A[(size_t)r * ld + c] = Sm[i] - V.C[i]; }
    for (int i = tid; i < n; i += 256) z[i] = gss[i] - V.c[i];
    __syncthreads();
    for (int j0 = 0; j0 < n; j0 += kRedNB) {
        const int nb = min(kRedNB, n - j0);
        const int rows = n + 1 - j0;                      // rows j0 .. n, one thread each
        const bool has_row = tid < rows;
        double row[kRedNB];
#pragma unroll
        for (int c = 0; c < kRedNB; ++c) row[c] = (has_row && c < nb && (c <= tid || tid >= nb)) ? A[(size_t)(j0 + tid) * ld + j0 + c] : 0.0;
#pragma unroll
        for (int c = 0; c < kRedNB; ++c) {
            if (c < nb) {
                if (tid == c) {
                    double s = row[c];
#pragma unroll
                    for (int k = 0; k < c; ++k) s = fma(-row[k], row[k], s);
                    if (!(s > 0.0) || !isfinite(s)) bad = 1;
                    const double d = sqrt(s);
                    row[c] = d;
#pragma unroll
                    for (int k = 0; k <= c; ++k) dg[c * kRedNB + k] = row[k];
                    dgi[c] = 1.0 / d;
                }
                __syncthreads();
                if (has_row && tid > c && !bad) {
                    double s = row[c];
#pragma unroll
                    for (int k = 0; k < c; ++k) s = fma(-row[k], dg[c * kRedNB + k], s);
                    row[c] = s * dgi[c];
                }
            }
        }
        if (bad) break;                                   // (uniform: read after a barrier, written before it)
        // the factor's columns go back (the diagonal holds 1 / L_jj: only the back substitution reads it again), and
        // the rows below the diagonal block go to the transposed panel, zero-padded to the tile grid
        const int m = rows - nb, nt = (m + 7) / 8;        // trailing rows (the right-hand side is the last one)
#pragma unroll
        for (int c = 0; c < kRedNB; ++c) {
            if (c < nb) {
                if (has_row && (c <= tid || tid >= nb)) A[(size_t)(j0 + tid) * ld + j0 + c] = (tid == c) ? dgi[c] : row[c];
                if (tid >= nb && tid < nb + nt * 8) Pt[c * kRedPtLd + tid - nb] = row[c];   // (row[] is zero past the last row)
            }
        }
        __syncthreads();
        if (m > 1) {
            int ti = -1, tj = -1;
            { int t = tid, r = 0; while (r < nt && t >= nt - r) { t -= nt - r; ++r; } if (r < nt) { ti = r; tj = r + t; } }
            if (ti >= 0) {
                double acc[8][8];
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;
                for (int c = 0; c < nb; ++c) {
                    double fa[8], fb[8];
#pragma unroll
                    for (int a = 0; a < 8; ++a) { fa[a] = Pt[c * kRedPtLd + ti + nt * a]; fb[a] = Pt[c * kRedPtLd + tj + nt * a]; }
#pragma unroll
                    for (int a = 0; a < 8; ++a)
#pragma unroll
                        for (int b = 0; b < 8; ++b) acc[a][b] = fma(fa[a], fb[b], acc[a][b]);
                }
                double* const T = A + (size_t)(j0 + nb) * ld + j0 + nb;   // the trailing block
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        const int r = ti + nt * a, cc = tj + nt * b;
                        const int hi = max(r, cc), lo = min(r, cc);
                        if (hi < m && lo < m - 1 && (ti != tj || a >= b)) T[(size_t)hi * ld + lo] -= acc[a][b];
                    }
            }
        }
        __syncthreads();
    }
    if (!bad) {   // L^T y = z, panels in reverse
        for (int j0 = (n - 1) / kRedNB * kRedNB; j0 >= 0; j0 -= kRedNB) {
            const int nb = min(kRedNB, n - j0);
            if (tid < 32) {
                double x = tid < nb ? z[j0 + tid] : 0.0;
                for (int c = nb - 1; c >= 0; --c) {
                    if (tid == c) x *= A[(size_t)(j0 + c) * ld + j0 + c];   // (the stored reciprocal)
                    const double xc = __shfl_sync(0xffffffffu, x, c);
                    if (tid < c) x = fma(-A[(size_t)(j0 + c) * ld + j0 + tid], xc, x);
                }
                if (tid < nb) z[j0 + tid] = x;
            }
            __syncthreads();
            if (tid < j0) {
                double s = z[tid];
                for (int c = 0; c < nb; ++c) s = fma(-A[(size_t)(j0 + c) * ld + tid], z[j0 + c], s);
                z[tid] = s;
            }
            __syncthreads();
        }
        for (int i = tid; i < n; i += 256) if (!isfinite(z[i])) bad = 1;
    }
    __syncthreads();
    for (int i = tid; i < n; i += 256) V.y_shared[i] = z[i];
    if (tid == 0) info[0] = bad;
}

// C_aug = sum_v F_v^T F_v — a SYRK with a long inner dimension (6 rows per view, 600 000 rows at 100 k views) and a
// small output ((ns+1)^2, ns + 1 <= 176): the one GEMM-shaped operation of the path, on the FP64 TENSOR-CORE instruction
// (mma.sync m8n8k4 f64, SASS DMMA.8x8x4).  On B200 DMMA and DFMA share one FP64 roof (tools/ubench_dmma.cu: 37.0 against
// 36.3 TFLOP/s, no co-issue), so the gain is not a higher peak but what a DMMA does not need: one instruction per 256
// multiply-adds and two 8-byte operands per lane, where the register-tiled DFMA version issued 16 shared-memory loads
// per 64 DFMAs and stopped at 50 % of the FP64 pipe (an LDS costs the FP64 pipe 1.6 DFMA issue slots).
// The upper triangle of the output is cut into blocks of BT x BT tiles of 8x8; a WARP owns one block (its accumulators:
// 2 BT^2 doubles per lane) and per k-step of four rows loads BT + BT operand fragments for BT^2 DMMAs.  The rows
// [F_v | f_v] of kSyrkViews consecutive views are ONE contiguous block of Fd, staged by a TMA bulk copy
// (cp.async.bulk + mbarrier) into a two-stage shared-memory ring: the copy of the next views runs under the DMMAs of
// the current ones.  Per-CTA partial results are summed in a fixed order by k_schur_reduce (no floating-point atomics).
// Dynamic shared memory: 2 * kSyrkViews * 6 * ncp doubles + 2 mbarriers.
constexpr int kSyrkViews = 4;
// block edge (in 8x8 tiles) and warps per CTA for a shared block of ns columns
__host__ __device__ inline void syrk_shape(int ns, int* bt, int* warps) {
    const int nt8 = (ns + 1 + 7) / 8;
    const int b = nt8 <= 12 ? 3 : 5, nbk = (nt8 + b - 1) / b;   // at most 10 warps up to ns + 1 = 160, 15 warps above (the <5, 480> instance)
    *bt = b; *warps = nbk * (nbk + 1) / 2;
}
// D(8x8) += A(8x4) B(4x8): lane l holds A[l >> 2][l & 3], B[l & 3][l >> 2] and D[l >> 2][2 (l & 3) + {0, 1}]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
#if defined(__CUDACC__)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
#else   // host build (tests/host_emul): the same contraction from shuffles
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int k = 0; k < 4; ++k) {
        const double ak = __shfl_sync(0xffffffffu, a, g * 4 + k);
        const double b0 = __shfl_sync(0xffffffffu, b, (2 * t) * 4 + k), b1 = __shfl_sync(0xffffffffu, b, (2 * t + 1) * 4 + k);
        d0 = fma(ak, b0, d0); d1 = fma(ak, b1, d1);
    }
#endif
}
template <int BT, int MAXT>
__global__ void __launch_bounds__(MAXT) k_schur_syrk(ProblemShape S, ViewBuffers V, int ns, int views_per_cta) {
#if defined(__CUDACC__)
    extern __shared__ __align__(128) unsigned char syrk_smem[];
#else   // host build of this source (tests/host_emul): one CTA runs at a time
    static __attribute__((aligned(128))) unsigned char syrk_smem[2 * kSyrkViews * 6 * kSyrkMaxN * 8 + 16];
#endif
    const int ncp = V.ncp;
    const int stage_doubles = kSyrkViews * 6 * ncp;
    double* const ring = reinterpret_cast<double*>(syrk_smem);
    unsigned long long* const bar = reinterpret_cast<unsigned long long*>(syrk_smem + (size_t)2 * stage_doubles * 8);
    const int na = ns + 1, nt8 = (na + 7) / 8, nbk = (nt8 + BT - 1) / BT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    int bi = 0, bj = 0;                                   // this warp's block of the upper triangle
    { int w = warp; while (bi < nbk && w >= nbk - bi) { w -= nbk - bi; ++bi; } bj = bi + w; }
    int ca[BT], cb[BT];                                   // column of lane's operand in each tile (tiles past the edge repeat the last one: computed, never stored)
#pragma unroll
    for (int a = 0; a < BT; ++a) { ca[a] = min(bi * BT + a, nt8 - 1) * 8 + g; cb[a] = min(bj * BT + a, nt8 - 1) * 8 + g; }
    double acc[BT][BT][2];
#pragma unroll
    for (int a = 0; a < BT; ++a)
#pragma unroll
        for (int b = 0; b < BT; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    const int v0 = blockIdx.x * views_per_cta, v1 = min(S.n_views, v0 + views_per_cta);
    const int n_steps = v1 > v0 ? (v1 - v0 + kSyrkViews - 1) / kSyrkViews : 0;
    // stage `k & 1` <- the rows of the views of step k (leader thread only)
    auto issue = [&](int k) {
        if (k >= n_steps) return;
        const int vb = v0 + k * kSyrkViews, nvb = min(kSyrkViews, v1 - vb);
        const double* src = V.Fd + (int64_t)vb * 6 * ncp;
        double* dst = ring + (k & 1) * stage_doubles;
        const unsigned bytes = (unsigned)(nvb * 6 * ncp * 8);
#if defined(__CUDACC__)
        const unsigned mb = (unsigned)__cvta_generic_to_shared(&bar[k & 1]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"(mb)
                     : "memory");
#else
        for (unsigned i = 0; i < bytes / 8; ++i) dst[i] = src[i];
#endif
    };
#if defined(__CUDACC__)
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
#endif
    __syncthreads();
    if (threadIdx.x == 0) { issue(0); issue(1); }
#if !defined(__CUDACC__)
    __syncthreads();
#endif
    for (int k = 0; k < n_steps; ++k) {
        const int nvb = min(kSyrkViews, v1 - (v0 + k * kSyrkViews));
        double* const frow = ring + (k & 1) * stage_doubles;
        if (nvb * 6 % 4 != 0) {   // a last, partial step of 1 or 3 views: zero the two rows that complete its last k-step (not covered by the copy)
            for (int i = threadIdx.x; i < 2 * ncp; i += blockDim.x) frow[nvb * 6 * ncp + i] = 0.0;
            __syncthreads();
        }
#if defined(__CUDACC__)
        {
            const unsigned mb = (unsigned)__cvta_generic_to_shared(&bar[k & 1]);
            const unsigned parity = (unsigned)(k >> 1) & 1u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@!p bra WAIT_%=;\n\t}" ::"r"(mb), "r"(parity) : "memory");
        }
#endif
        const int ksteps = (nvb * 6 + 3) / 4;
        // (Loading the fragments of k-step ks + 1 before the DMMAs of k-step ks — explicit double buffering — was measured on the
        // B200: 0.358 against 0.337 ms; ptxas already spreads the ten loads among the DMMAs of the step.)
        if (bi != bj) {
            for (int ks = 0; ks < ksteps; ++ks) {
                const double* const row = frow + (4 * ks + t) * ncp;
                double fa[BT], fb[BT];
#pragma unroll
                for (int a = 0; a < BT; ++a) { fa[a] = row[ca[a]]; fb[a] = row[cb[a]]; }
#pragma unroll
                for (int a = 0; a < BT; ++a)
#pragma unroll
                    for (int b = 0; b < BT; ++b) dmma884(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
            }
        } else {   // a block on the diagonal: both operands are the same fragments, and only its tiles a <= b are stored
            for (int ks = 0; ks < ksteps; ++ks) {
                const double* const row = frow + (4 * ks + t) * ncp;
                double fa[BT];
#pragma unroll
                for (int a = 0; a < BT; ++a) fa[a] = row[ca[a]];
#pragma unroll
                for (int a = 0; a < BT; ++a)
#pragma unroll
                    for (int b = a; b < BT; ++b) dmma884(acc[a][b][0], acc[a][b][1], fa[a], fa[b]);
            }
        }
        __syncthreads();                               // every warp is done with this stage
        if (threadIdx.x == 0) issue(k + 2);            // refill it
#if !defined(__CUDACC__)
        __syncthreads();
#endif
    }
    double* out = V.partialC + (int64_t)blockIdx.x * na * na;
#pragma unroll
    for (int a = 0; a < BT; ++a)
#pragma unroll
        for (int b = 0; b < BT; ++b) {
            const int ta = bi * BT + a, tb = bj * BT + b;
            if (ta >= nt8 || tb >= nt8 || ta > tb) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = ta * 8 + g, c = tb * 8 + 2 * t + h;
                if (r < na && c < na) { out[(int64_t)r * na + c] = acc[a][b][h]; if (ta != tb) out[(int64_t)c * na + r] = acc[a][b][h]; }
            }
        }
}

#if !defined(__CUDACC__)
// host build (tests/host_emul): the body a launcher runs per thread, with launch_schur's dispatch on the block edge
inline void k_schur_syrk_any(int bt, ProblemShape S, ViewBuffers V, int ns, int views_per_cta) {
    if (bt == 3) k_schur_syrk<3, 320>(S, V, ns, views_per_cta);
    else k_schur_syrk<5, 320>(S, V, ns, views_per_cta);   // (the <5, 480> instance is the same source under another register cap)
}
#endif

// The per-CTA partials of the SYRK summed in CTA order.  One thread per entry walking 444 partials is a chain of 444
// dependent additions behind loads issued one at a time (113 us at ns = 114, round 2): four threads share an entry, each
// sums a quarter of the partials with eight loads in flight, and the quarters are added in order — a fixed order, so
// the result does not depend on scheduling.
constexpr int kSchurReduceEntries = 32;   // entries per CTA (x 4 quarter sums = 128 threads)
__global__ void __launch_bounds__(4 * kSchurReduceEntries) k_schur_reduce(ViewBuffers V, int n_cta, int ns) {
    __shared__ double part[4][kSchurReduceEntries];
    const int na = ns + 1;
    const int e = threadIdx.x % kSchurReduceEntries, qd = threadIdx.x / kSchurReduceEntries;
    const int i = blockIdx.x * kSchurReduceEntries + e;
    const int per = (n_cta + 3) / 4, c0 = qd * per, c1 = min(n_cta, c0 + per);
    double s = 0.0;
    if (i < na * na) {
        const double* p = V.partialC + i;
        const int64_t stride = (int64_t)na * na;
        int c = c0;
        for (; c + 8 <= c1; c += 8) {
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = p[(c + u) * stride];
#pragma unroll
            for (int u = 0; u < 8; ++u) s += t[u];
        }
        for (; c < c1; ++c) s += p[c * stride];
    }
    part[qd][e] = s;
    __syncthreads();
    if (qd != 0 || i >= na * na) return;
    s = ((part[0][e] + part[1][e]) + part[2][e]) + part[3][e];
    const int r = i / na, cc = i % na;
    if (r < ns && cc < ns) V.C[(int64_t)r * ns + cc] = s;
    else if (r < ns && cc == ns) V.c[r] = s;
}

// y_p = L^-T (f - F_v y_s), step_p = -y_p, delta_p = step_p o sp, and the per-view terms of step'g and
// step'H step (H undamped, Jacobi-scaled).  A CTA takes kBacksubViews = 32 consecutive views.  Phase 1: each of the 8
// warps forms q = F_v y_s for FOUR views at once — six dot products per view over the dense rows, lanes striding over
// the shared columns (coalesced), 24 independent loads in flight per lane — summed by a fixed shuffle tree and left in
// shared memory.  Phase 2: warp 0, one LANE per view, finishes the 6x6 part for the 32 views side by side.  (One warp
// per view with lane 0 walking the serial 6x6 part kept 16 warp slots per SM waiting on one lane's dependent chain:
// 0.27 ms = 2.1 TB/s for a 576 MB read, round 2.)
constexpr int kBacksubViews = 32;
__global__ void __launch_bounds__(256) k_backsub(ProblemShape S, DevLayout L, ViewBuffers V, int ns) {
    __shared__ double sq[kBacksubViews][6];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int vbase = blockIdx.x * kBacksubViews;
    {
        const int v0 = vbase + 4 * w;                      // this warp's four views
        double q[4][6];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int i = 0; i < 6; ++i) q[a][i] = 0.0;
        const double* F[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) F[a] = V.Fd + (int64_t)min(v0 + a, S.n_views - 1) * 6 * V.ncp;   // (clamped: a view past the end is not used)
        for (int c = lane; c < ns; c += 32) {
            const double ys = V.y_shared[c];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int i = 0; i < 6; ++i) q[a][i] = fma(F[a][i * V.ncp + c], ys, q[a][i]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int i = 0; i < 6; ++i) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) q[a][i] += __shfl_xor_sync(0xffffffffu, q[a][i], o);
                if (lane == 0) sq[4 * w + a][i] = q[a][i];
            }
    }
    __syncthreads();
    if (w != 0) return;
    const int v = vbase + lane;
    if (v >= S.n_views) return;
    double* red = V.red + (int64_t)v * 4;
    if (!V.view_free[v] || V.view_blk_off[v + 1] == V.view_blk_off[v]) {   // fixed, or seen by no camera: no step
        for (int i = 0; i < 6; ++i) V.delta_p[(int64_t)v * 6 + i] = 0.0;
        red[0] = red[1] = 0.0;
        return;
    }
    double q[6];
    for (int i = 0; i < 6; ++i) q[i] = sq[lane][i];
    double Lm[36], f[6], sp[6];
    for (int i = 0; i < 36; ++i) Lm[i] = V.Lp[(int64_t)v * 36 + i];
    for (int i = 0; i < 6; ++i) { f[i] = V.view_f[(int64_t)v * 6 + i]; sp[i] = V.sp[(int64_t)v * 6 + i]; }
    double y[6];
    for (int i = 0; i < 6; ++i) y[i] = f[i] - q[i];
    for (int i = 5; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < 6; ++k) s -= Lm[6 * k + i] * y[k]; y[i] = s / Lm[7 * i]; }
    double step[6], sg = 0.0;
    for (int i = 0; i < 6; ++i) { step[i] = -y[i]; sg += step[i] * V.gp[(int64_t)v * 6 + i] * sp[i]; V.delta_p[(int64_t)v * 6 + i] = step[i] * sp[i]; }
    // quad = step' A0 step + 2 step' E_s step_s ; E_s step_s = -L q
    double quad = 0.0;
    for (int i = 0; i < 6; ++i) {
        double row = 0.0;
        for (int j = 0; j < 6; ++j) row += V.Hpp[(int64_t)v * 36 + 6 * i + j] * sp[j] * step[j];
        double lq = 0.0;
        for (int k = 0; k <= i; ++k) lq += Lm[6 * i + k] * q[k];
        quad += step[i] * (sp[i] * row - 2.0 * lq);
    }
    red[0] = sg; red[1] = quad;
}

// x_cand(view v) = x(view v) [+] t * delta_p ; red[2] = |dx|^2
__global__ void k_view_plus(ProblemShape S, EvalBuffers B, ViewBuffers V, double t) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= S.n_views) return;
    const double* q = B.x + S.off_viewq + 4 * (int64_t)v; const double* tr = B.x + S.off_viewt + 3 * (int64_t)v;
    double* qo = V.x_cand + S.off_viewq + 4 * (int64_t)v; double* to = V.x_cand + S.off_viewt + 3 * (int64_t)v;
    double d[6];
    for (int i = 0; i < 6; ++i) d[i] = t * V.delta_p[(int64_t)v * 6 + i];
    double qn[4]; quat_plus(q, d, qn);
    double dx2 = 0.0;
    for (int i = 0; i < 4; ++i) { qo[i] = qn[i]; const double e = qn[i] - q[i]; dx2 += e * e; }
    for (int i = 0; i < 3; ++i) { const double e = d[3 + i]; to[i] = tr[i] + e; dx2 += (to[i] - tr[i]) * (to[i] - tr[i]); }
    V.red[(int64_t)v * 4 + 2] = dx2;
}

// red[0] = |x_v|^2, red[3] = max |x_v - plus(x_v, -g_v)| (gradient max-norm piece)
__global__ void k_view_norms(ProblemShape S, EvalBuffers B, ViewBuffers V) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= S.n_views) return;
    const double* q = B.x + S.off_viewq + 4 * (int64_t)v; const double* tr = B.x + S.off_viewt + 3 * (int64_t)v;
    double x2 = 0.0, gm = 0.0;
    for (int i = 0; i < 4; ++i) x2 += q[i] * q[i];
    for (int i = 0; i < 3; ++i) x2 += tr[i] * tr[i];
    if (V.view_free[v]) {
        double d[6]; for (int i = 0; i < 6; ++i) d[i] = -V.gp[(int64_t)v * 6 + i];
        double qn[4]; quat_plus(q, d, qn);
        for (int i = 0; i < 4; ++i) gm = fmax(gm, fabs(q[i] - qn[i]));
        for (int i = 0; i < 3; ++i) gm = fmax(gm, fabs(tr[i] - (tr[i] + d[3 + i])));
    }
    V.red[(int64_t)v * 4 + 0] = x2; V.red[(int64_t)v * 4 + 3] = gm;
}

// red_out[0..2] = column sums, red_out[3] = column max.  Every CTA reduces a contiguous slice of the views with a fixed
// tree; the CTA that arrives last (integer ticket) combines the per-CTA partials in CTA order: deterministic, one launch.
__global__ void __launch_bounds__(256) k_reduce_views(ViewBuffers V, int n_views) {
    __shared__ double sm[4][256];
#if defined(__CUDACC__)
    __shared__ int last;
#else
    static int last;   // host build of this source (tests/host_emul): one CTA runs at a time
#endif
    const int per = (n_views + gridDim.x - 1) / gridDim.x;
    const int v0 = blockIdx.x * per, v1 = min(n_views, v0 + per);
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (int v = v0 + threadIdx.x; v < v1; v += 256) {
        const double* r = V.red + (int64_t)v * 4;
        a[0] += r[0]; a[1] += r[1]; a[2] += r[2]; a[3] = fmax(a[3], r[3]);
    }
    for (int k = 0; k < 4; ++k) sm[k][threadIdx.x] = a[k];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            for (int k = 0; k < 3; ++k) sm[k][threadIdx.x] += sm[k][threadIdx.x + s];
            sm[3][threadIdx.x] = fmax(sm[3][threadIdx.x], sm[3][threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x < 4) V.red_part[blockIdx.x * 4 + threadIdx.x] = sm[threadIdx.x][0];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(V.red_ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < 4) {
        double r = 0.0;
        for (unsigned c = 0; c < gridDim.x; ++c) {
            const double p = __ldcg(V.red_part + c * 4 + threadIdx.x);
            r = threadIdx.x == 3 ? fmax(r, p) : r + p;
        }
        V.red_out[threadIdx.x] = r;
    }
    if (threadIdx.x == 0) *V.red_ticket = 0u;   // ready for the next launch
}

// ---- block-structured covariance of the per-view kinds (description and launcher: refine_kernels.cu) ----
__global__ void __launch_bounds__(64) k_cov_view_prep(ProblemShape S, DevLayout L, ViewBuffers V, int ns, const double* __restrict__ W,
                                                      double* __restrict__ Z, double* __restrict__ G, double* __restrict__ Ainv) {
    __shared__ double Y[6][kSyrkMaxN];
    __shared__ double Lm[36];
    const int v = blockIdx.x;
    if (!V.view_free[v]) return;  // uniform
    for (int i = threadIdx.x; i < 6 * kSyrkMaxN; i += 64) (&Y[0][0])[i] = 0.0;
    if (threadIdx.x < 36) Lm[threadIdx.x] = V.Lp[(int64_t)v * 36 + threadIdx.x];
    __syncthreads();
    for (int idx = threadIdx.x; idx < 6 * ns; idx += 64) {   // Y = L_v^-1 E_v: the dense rows of the view (k_schur_factor)
        const int i = idx / ns, c = idx % ns;
        Y[i][c] = V.Fd[((int64_t)v * 6 + i) * V.ncp + c];
    }
    __syncthreads();
    // Z = L^-T Y, column by column
    for (int c = threadIdx.x; c < ns; c += 64) {
        double z[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) z[i] = Y[i][c];
#pragma unroll
        for (int i = 5; i >= 0; --i) { double s = z[i];
#pragma unroll
            for (int k = i + 1; k < 6; ++k) s -= Lm[6 * k + i] * z[k]; z[i] = s / Lm[7 * i]; }
#pragma unroll
        for (int i = 0; i < 6; ++i) { Y[i][c] = z[i]; Z[((int64_t)v * 6 + i) * ns + c] = z[i]; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < ns; c += 64) {
        double g[6] = {0, 0, 0, 0, 0, 0};
        for (int k = 0; k < ns; ++k) {
            const double w = W[(int64_t)k * ns + c];
#pragma unroll
            for (int i = 0; i < 6; ++i) g[i] = fma(Y[i][k], w, g[i]);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) G[((int64_t)v * 6 + i) * ns + c] = g[i];
    }
    if (threadIdx.x < 6) {  // column j of A^-1 = L^-T L^-1 e_j
        const int j = threadIdx.x;
        double e[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) e[i] = i == j ? 1.0 : 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) { double s = e[i];
#pragma unroll
            for (int k = 0; k < i; ++k) s -= Lm[6 * i + k] * e[k]; e[i] = s / Lm[7 * i]; }
#pragma unroll
        for (int i = 5; i >= 0; --i) { double s = e[i];
#pragma unroll
            for (int k = i + 1; k < 6; ++k) s -= Lm[6 * k + i] * e[k]; e[i] = s / Lm[7 * i]; }
#pragma unroll
        for (int i = 0; i < 6; ++i) Ainv[(int64_t)v * 36 + 6 * i + j] = e[i];
    }
}

constexpr int kCovTile = 16, kCovK = 16;
__global__ void __launch_bounds__(kCovTile * kCovTile) k_cov_vv(ProblemShape S, ViewBuffers V, const double* __restrict__ x, int ns,
                                                                  const double* __restrict__ Z, const double* __restrict__ G,
                                                                  const double* __restrict__ Ainv, double* __restrict__ cov, int64_t na) {
    __shared__ double sG[kCovTile][6][kCovK + 1], sZ[kCovTile][6][kCovK + 1];
    const int tv = threadIdx.x / kCovTile, tw = threadIdx.x % kCovTile;
    const int v = blockIdx.y * kCovTile + tv, w = blockIdx.x * kCovTile + tw;
    double c[36];
#pragma unroll
    for (int i = 0; i < 36; ++i) c[i] = 0.0;
    for (int k0 = 0; k0 < ns; k0 += kCovK) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < kCovTile * 6 * kCovK; idx += kCovTile * kCovTile) {
            const int t = idx / (6 * kCovK), rem = idx % (6 * kCovK), i = rem / kCovK, k = rem % kCovK;
            const int vv = blockIdx.y * kCovTile + t, ww = blockIdx.x * kCovTile + t;
            sG[t][i][k] = (vv < S.n_views && k0 + k < ns && V.view_free[vv]) ? G[((int64_t)vv * 6 + i) * ns + k0 + k] : 0.0;
            sZ[t][i][k] = (ww < S.n_views && k0 + k < ns && V.view_free[ww]) ? Z[((int64_t)ww * 6 + i) * ns + k0 + k] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < kCovK; ++k) {
            double g[6], z[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) { g[i] = sG[tv][i][k]; z[i] = sZ[tw][i][k]; }
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) c[6 * i + j] = fma(g[i], z[j], c[6 * i + j]);
        }
    }
    if (v >= S.n_views || w >= S.n_views || !V.view_free[v] || !V.view_free[w]) return;  // constant blocks: zero rows / columns
    if (v == w) {
#pragma unroll
        for (int i = 0; i < 36; ++i) c[i] += Ainv[(int64_t)v * 36 + i];
    }
    double sv[6], sw[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { sv[i] = V.sp[(int64_t)v * 6 + i]; sw[i] = V.sp[(int64_t)w * 6 + i]; }
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) c[6 * i + j] *= sv[i] * sw[j];
    // ambient lift: rows / columns 0..2 through the quaternion plus-Jacobian (4x3), 3..5 unchanged
    const double* qv = x + S.off_viewq + 4 * (int64_t)v; const double* qw = x + S.off_viewq + 4 * (int64_t)w;
    const double Pv[12] = {-qv[1], -qv[2], -qv[3], qv[0], qv[3], -qv[2], -qv[3], qv[0], qv[1], qv[2], -qv[1], qv[0]};
    const double Pw[12] = {-qw[1], -qw[2], -qw[3], qw[0], qw[3], -qw[2], -qw[3], qw[0], qw[1], qw[2], -qw[1], qw[0]};
    double R[7][6];  // rows lifted: [quat(4); tran(3)] x tangent columns(6)
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
        for (int a = 0; a < 4; ++a) R[a][j] = Pv[3 * a] * c[j] + Pv[3 * a + 1] * c[6 + j] + Pv[3 * a + 2] * c[12 + j];
#pragma unroll
        for (int a = 0; a < 3; ++a) R[4 + a][j] = c[6 * (3 + a) + j];
    }
    const int64_t rq = S.off_viewq + 4 * (int64_t)v, rt = S.off_viewt + 3 * (int64_t)v;
    const int64_t cq = S.off_viewq + 4 * (int64_t)w, ct = S.off_viewt + 3 * (int64_t)w;
#pragma unroll
    for (int a = 0; a < 7; ++a) {
        double* row = cov + (a < 4 ? rq + a : rt + a - 4) * na;
#pragma unroll
        for (int b = 0; b < 4; ++b) row[cq + b] = R[a][0] * Pw[3 * b] + R[a][1] * Pw[3 * b + 1] + R[a][2] * Pw[3 * b + 2];
#pragma unroll
        for (int b = 0; b < 3; ++b) row[ct + b] = R[a][3 + b];
    }
}

}  // namespace calk
