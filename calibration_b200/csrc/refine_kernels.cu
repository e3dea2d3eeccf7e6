// CUDA kernels of the refinement hot path (sm_100a).  FP64 throughout: the
// reference solves in double with Ceres and the north star asks for parity at
// 1e-8 relative, so tensor cores are deliberately not used (nothing here is a
// large dense contraction; see DESIGN.md §5 for the roofline argument).
//
//   k_repack        one-time gather of the uploaded SoA observations into the
//                   tile-transposed layout (refine_kernels.cuh)
//   k_block_setup
//                   per evaluation: camera constants, composite block poses
//                   (pose chains of src/estimation/residuals/*.h), sensor-frame
//                   frames and the 6x6 chain-rule transforms
//   (K1 itself, the fused residual + Jacobian + J^T J + per-block epilogue kernel, is k1_fused.cu)
//   k_cost          residual-only pass
//   k_assemble      per block: Huber weight (per residual block),
//                   chain rule, deterministic warp transpose-reductions into
//                   per-camera sums (no floating-point atomics)
//   k_view_* / k_schur_* / k_backsub
//                   K2: batched 6x6 Cholesky of the damped per-view pose blocks,
//                   Schur complement onto the shared block as a tiled rank-6
//                   SYRK, back-substitution
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "refine_kernels.cuh"
#include "refine_setup_kernels.cuh"
#include "refine_schur_kernels.cuh"
#include "refine_assemble_kernels.cuh"
#include "tile_stage.cuh"
#include "refine_cost_kernel.cuh"

namespace calk {

// ---------------------------------------------------------------------------
// layout
// ---------------------------------------------------------------------------
void launch_repack(const DevLayout& L, const double* sx, const double* sy, const double* su, const double* sv,
                   const int64_t* seg_src, int board_n, cudaStream_t st) {
    if (L.n_tiles == 0) return;
    k_repack<<<(unsigned)L.n_tiles, 128, 0, st>>>(L, sx, sy, su, sv, seg_src, board_n);
}

void launch_btg_permute(const DevLayout& L, const double* src, cudaStream_t st) {
    const int64_t n = L.n_blk * 12;
    if (n > 0) k_btg_permute<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(L, src);
}

// ---------------------------------------------------------------------------
// per-evaluation setup
// ---------------------------------------------------------------------------
void launch_setup(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, cudaStream_t st) {
    const int64_t n = L.n_blk > S.n_cams ? L.n_blk : S.n_cams;
    k_block_setup<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(S, L, B);
}

#define CALK_DISPATCH(FN, ...)                                                             \
    do {                                                                                   \
        if (S.model == 0 && S.imode == 0) FN<0, 0>(__VA_ARGS__);                           \
        else if (S.model == 0 && S.imode == 1) FN<0, 1>(__VA_ARGS__);                      \
        else if (S.model == 0 && S.imode == 2) FN<0, 2>(__VA_ARGS__);                      \
        else if (S.model == 1 && S.imode == 0) FN<1, 0>(__VA_ARGS__);                      \
        else if (S.model == 1 && S.imode == 1) FN<1, 1>(__VA_ARGS__);                      \
        else FN<1, 2>(__VA_ARGS__);                                                        \
    } while (0)

// residual-only pass (device code: refine_cost_kernel.cuh)
void launch_cost(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, cudaStream_t st) {
    if (L.n_tiles == 0) return;
    const unsigned g = (unsigned)((L.n_tiles + 3) / 4);
    constexpr int smem = 4 * kWarpStageBytes;
    static PerDeviceOnce once;
    if (once.first()) {
        cudaFuncSetAttribute(k_cost<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k_cost<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    if (S.model == 0) k_cost<0><<<g, 128, smem, st>>>(L, B);
    else k_cost<1><<<g, 128, smem, st>>>(L, B);
}

// ---------------------------------------------------------------------------
// per-block assembly + deterministic per-camera reduction
// ---------------------------------------------------------------------------
// k_block_weight   per block: s = sum of squared residuals (sum over its
//                  segments), Huber rho / weight (per residual BLOCK),
//                  cost row, and the weight copied to each of its segments
// k_view_part      per block: chain rule of the view-type pose block,
//                  Q = T^T N_xixi, H_vv = Q T, g_v = T^T N_xir, E_vi = T^T N_xii
// k_colsum         per-camera weighted column sums of the [row][column]
//                  matrices (segment- or block-indexed) in fixed order: strided
//                  per-thread sums, shuffle tree, shared-memory tree — no
//                  floating-point atomics, so results are run-to-run identical
// (device code: refine_assemble_kernels.cuh)
template <int MODEL, int IMODE>
static void launch_view_part_t(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, cudaStream_t st) {
    const unsigned g = (unsigned)((L.n_blk + 127) / 128);
    if (L.one_seg_per_blk) k_view_part<MODEL, IMODE, true><<<g, 128, 0, st>>>(S, L, B);
    else k_view_part<MODEL, IMODE, false><<<g, 128, 0, st>>>(S, L, B);
}

int launch_assemble(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, const ReduceDesc& R, int jac,
                    const calcomm::PeerArgs* peer, bool* peer_done, cudaStream_t st) {
    if (peer_done) *peer_done = false;
    if (jac && L.fused) {   // K1 already weighted, transformed and summed per tile
        if (peer_done) *peer_done = peer != nullptr;
        return launch_tile_reduce(S, B, R, R.nvt, peer ? *peer : calcomm::PeerArgs{}, st);
    }
    int launches = 0;
    const unsigned gb = (unsigned)((L.n_blk + 127) / 128);
    const int rr_row = S.NL * S.NC - S.NC * (S.NC - 1) / 2;  // idx(NC, NC)
    if (jac) k_block_weight<1><<<gb, 128, 0, st>>>(S, L, B, rr_row); else k_block_weight<0><<<gb, 128, 0, st>>>(S, L, B, rr_row);
    ++launches;
    const int NV = jac ? S.NV : 1;
    const bool view_part = jac && (S.kind == 2 ? S.view_free_global != 0 : S.n_views > 0);
    if (view_part) { CALK_DISPATCH(launch_view_part_t, S, L, B, st); ++launches; }
    if (jac) {
        // segment-indexed rows: the local systems, weighted per segment
        k_colsum<true><<<dim3(R.n_seg_chunks, S.NE), 256, 0, st>>>(B.segN, L.n_seg, B.seg_w, R.seg_chunks, B.partial, S.NE);
        k_final_reduce<<<(S.n_cams * S.NE + 127) / 128, 128, 0, st>>>(B.partial, R.seg_cam_chunk_off, S.n_cams, S.NE, B.cam_sums, NV, 0);
        launches += 2;
    }
    // block-indexed rows: cost (+ the view-type part of the bundle kind)
    const int n_brows = jac ? NV - S.NE : 1;
    k_colsum<false><<<dim3(R.n_blk_chunks, n_brows), 256, 0, st>>>(B.blk_rows, L.n_blk, nullptr, R.blk_chunks, B.partial_blk, n_brows);
    k_final_reduce<<<(S.n_cams * n_brows + 127) / 128, 128, 0, st>>>(B.partial_blk, R.blk_cam_chunk_off, S.n_cams, n_brows, B.cam_sums, NV, jac ? S.NE : 0);
    launches += 2;
    return launches;
}

// ---------------------------------------------------------------------------
// K2: per-view pose blocks — launchers (device code: refine_schur_kernels.cuh)
// ---------------------------------------------------------------------------
void launch_view_gather(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, const ViewBuffers& V,
                        cudaStream_t st) {
    if (S.n_views == 0) return;
    k_view_gather<<<(S.n_views + 127) / 128, 128, 0, st>>>(S, L, B, V);
}
void launch_view_scale(const ProblemShape& S, const ViewBuffers& V, int compute_scale, cudaStream_t st) {
    if (S.n_views == 0) return;
    k_view_scale<<<(S.n_views * 6 + 127) / 128, 128, 0, st>>>(S, V, compute_scale);
}
void launch_backsub(const ProblemShape& S, const DevLayout& L, const ViewBuffers& V, int ns, cudaStream_t st) {
    if (S.n_views == 0) return;
    k_backsub<<<(unsigned)((S.n_views + kBacksubViews - 1) / kBacksubViews), 256, 0, st>>>(S, L, V, ns);   // 32 views per CTA
}
void launch_view_plus(const ProblemShape& S, const EvalBuffers& B, const ViewBuffers& V, double t, cudaStream_t st) {
    if (S.n_views == 0) return;
    k_view_plus<<<(S.n_views + 127) / 128, 128, 0, st>>>(S, B, V, t);
}
void launch_view_norms(const ProblemShape& S, const EvalBuffers& B, const ViewBuffers& V, cudaStream_t st) {
    if (S.n_views == 0) return;
    k_view_norms<<<(S.n_views + 127) / 128, 128, 0, st>>>(S, B, V);
}
void launch_reduce_views(const ViewBuffers& V, int n_views, cudaStream_t st) {
    const int ctas = n_views >= 64 * 256 ? kReduceViewsCtas : (n_views + 1023) / 1024 > 0 ? (n_views + 1023) / 1024 : 1;   // >= 1024 views per CTA
    k_reduce_views<<<ctas < kReduceViewsCtas ? ctas : kReduceViewsCtas, 256, 0, st>>>(V, n_views);
}

// up to three CTAs per SM (4-warp CTAs at the common shared-block widths), at least 16 views each
int schur_num_ctas(int n_views) { return n_views < 64 ? 1 : (n_views < kSchurMaxCtas * 16 ? (n_views + 15) / 16 : kSchurMaxCtas); }   // an upper bound (partialC is sized by it)

void launch_schur(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, const ViewBuffers& V, int ns,
                  double radius, cudaStream_t st) {
    if (S.n_views == 0) return;
    k_view_chol<<<(unsigned)((S.n_views + 127) / 128), 128, 0, st>>>(S, V, 1.0 / radius);                                                             // one thread per view
    k_schur_factor<<<(unsigned)(((int64_t)S.n_views * 32 + kFactorThreads - 1) / kFactorThreads), kFactorThreads, 0, st>>>(S, L, B, V);              // one warp per view
    int bt = 0, warps = 0;
    syrk_shape(ns, &bt, &warps);
    const int threads = 32 * warps;                                           // one warp per block of bt x bt tiles of the upper triangle
    const int smem = 2 * kSyrkViews * 6 * V.ncp * (int)sizeof(double) + 16;   // two stages of dense rows + two mbarriers
    const int smem_max = 2 * kSyrkViews * 6 * kSyrkMaxN * (int)sizeof(double) + 16;
    void (*kern)(ProblemShape, ViewBuffers, int, int) = bt == 3 ? k_schur_syrk<3, 320> : (threads <= 320 ? k_schur_syrk<5, 320> : k_schur_syrk<5, 480>);
    static PerDeviceOnce once;
    if (once.first()) {
        cudaFuncSetAttribute(k_schur_syrk<3, 320>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
        cudaFuncSetAttribute(k_schur_syrk<5, 320>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
        cudaFuncSetAttribute(k_schur_syrk<5, 480>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    }
    int n_cta = schur_num_ctas(S.n_views);
    if (n_cta == kSchurMaxCtas) {   // a large problem: exactly one wave of resident CTAs (444 CTAs on 296 slots ran as one and a half)
        int occ = 0, dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) == cudaSuccess && occ > 0)
            n_cta = std::min(n_cta, occ * sms);
    }
    const int per = (S.n_views + n_cta - 1) / n_cta;
    kern<<<n_cta, threads, smem, st>>>(S, V, ns, per);
    const int na = ns + 1;
    k_schur_reduce<<<(na * na + kSchurReduceEntries - 1) / kSchurReduceEntries, 4 * kSchurReduceEntries, 0, st>>>(V, n_cta, ns);
}

// reduced system of one LM iteration on the device (k_reduced_solve); false: too wide for one CTA's shared memory
bool launch_reduced_solve(const double* Sm, const double* gss, const ViewBuffers& V, int ns, int32_t* info, cudaStream_t st) {
    if (ns > kReducedMaxN) return false;
    const int smem = (int)(reduced_solve_smem_doubles(ns) * sizeof(double));
    static PerDeviceOnce once;
    if (once.first()) cudaFuncSetAttribute(k_reduced_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(reduced_solve_smem_doubles(kReducedMaxN) * sizeof(double)));
    k_reduced_solve<<<1, 256, smem, st>>>(Sm, gss, V, ns, info);
    return true;
}

// ---------------------------------------------------------------------------
// Covariance of the kinds with per-view pose blocks (compute_covariance, ceresutils.h:69-126;
// SURVEY B.5), block-structured.  With the (Jacobi-scaled, undamped) normal matrix
//     H = [[A, E], [E^T, Hss]],  A = blockdiag(A_v) (6x6 per view),
// the inverse is  C_ss = W = (Hss - sum_v E_v^T A_v^-1 E_v)^-1,  C_vs = -Z_v W,
// C_vw = delta_vw A_v^-1 + Z_v W Z_w^T  with  Z_v = A_v^-1 E_v = L_v^-T (L_v^-1 E_v).
// launch_schur with an infinite radius leaves L_v and F_b = L_v^-1 E_b on the device and the
// Schur complement for the host, which inverts the small shared block; k_cov_view_prep forms
// Z_v, G_v = Z_v W and A_v^-1 per view, k_cov_vv the n_views^2 view-view blocks, lifted to
// ambient coordinates with the quaternion plus-Jacobians and un-scaled, straight into the dense
// covariance in the reference's block order.
// ---------------------------------------------------------------------------
// (device code of k_cov_view_prep / k_cov_vv: refine_schur_kernels.cuh)
void launch_cov_views(const ProblemShape& S, const DevLayout& L, const ViewBuffers& V, const double* x, int ns, const double* W,
                      double* Z, double* G, double* Ainv, double* cov, int64_t na, cudaStream_t st) {
    if (S.n_views == 0) return;
    k_cov_view_prep<<<S.n_views, 64, 0, st>>>(S, L, V, ns, W, Z, G, Ainv);
    const unsigned t = (unsigned)((S.n_views + kCovTile - 1) / kCovTile);
    k_cov_vv<<<dim3(t, t), kCovTile * kCovTile, 0, st>>>(S, V, x, ns, Z, G, Ainv, cov, na);
}

// FP64 FMA-chain microbenchmark: the sustained DFMA rate of this GPU under its
// actual clocks, used as the denominator of the FP64 roofline fraction.
__global__ void k_dfma_peak(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
float dfma_peak_ms(double* scratch, int blocks, int threads, int iters, cudaStream_t st) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_dfma_peak<<<blocks, threads, 0, st>>>(scratch, iters / 8);
    cudaEventRecord(e0, st);
    k_dfma_peak<<<blocks, threads, 0, st>>>(scratch, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
}

}  // namespace calk
