// TEST INFRASTRUCTURE — the launch layer of the product's refinement path for the CPU: every launch_* function that
// calibration_b200/csrc/refine_kernels.cu and k1_fused.cu implement with <<<...>>> is implemented here with
// simt::launch on the SAME kernel sources (same grids, same block sizes, same dispatch), so that the product's host
// code — refine_host.cu, compiled unmodified by g++ against tests/host_emul/fake_cuda/cuda_runtime.h — runs its whole
// path on the CPU: cal_refine_create (layout, repack), cal_refine_eval / cal_refine_cost, cal_refine_solve (the
// Levenberg–Marquardt loop, Schur steps, covariance), cal_refine_view_errors.  Linked together they form
// tests/host_emul/_build/libcalib_b200_simt.so, which exports the refinement part of the C ABI.
#define CALIB_SIMT_SHIM 1
#include "simt_shim.hpp"

#include <cuda_runtime.h>   // the fake one (include path order)

namespace calk {
__attribute__((aligned(128))) unsigned char k1_smem[232448];
}

#include "../../calibration_b200/csrc/k1_kernel.cuh"             // extern __shared__ users first ...
#include "../../calibration_b200/csrc/refine_cost_kernel.cuh"
#include "../../calibration_b200/csrc/refine_setup_kernels.cuh"
#undef __shared__
#define __shared__ static                                          // ... then the kernels with static shared arrays
#include "../../calibration_b200/csrc/refine_schur_kernels.cuh"
#include "../../calibration_b200/csrc/refine_assemble_kernels.cuh"
#include "../../calibration_b200/csrc/comm.h"

namespace calk {

#define SIMT_DISPATCH(FN, ...)                                             \
    do {                                                                   \
        if (S.model == 0 && S.imode == 0) FN<0, 0>(__VA_ARGS__);           \
        else if (S.model == 0 && S.imode == 1) FN<0, 1>(__VA_ARGS__);      \
        else if (S.model == 0 && S.imode == 2) FN<0, 2>(__VA_ARGS__);      \
        else if (S.model == 1 && S.imode == 0) FN<1, 0>(__VA_ARGS__);      \
        else if (S.model == 1 && S.imode == 1) FN<1, 1>(__VA_ARGS__);      \
        else FN<1, 2>(__VA_ARGS__);                                        \
    } while (0)

// ---- refine_kernels.cu: layout, setup, residual-only pass ----
void launch_repack(const DevLayout& L, const double* sx, const double* sy, const double* su, const double* sv, const int64_t* seg_src, int board_n,
                   cudaStream_t) {
    if (L.n_tiles == 0) return;
    simt::launch((unsigned)L.n_tiles, 128, [&] { k_repack(L, sx, sy, su, sv, seg_src, board_n); });
}
void launch_btg_permute(const DevLayout& L, const double* src, cudaStream_t) {
    const int64_t n = L.n_blk * 12;
    if (n > 0) simt::launch((unsigned)((n + 255) / 256), 256, [&] { k_btg_permute(L, src); });
}
void launch_setup(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, cudaStream_t) {
    const int64_t n = L.n_blk > S.n_cams ? L.n_blk : S.n_cams;
    simt::launch((unsigned)((n + 127) / 128), 128, [&] { k_block_setup(S, L, B); });
}
void launch_cost(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, cudaStream_t) {
    if (L.n_tiles == 0) return;
    const unsigned g = (unsigned)((L.n_tiles + 3) / 4);
    if (4 * kWarpStageBytes > (int)sizeof k1_smem) std::abort();
    if (S.model == 0) simt::launch(g, 128, [&] { k_cost<0>(L, B); });
    else simt::launch(g, 128, [&] { k_cost<1>(L, B); });
}

// ---- k1_fused.cu ----
template <int MODEL, int IMODE>
static void launch_k1_t(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B) {
    using RT = K1Roles<MODEL, IMODE>;
    if (K1Smem<MODEL, IMODE>::kBytes > (int)sizeof k1_smem) std::abort();
    K1Args P{L, B, S.huber_delta, 0};
    const unsigned g = (unsigned)L.n_tiles, t = RT::NROLE * 32;
    if (!L.fused) { simt::launch(g, t, [&] { k1_kernel<MODEL, IMODE, NOT_FUSED>(P); }); return; }
    const bool reduce_rows = S.kind == 2 && S.view_free_global;
    P.nvt = RT::nvt(reduce_rows);
    if (S.kind != 2) simt::launch(g, t, [&] { k1_kernel<MODEL, IMODE, VIEW_STORE>(P); });
    else if (reduce_rows) simt::launch(g, t, [&] { k1_kernel<MODEL, IMODE, VIEW_REDUCE>(P); });
    else simt::launch(g, t, [&] { k1_kernel<MODEL, IMODE, VIEW_NONE>(P); });
}
void launch_k1(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, cudaStream_t) {
    if (L.n_tiles == 0) return;
    SIMT_DISPATCH(launch_k1_t, S, L, B);
}
template <int MODEL, int IMODE>
static void roles_info_t(bool view_rows, int* n_roles, int* nvt, std::vector<int32_t>* map) {
    using RT = K1Roles<MODEL, IMODE>;
    if (n_roles) *n_roles = RT::NROLE;
    if (nvt) *nvt = RT::nvt(view_rows);
    if (map) RT::value_map(view_rows, *map);
}
void k1_tile_value_map(const ProblemShape& S, int* n_roles, int* nvt, std::vector<int32_t>* map) {
    const bool view_rows = S.kind == 2 && S.view_free_global;
    SIMT_DISPATCH(roles_info_t, view_rows, n_roles, nvt, map);
}
int k1_num_passes(const ProblemShape& S) { int r = 1; k1_tile_value_map(S, &r, nullptr, nullptr); return r; }
int launch_tile_reduce(const ProblemShape& S, const EvalBuffers& B, const ReduceDesc& R, int nvt, const calcomm::PeerArgs& peer, cudaStream_t) {
    const TileReduceArgs A{B.tile_vals, nvt, R.tile_chunks, B.partial_tile, R.tile_cam_chunk_off, S.n_cams, B.tile_vmap, B.cam_sums, S.NV, R.tile_tickets, R.n_active_cams};
    simt::launch((unsigned)R.n_tile_chunks, 256, [&] { k_tile_reduce(A, peer); });
    return 1;
}
float dfma_peak_ms(double*, int, int, int, cudaStream_t) { return 1.0f; }   // a benchmark utility: no meaning on the CPU

// ---- refine_kernels.cu: launch_assemble ----
template <int MODEL, int IMODE>
static void launch_view_part_t(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B) {
    const unsigned g = (unsigned)((L.n_blk + 127) / 128);
    if (L.one_seg_per_blk) simt::launch(g, 128, [&] { k_view_part<MODEL, IMODE, true>(S, L, B); });
    else simt::launch(g, 128, [&] { k_view_part<MODEL, IMODE, false>(S, L, B); });
}
int launch_assemble(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, const ReduceDesc& R, int jac, const calcomm::PeerArgs*, bool* peer_done,
                    cudaStream_t st) {
    if (peer_done) *peer_done = false;
    if (jac && L.fused) return launch_tile_reduce(S, B, R, R.nvt, calcomm::PeerArgs{}, st);
    int launches = 0;
    const unsigned gb = (unsigned)((L.n_blk + 127) / 128);
    const int rr_row = S.NL * S.NC - S.NC * (S.NC - 1) / 2;
    if (jac) simt::launch(gb, 128, [&] { k_block_weight<1>(S, L, B, rr_row); }); else simt::launch(gb, 128, [&] { k_block_weight<0>(S, L, B, rr_row); });
    ++launches;
    const int NV = jac ? S.NV : 1;
    const bool view_part = jac && (S.kind == 2 ? S.view_free_global != 0 : S.n_views > 0);
    if (view_part) { SIMT_DISPATCH(launch_view_part_t, S, L, B); ++launches; }
    if (jac) {
        simt::launch((unsigned)R.n_seg_chunks, 256, [&] { k_colsum<true>(B.segN, L.n_seg, B.seg_w, R.seg_chunks, B.partial, S.NE); }, (unsigned)S.NE);
        simt::launch((unsigned)((S.n_cams * S.NE + 127) / 128), 128, [&] { k_final_reduce(B.partial, R.seg_cam_chunk_off, S.n_cams, S.NE, B.cam_sums, NV, 0); });
        launches += 2;
    }
    const int n_brows = jac ? NV - S.NE : 1;
    simt::launch((unsigned)R.n_blk_chunks, 256, [&] { k_colsum<false>(B.blk_rows, L.n_blk, nullptr, R.blk_chunks, B.partial_blk, n_brows); }, (unsigned)n_brows);
    simt::launch((unsigned)((S.n_cams * n_brows + 127) / 128), 128,
                 [&] { k_final_reduce(B.partial_blk, R.blk_cam_chunk_off, S.n_cams, n_brows, B.cam_sums, NV, jac ? S.NE : 0); });
    return launches + 2;
}

// ---- refine_kernels.cu: K2 and covariance ----
void launch_view_gather(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, const ViewBuffers& V, cudaStream_t) {
    if (S.n_views == 0) return;
    simt::launch((unsigned)((S.n_views + 127) / 128), 128, [&] { k_view_gather(S, L, B, V); });
}
void launch_view_scale(const ProblemShape& S, const ViewBuffers& V, int compute_scale, cudaStream_t) {
    if (S.n_views == 0) return;
    simt::launch((unsigned)((S.n_views * 6 + 127) / 128), 128, [&] { k_view_scale(S, V, compute_scale); });
}
void launch_backsub(const ProblemShape& S, const DevLayout& L, const ViewBuffers& V, int ns, cudaStream_t) {
    if (S.n_views == 0) return;
    simt::launch((unsigned)((S.n_views + kBacksubViews - 1) / kBacksubViews), 256, [&] { k_backsub(S, L, V, ns); });
}
void launch_view_plus(const ProblemShape& S, const EvalBuffers& B, const ViewBuffers& V, double t, cudaStream_t) {
    if (S.n_views == 0) return;
    simt::launch((unsigned)((S.n_views + 127) / 128), 128, [&] { k_view_plus(S, B, V, t); });
}
void launch_view_norms(const ProblemShape& S, const EvalBuffers& B, const ViewBuffers& V, cudaStream_t) {
    if (S.n_views == 0) return;
    simt::launch((unsigned)((S.n_views + 127) / 128), 128, [&] { k_view_norms(S, B, V); });
}
void launch_reduce_views(const ViewBuffers& V, int n_views, cudaStream_t) { simt::launch(n_views >= 2048 ? 2u : 1u, 256, [&] { k_reduce_views(V, n_views); }); }
int schur_num_ctas(int n_views) { return n_views < 64 ? 1 : (n_views < 444 * 16 ? (n_views + 15) / 16 : 444); }
void launch_schur(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, const ViewBuffers& V, int ns, double radius, cudaStream_t) {
    if (S.n_views == 0) return;
    simt::launch((unsigned)((S.n_views + 127) / 128), 128, [&] { k_view_chol(S, V, 1.0 / radius); });
    simt::launch((unsigned)(((int64_t)S.n_views * 32 + kFactorThreads - 1) / kFactorThreads), kFactorThreads, [&] { k_schur_factor(S, L, B, V); });
    const int n_cta = schur_num_ctas(S.n_views);
    const int per = (S.n_views + n_cta - 1) / n_cta;
    const int nt = (ns + 1 + kSyrkTile - 1) / kSyrkTile;
    const int threads = (nt * (nt + 1) / 2 + 31) / 32 * 32;
    { int bt = 0, warps = 0; syrk_shape(ns, &bt, &warps); simt::launch((unsigned)n_cta, (unsigned)(32 * warps), [&] { k_schur_syrk_any(bt, S, V, ns, per); }); }
    const int na = ns + 1;
    simt::launch((unsigned)((na * na + kSchurReduceEntries - 1) / kSchurReduceEntries), 4 * kSchurReduceEntries, [&] { k_schur_reduce(V, n_cta, ns); });
}
bool launch_reduced_solve(const double* Sm, const double* gss, const ViewBuffers& V, int ns, int32_t* info, cudaStream_t) {
    if (ns > kReducedMaxN) return false;
    simt::launch(1, 256, [&] { k_reduced_solve(Sm, gss, V, ns, info); });
    return true;
}
void launch_cov_views(const ProblemShape& S, const DevLayout& L, const ViewBuffers& V, const double* x, int ns, const double* W, double* Z, double* G,
                      double* Ainv, double* cov, int64_t na, cudaStream_t) {
    if (S.n_views == 0) return;
    simt::launch((unsigned)S.n_views, 64, [&] { k_cov_view_prep(S, L, V, ns, W, Z, G, Ainv); });
    const unsigned t = (unsigned)((S.n_views + kCovTile - 1) / kCovTile);
    simt::launch(t, kCovTile * kCovTile, [&] { k_cov_vv(S, V, x, ns, Z, G, Ainv, cov, na); }, t);
}

}  // namespace calk

// ---- comm.cpp / comm_peer.cu: the multi-GPU exchange needs devices; a single process never attaches a communicator ----
namespace calcomm {
bool Comm::unique_id(uint8_t*, std::string* err) { if (err) *err = "no communicator on the CPU"; return false; }
Comm* Comm::create(const uint8_t*, int, int, std::string* err) { if (err) *err = "no communicator on the CPU"; return nullptr; }
Comm::~Comm() = default;
bool Comm::allreduce_sum(double*, size_t, cudaStream_t) { return false; }
bool Comm::allreduce_host(double*, size_t, bool) { return false; }
bool Comm::peer_export(uint8_t*) { return false; }
bool Comm::peer_enable(const uint8_t*) { return false; }
bool Comm::allreduce_test(double*, size_t, bool) { return false; }
}  // namespace calcomm
