// K1: fused project -> residual -> analytic Jacobian -> J^T J accumulation -> per-block
// Huber weight and chain rule -> per-tile sums, in one kernel (sm_100a, FP64).
//
// Replaces, per Levenberg-Marquardt iteration, what the reference does with
// ceres::AutoDiffCostFunction over src/estimation/residuals/{intrinsicresidual,
// extrinsicsresidual,bundleresidual}.h plus Ceres' loss correction and normal-equation
// formation (reached through solve_problem, src/estimation/detail/ceresutils.h:27-43).
//
// Work decomposition.  One lane owns one segment (in the fused mode: one residual block)
// and keeps the upper triangle of [J | r]^T [J | r] of that segment in registers.  The
// NE = 28 ... 190 entries do not fit one thread, so NROLE = 1 ... 3 warps ("roles") of the
// same CTA work on the same tile of 32 segments and each accumulates its own subset of the
// entries (K1Roles below).  The subsets are chosen so that every product the per-block
// epilogue needs is role-local: role 0 holds the twist-twist and twist-residual entries,
// and all six entries that couple the twist with one intrinsic column always sit in one role.
//
// Fused epilogue (one block per lane).  s_b = |r_b|^2 gives the Huber weight of the
// residual BLOCK (rho is applied to the whole block, so it is only known after
// the block's last corner); the role then forms w_b N_b for its entries, applies the 6x6
// chain rule T_b of the view-type pose (k1_math.cuh) from registers, and reduces over the
// 32 blocks of the tile through a padded shared-memory transpose in a fixed order (no
// floating-point atomics: results are run-to-run identical).  Per tile one row of NVT
// values is written; k_tile_reduce adds the rows per camera (and all-reduces them over NVLink).  For the
// kinds with per-view unknowns the per-block products (H_vv, g_v, E_vc, E_vi) are written
// per block for the Schur kernels instead of being reduced.
//
// Without the fused epilogue (small problems, whose blocks are cut into several segments
// to fill the machine) the roles store their entries to segN and the assembly kernels of
// refine_kernels.cu take over.
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "k1_kernel.cuh"
#include "k1_roles.hpp"
#include "refine_kernels.cuh"
#include "tile_stage.cuh"

namespace calk {

template <int MODEL, int IMODE, int VIEW>
static void launch_k1_v(const K1Args& P, cudaStream_t st) {
    using RT = K1Roles<MODEL, IMODE>;
    constexpr int threads = RT::NROLE * 32;
    constexpr int smem = K1Smem<MODEL, IMODE>::kBytes;
    static PerDeviceOnce once;
    if (once.first()) cudaFuncSetAttribute(k1_kernel<MODEL, IMODE, VIEW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k1_kernel<MODEL, IMODE, VIEW><<<(unsigned)P.L.n_tiles, threads, smem, st>>>(P);
}

template <int MODEL, int IMODE>
static void launch_k1_t(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, cudaStream_t st) {
    using RT = K1Roles<MODEL, IMODE>;
    K1Args P{L, B, S.huber_delta, 0};
    if (!L.fused) { launch_k1_v<MODEL, IMODE, NOT_FUSED>(P, st); return; }
    const bool reduce_rows = S.kind == 2 && S.view_free_global;
    P.nvt = RT::nvt(reduce_rows);
    if (S.kind != 2) launch_k1_v<MODEL, IMODE, VIEW_STORE>(P, st);
    else if (reduce_rows) launch_k1_v<MODEL, IMODE, VIEW_REDUCE>(P, st);
    else launch_k1_v<MODEL, IMODE, VIEW_NONE>(P, st);
}

#if defined(CALK_K1_ONLY_PINHOLE_NOSKEW)   // kernel experiments: compile the benchmark's instance only (never the shipped build)
#define CALK_K1_DISPATCH(FN, ...) FN<0, 1>(__VA_ARGS__)
#else
#define CALK_K1_DISPATCH(FN, ...)                                                          \
    do {                                                                                   \
        if (S.model == 0 && S.imode == 0) FN<0, 0>(__VA_ARGS__);                           \
        else if (S.model == 0 && S.imode == 1) FN<0, 1>(__VA_ARGS__);                      \
        else if (S.model == 0 && S.imode == 2) FN<0, 2>(__VA_ARGS__);                      \
        else if (S.model == 1 && S.imode == 0) FN<1, 0>(__VA_ARGS__);                      \
        else if (S.model == 1 && S.imode == 1) FN<1, 1>(__VA_ARGS__);                      \
        else FN<1, 2>(__VA_ARGS__);                                                        \
    } while (0)
#endif

void launch_k1(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, cudaStream_t st) {
    if (L.n_tiles == 0) return;
    CALK_K1_DISPATCH(launch_k1_t, S, L, B, st);
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
    CALK_K1_DISPATCH(roles_info_t, view_rows, n_roles, nvt, map);
}
int k1_num_passes(const ProblemShape& S) { int r = 1; k1_tile_value_map(S, &r, nullptr, nullptr); return r; }

// ---------------------------------------------------------------------------
// per-camera sums of the per-tile rows (fixed order, no atomics)
// ---------------------------------------------------------------------------
// (device code of k_tile_reduce: k1_kernel.cuh)
int launch_tile_reduce(const ProblemShape& S, const EvalBuffers& B, const ReduceDesc& R, int nvt, const calcomm::PeerArgs& peer, cudaStream_t st) {
    const TileReduceArgs A{B.tile_vals, nvt, R.tile_chunks, B.partial_tile, R.tile_cam_chunk_off, S.n_cams, B.tile_vmap, B.cam_sums, S.NV, R.tile_tickets, R.n_active_cams};
    k_tile_reduce<<<R.n_tile_chunks, 256, 0, st>>>(A, peer);
    return 1;
}

}  // namespace calk
