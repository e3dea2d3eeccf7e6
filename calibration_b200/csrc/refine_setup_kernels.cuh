// Device code of the layout and per-evaluation setup kernels (see refine_kernels.cu for the description).  Kept in a
// header so that tests/host_emul can run this very source on the CPU under a lock-step SIMT shim (test-only).
#pragma once
#include "refine_kernels.cuh"

namespace calk {

// layout: SoA observations of the caller -> tile-transposed device layout (refine_kernels.cuh).
// board_n > 0 (shared-board form of cal_problem_desc): sx / sy hold ONE board of board_n points and every block
// has board_n observations in board order, so observation i reads board point i % board_n; a segment never leaves
// its block, hence one modulo per thread.
__global__ void k_repack(DevLayout L, const double* __restrict__ sx, const double* __restrict__ sy,
                         const double* __restrict__ su, const double* __restrict__ sv,
                         const int64_t* __restrict__ seg_src, int board_n) {
    const int64_t tile = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const int64_t s = tile * 32 + lane;
    const int len = L.seg_len[s];
    const int depth = L.tile_depth[tile];
    const int64_t src = seg_src[s];
    const int64_t osrc = board_n > 0 ? src % board_n : src;   // where this segment's object points start
    double* dst = L.obs + L.tile_off[tile] * 128 + lane;
    for (int k = threadIdx.x >> 5; k < depth; k += blockDim.x >> 5) {
        const bool ok = k < len;
        dst[(int64_t)k * 128 + 0] = ok ? sx[osrc + k] : 0.0;
        dst[(int64_t)k * 128 + 32] = ok ? sy[osrc + k] : 0.0;
        dst[(int64_t)k * 128 + 64] = ok ? su[src + k] : 0.0;
        dst[(int64_t)k * 128 + 96] = ok ? sv[src + k] : 0.0;
    }
}

// bundle: robot poses b_se3_g, AoS [orig block][12] -> device-block-ordered SoA [12][n_blk]
__global__ void k_btg_permute(DevLayout L, const double* __restrict__ src) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.n_blk * 12) return;
    const int64_t b = i / 12; const int k = (int)(i % 12);
    const int64_t o = L.blk_orig[b];
    L.blk_bTg[(int64_t)k * L.n_blk + b] = o >= 0 ? src[o * 12 + k] : ((k % 4 == 0 && k < 9) ? 1.0 : 0.0);
}

// One launch per evaluation: threads c < n_cams additionally publish the camera constants and the
// per-camera chain-rule transform T_c that K1 / k_cost read; every thread derives the sensor rotation
// of ITS camera from the intrinsics itself (identity for the pinhole model), so no second launch and
// no dependency between the two parts.
__global__ void k_block_setup(ProblemShape S, DevLayout L, EvalBuffers B) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < S.n_cams) {
        const int c = (int)b;
        const double* intr = B.x + S.off_intr + (S.kind == 0 ? 0 : c * S.P);
        CamConst cc; cam_const_from_intr(intr, S.model, cc);
        B.camc[c] = cc;
        double T[36];
        for (int i = 0; i < 36; ++i) T[i] = 0.0;
        if (S.cam_pose_kind == 1) cam_transform_extrinsics(B.x + S.off_camt + 3 * c, cc.Rs, T);
        else if (S.cam_pose_kind == 2) cam_transform_bundle(B.x + S.off_camq + 4 * c, cc.Rs, T);
        for (int i = 0; i < 36; ++i) B.camT[c * 36 + i] = T[i];
    }
    if (b >= L.n_blk) return;
    double A[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double T[36];
    for (int i = 0; i < 36; ++i) T[i] = 0.0;
    if (L.blk_orig[b] >= 0) {
        const int cam = L.blk_cam[b];
        double Rs[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (S.model == 1) {  // rot_sensor of this block's camera (scheimpflug.h:150-153)
            const double* intr = B.x + S.off_intr + (S.kind == 0 ? 0 : cam * S.P);
            CamConst cc; cam_const_from_intr(intr, S.model, cc);
            for (int i = 0; i < 9; ++i) Rs[i] = cc.Rs[i];
        }
        BlockPose bp;
        block_pose(S, L, B.x, b, cam, true, bp);
        block_frame(bp, Rs, A);
        view_transform(bp, Rs, T);
    }
    // T = [[TL, 0], [BL, TL / 2]] (view_transform): the fused K1 reads only the 18 entries of TL and BL
    for (int i = 0; i < 36; ++i)
        if (!L.fused || i % 6 < 3) B.blk_Tv[(int64_t)i * L.n_blk + b] = T[i];
    for (int s = L.blk_seg_off[b]; s < L.blk_seg_off[b + 1]; ++s)
        for (int i = 0; i < 9; ++i) B.seg_frame[(int64_t)i * L.n_seg + s] = A[i];
}

}  // namespace calk
