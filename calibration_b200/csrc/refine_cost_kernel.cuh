// Device code of the residual-only pass (candidate steps of the LM; see refine_kernels.cu for the launcher).  Kept
// in a header so that tests/host_emul can run this very source on the CPU under the lock-step SIMT shim (test-only).
#pragma once
#include "k1_math.cuh"
#include "refine_kernels.cuh"
#include "tile_stage.cuh"

namespace calk {

// residual-only pass
template <int MODEL>
__global__ void __launch_bounds__(128) k_cost(DevLayout L, EvalBuffers B) {
    const int64_t tile = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (tile >= L.n_tiles) return;
    const int lane = threadIdx.x & 31;
    extern __shared__ __align__(128) unsigned char k1_smem[];
    const int64_t s = tile * 32 + lane;
    const int len = L.seg_len[s];
    const int depth = L.tile_depth[tile];
    TileStage ts; ts.init(k1_smem + (threadIdx.x >> 5) * kWarpStageBytes, L.obs + L.tile_off[tile] * 128, depth, lane);
    ts.issue(0, lane); ts.issue(1, lane);
    double A[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) A[i] = B.seg_frame[(int64_t)i * L.n_seg + s];
    const CamConst c = B.camc[L.seg_cam[s]];
    double acc = 0.0;
    for (int ch = 0; ch < ts.n_chunks; ++ch) {
        ts.wait(ch);
        const int k0 = ch * kChunk, kn = min(kChunk, depth - k0);
#pragma unroll 4
        for (int kk = 0; kk < kn; ++kk) {
            const double* q = ts.row(ch, kk, lane);
            const double X = q[0], Y = q[32], U = q[64], V = q[96];
            if (k0 + kk < len) acc += obs_ssr<MODEL>(c, A, X, Y, U, V);
        }
        __syncwarp();
        ts.issue(ch + 2, lane);
    }
    B.seg_ssr[s] = acc;
}

}  // namespace calk
