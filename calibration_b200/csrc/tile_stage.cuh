// TMA staging of a warp's observation tile (shared by K1 and the residual-only pass).
// Under CALIB_SIMT_SHIM (tests/host_emul, test-only) the bulk copy is a memcpy by the issuing lane and the mbarrier a
// completion counter; under nvcc those branches do not exist.
#pragma once
#include <stdint.h>

namespace calk {

// ---------------------------------------------------------------------------
// TMA staging of a warp's tile: bulk asynchronous copies (cp.async.bulk, SASS UBLKCP) of
// kChunk k-slices (kChunk x 1 KB, contiguous in the tile-transposed layout) into a per-warp
// two-stage shared-memory ring, completion signalled on an mbarrier.  One elected lane
// issues; all lanes then read their column with conflict-free 8-byte shared loads.
// ---------------------------------------------------------------------------
constexpr int kChunk = 8;                       // k-slices per stage (8 KB)
constexpr int kStageDoubles = kChunk * 128;
constexpr int kWarpStageBytes = 2 * kStageDoubles * 8 + 16;  // two stages + two mbarriers

#if !defined(CALIB_SIMT_SHIM)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
#endif

struct TileStage {
    double* buf;                // [2][kChunk][4][32]
    unsigned long long* bar;    // [2]
    const double* src;          // tile base in global memory
    int depth, n_chunks;

    __device__ __forceinline__ void init(unsigned char* warp_smem, const double* tile_src, int tile_depth, int lane) {
        buf = reinterpret_cast<double*>(warp_smem);
        bar = reinterpret_cast<unsigned long long*>(warp_smem + 2 * kStageDoubles * 8);
        src = tile_src; depth = tile_depth; n_chunks = (tile_depth + kChunk - 1) / kChunk;
        if (lane == 0) {
#if defined(CALIB_SIMT_SHIM)
            simt::mbar_init(&bar[0]); simt::mbar_init(&bar[1]);
#else
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
        }
        __syncwarp();
    }
    // issue chunk c into stage c & 1 (lane 0 only)
    __device__ __forceinline__ void issue(int c, int lane) {
        if (lane == 0 && c < n_chunks) {
            const int ks = min(kChunk, depth - c * kChunk);
            const unsigned bytes = (unsigned)ks * 1024u;
#if defined(CALIB_SIMT_SHIM)
            simt::bulk_copy_and_complete(buf + (c & 1) * kStageDoubles, src + (int64_t)c * kStageDoubles, bytes, &bar[c & 1]);
#else
            const unsigned mb = smem_u32(&bar[c & 1]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(buf + (c & 1) * kStageDoubles)), "l"(src + (int64_t)c * kStageDoubles), "r"(bytes), "r"(mb)
                         : "memory");
#endif
        }
    }
    // wait until chunk c has landed
    __device__ __forceinline__ void wait(int c) {
#if defined(CALIB_SIMT_SHIM)
        simt::mbar_wait(&bar[c & 1], (c >> 1) + 1);   // the (c / 2 + 1)-th completion of this stage
#else
        const unsigned mb = smem_u32(&bar[c & 1]);
        const unsigned parity = (unsigned)(c >> 1) & 1u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@!p bra WAIT_%=;\n\t}" ::"r"(mb), "r"(parity) : "memory");
#endif
    }
    __device__ __forceinline__ const double* row(int c, int kk, int lane) const { return buf + (c & 1) * kStageDoubles + kk * 128 + lane; }
};


}  // namespace calk
