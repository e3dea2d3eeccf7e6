// Device side of the NVLink peer-memory all-reduce (comm_peer.cu), as a function one CTA calls: the stand-alone kernel
// k_peer_allreduce and the last CTA of K1's per-camera reduction (k_tile_reduce, k1_kernel.cuh) — there the
// all-reduce is fused with the reduction that feeds it, one launch per pass instead of three.
#pragma once
#include <stdint.h>

namespace calcomm {

constexpr int kPeerMaxDoubles = 4096;   // largest block the peer-memory all-reduce takes (32 KB)

// what a kernel needs to take part in all-reduce number `epoch` of a communicator (Comm::peer_args)
struct PeerArgs {
    double* const* peers = nullptr;     // device array [world] of the ranks' receive regions (own region included)
    int rank = 0, world = 1;            // world == 1: no exchange
    unsigned long long epoch = 0;
    int* timed_out = nullptr;           // host-mapped flag raised by the bounded wait
};

#if defined(__CUDACC__)
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// region of one rank: slots [2][world][kPeerMaxDoubles] doubles, then flags [2][world] u64.
// All threads of ONE CTA call this with the same arguments; buf[0, n) is replaced by the sum over the ranks, added in
// rank order (bitwise identical on every rank).  `bad` is a shared-memory int of the caller.
__device__ __forceinline__ void peer_allreduce_cta(double* __restrict__ buf, int n, const PeerArgs& a, int* bad) {
    const int p = (int)(a.epoch & 1ULL), rank = a.rank, world = a.world;
    if (threadIdx.x == 0) *bad = 0;
    // 1. publish my block to every rank (my own region included); the block may have been written by other CTAs
    for (int r = 0; r < world; ++r) {
        double* dst = a.peers[r] + ((size_t)p * world + rank) * kPeerMaxDoubles;
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldcg(buf + i);
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) {
        unsigned long long* flag = reinterpret_cast<unsigned long long*>(a.peers[threadIdx.x] + (size_t)2 * world * kPeerMaxDoubles) + p * world + rank;
        st_release_sys(flag, a.epoch);
        // 2. wait for source threadIdx.x in my own region
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(a.peers[rank] + (size_t)2 * world * kPeerMaxDoubles) + p * world + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(mine) < a.epoch) {
            if (clock64() - t0 > 4000000000LL) { *bad = 1; break; }  // ~2 s at 1.9 GHz
        }
    }
    __syncthreads();
    if (*bad) {   // a peer never arrived: poison the block and tell the host, which fails the call with CAL_ERR_COMM
        if (threadIdx.x == 0) { *a.timed_out = 1; __threadfence_system(); }
        for (int i = threadIdx.x; i < n; i += blockDim.x) buf[i] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    // 3. rank-ordered sum of the world slots (L1 bypassed: the slots were written by other GPUs)
    const double* base = a.peers[rank] + (size_t)p * world * kPeerMaxDoubles;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += __ldcg(base + (size_t)r * kPeerMaxDoubles + i);
        buf[i] = s;
    }
}
#endif

}  // namespace calcomm
