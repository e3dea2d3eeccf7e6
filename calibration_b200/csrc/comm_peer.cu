// NVLink peer-memory all-reduce for the small per-pass blocks (one FP64 sum-allreduce of
// [cost | g | upper(H)] per evaluation, <= 32 KB, latency-bound).  One process per GPU; every rank owns a
// receive region that its peers map through CUDA IPC.  ONE kernel per all-reduce:
//   1. every rank stores its block into its slot of EVERY rank's receive region (remote stores over
//      NVLink / NVSwitch), fences at system scope and raises a per-source flag carrying the epoch;
//   2. waits until all world flags of its own region carry the epoch;
//   3. adds the world slots in RANK ORDER — the sum is bitwise identical on every rank, which the
//      replicated host LM relies on.
// Slots and flags are double-buffered by epoch parity (a rank can be at most one epoch ahead of a peer:
// it cannot finish epoch e + 1 without that peer's e + 1 flag, which the peer raises only after it has
// consumed epoch e).  The wait is bounded: after ~2 s the kernel poisons the result with NaN instead of
// hanging the GPU, and the host falls back to NCCL for the rest of the run.
#include <stdint.h>

#include "comm.h"

namespace calcomm {

namespace {
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
}  // namespace

// region of one rank: slots [2][world][kPeerMaxDoubles] doubles, then flags [2][world] u64
__global__ void __launch_bounds__(256) k_peer_allreduce(double* __restrict__ buf, int n, double* const* __restrict__ peers, int rank,
                                                        int world, unsigned long long epoch, int* __restrict__ timed_out) {
    __shared__ int bad;
    const int p = (int)(epoch & 1ULL);
    if (threadIdx.x == 0) bad = 0;
    // 1. publish my block to every rank (my own region included)
    for (int r = 0; r < world; ++r) {
        double* dst = peers[r] + ((size_t)p * world + rank) * kPeerMaxDoubles;
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = buf[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        unsigned long long* flag = reinterpret_cast<unsigned long long*>(peers[threadIdx.x] + (size_t)2 * world * kPeerMaxDoubles) + p * world + rank;
        st_release_sys(flag, epoch);
        // 2. wait for source threadIdx.x in my own region
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(peers[rank] + (size_t)2 * world * kPeerMaxDoubles) + p * world + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(mine) < epoch) {
            if (clock64() - t0 > 4000000000LL) { bad = 1; break; }  // ~2 s at 1.9 GHz
        }
    }
    __syncthreads();
    if (bad) {
        if (threadIdx.x == 0) *timed_out = 1;
        for (int i = threadIdx.x; i < n; i += blockDim.x) buf[i] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    // 3. rank-ordered sum of the world slots (L1 bypassed: the slots were written by other GPUs)
    const double* base = peers[rank] + (size_t)p * world * kPeerMaxDoubles;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += __ldcg(base + (size_t)r * kPeerMaxDoubles + i);
        buf[i] = s;
    }
}

void launch_peer_allreduce(double* buf, int n, double* const* peers_dev, int rank, int world, unsigned long long epoch, int* timed_out,
                           cudaStream_t st) {
    k_peer_allreduce<<<1, 256, 0, st>>>(buf, n, peers_dev, rank, world, epoch, timed_out);
}

}  // namespace calcomm
