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
// hanging the GPU and raises a host-mapped flag; the host checks the flag after the stream synchronisation that
// follows every all-reduce and fails the call with CAL_ERR_COMM (Comm::check_timeout) — on every rank, because the
// peers of the rank that gave up time out at the next collective.
#include <stdint.h>

#include "comm.h"
#include "comm_peer.cuh"

namespace calcomm {

__global__ void __launch_bounds__(256) k_peer_allreduce(double* __restrict__ buf, int n, PeerArgs a) {
    __shared__ int bad;
    peer_allreduce_cta(buf, n, a, &bad);
}

void launch_peer_allreduce(double* buf, int n, const PeerArgs& a, cudaStream_t st) {
    k_peer_allreduce<<<1, 256, 0, st>>>(buf, n, a);
}

}  // namespace calcomm
