// Thin NCCL wrapper for the one collective on the refinement path: the FP64
// sum-allreduce of the per-camera normal-equation blocks (and, for the
// per-view kinds, of the Schur complement) once per evaluation.  One process
// per GPU; libnccl is dlopen'ed so single-GPU use has no NCCL dependency.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string>

#include "comm_peer.cuh"

namespace calcomm {

constexpr int kPeerHandleBytes = 64;    // sizeof(cudaIpcMemHandle_t)
void launch_peer_allreduce(double* buf, int n, const PeerArgs& a, cudaStream_t st);

class Comm {
  public:
    static bool unique_id(uint8_t out128[128], std::string* err);
    static Comm* create(const uint8_t id128[128], int rank, int world, std::string* err);
    ~Comm();
    bool allreduce_sum(double* dev_buf, size_t n, cudaStream_t st);
    // small host-side reductions (sum / max) routed through a device staging buffer
    bool allreduce_host(double* host_buf, size_t n, bool is_max = false);
    // NVLink peer-memory path (comm_peer.cu): export this rank's receive region, then map the peers' regions
    bool peer_export(uint8_t handle_out[kPeerHandleBytes]);
    bool peer_enable(const uint8_t* handles /*[world][kPeerHandleBytes]*/);
    bool peer_enabled() const { return peer_on_; }
    // arguments of the NEXT peer all-reduce of n doubles for a kernel that performs it inline (k_tile_reduce); false
    // when the peer path is off or the block too large — the caller then uses allreduce_sum.  Every rank must make
    // the same sequence of peer_args / allreduce_sum calls (one epoch each).
    bool peer_args(size_t n, PeerArgs* out) {
        if (!peer_on_ || n > (size_t)kPeerMaxDoubles) return false;
        *out = PeerArgs{peers_dev_, rank_, world_, ++epoch_, timed_out_dev_};
        return true;
    }
    // after a stream synchronisation: did a peer all-reduce give up waiting?  (then its result is NaN-poisoned)
    bool check_timeout() { if (timed_out_ && *timed_out_) { err_ = "peer all-reduce timed out waiting for another rank"; return false; } return true; }
    void peer_disable() { peer_on_ = false; }
    // test hook: all-reduce a host vector through the device (peer path if enabled and use_peer, else NCCL)
    bool allreduce_test(double* host_buf, size_t n, bool use_peer);
    const std::string& error() const { return err_; }
    int rank() const { return rank_; }
    int world() const { return world_; }

  private:
    Comm() = default;
    void* comm_ = nullptr;
    double* stage_ = nullptr;
    cudaStream_t st_ = nullptr;
    int rank_ = 0, world_ = 1;
    std::string err_;
    // peer path
    double* recv_ = nullptr;            // my region: slots [2][world][kPeerMaxDoubles] + flags [2][world]
    double** peers_dev_ = nullptr;      // device array [world] of region base pointers
    void* opened_[64] = {nullptr};      // cudaIpcOpenMemHandle results to close
    volatile int* timed_out_ = nullptr; // pinned host flag set by the kernel's bounded wait (mapped: no copy needed to read it)
    int* timed_out_dev_ = nullptr;      // its device address
    unsigned long long epoch_ = 0;
    bool peer_on_ = false;
};

}  // namespace calcomm

// the opaque communicator of the C ABI (cal_comm_create): shared by the translation units that accept one
struct cal_comm { calcomm::Comm* c = nullptr; };
