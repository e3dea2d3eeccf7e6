// Thin NCCL wrapper for the one collective on the refinement path: the FP64
// sum-allreduce of the per-camera normal-equation blocks (and, for the
// per-view kinds, of the Schur complement) once per evaluation.  One process
// per GPU; libnccl is dlopen'ed so single-GPU use has no NCCL dependency.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string>

namespace calcomm {

constexpr int kPeerMaxDoubles = 4096;   // largest block the peer-memory all-reduce takes (32 KB)
constexpr int kPeerHandleBytes = 64;    // sizeof(cudaIpcMemHandle_t)
void launch_peer_allreduce(double* buf, int n, double* const* peers_dev, int rank, int world, unsigned long long epoch, int* timed_out,
                           cudaStream_t st);

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
    int* timed_out_ = nullptr;          // device flag set by the kernel's bounded wait
    unsigned long long epoch_ = 0;
    bool peer_on_ = false;
};

}  // namespace calcomm
