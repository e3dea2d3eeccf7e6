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

class Comm {
  public:
    static bool unique_id(uint8_t out128[128], std::string* err);
    static Comm* create(const uint8_t id128[128], int rank, int world, std::string* err);
    ~Comm();
    bool allreduce_sum(double* dev_buf, size_t n, cudaStream_t st);
    // small host-side reductions (sum / max) routed through a device staging buffer
    bool allreduce_host(double* host_buf, size_t n, bool is_max = false);
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
};

}  // namespace calcomm
