#include "comm.h"

#include <dlfcn.h>

#include <cstring>
#include <string>

namespace calcomm {
namespace {

struct NcclUniqueId { char internal[128]; };
using ncclComm_t = void*;
typedef int (*fn_GetUniqueId)(NcclUniqueId*);
typedef int (*fn_CommInitRank)(ncclComm_t*, int, NcclUniqueId, int);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fn_CommDestroy)(ncclComm_t);
typedef const char* (*fn_GetErrorString)(int);

struct Api {
    void* lib = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
};

Api* api(std::string* err) {
    static Api a;
    static bool tried = false;
    if (!tried) {
        tried = true;
        // torch's bundled libnccl is already mapped when torch.distributed is in the
        // process; RTLD_NOLOAD finds it by soname, otherwise fall back to the system one.
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { a.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (a.lib) break; }
        if (!a.lib) for (const char* n : names) { a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.lib) break; }
        if (a.lib) {
            a.GetUniqueId = (fn_GetUniqueId)dlsym(a.lib, "ncclGetUniqueId");
            a.CommInitRank = (fn_CommInitRank)dlsym(a.lib, "ncclCommInitRank");
            a.AllReduce = (fn_AllReduce)dlsym(a.lib, "ncclAllReduce");
            a.CommDestroy = (fn_CommDestroy)dlsym(a.lib, "ncclCommDestroy");
            a.GetErrorString = (fn_GetErrorString)dlsym(a.lib, "ncclGetErrorString");
        }
    }
    if (!a.lib || !a.GetUniqueId || !a.CommInitRank || !a.AllReduce) {
        if (err) *err = "libnccl.so.2 not found or incomplete";
        return nullptr;
    }
    return &a;
}
constexpr int kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;

}  // namespace

bool Comm::unique_id(uint8_t out128[128], std::string* err) {
    Api* a = api(err);
    if (!a) return false;
    NcclUniqueId id;
    const int rc = a->GetUniqueId(&id);
    if (rc != 0) { if (err) *err = std::string("ncclGetUniqueId: ") + (a->GetErrorString ? a->GetErrorString(rc) : "error"); return false; }
    std::memcpy(out128, id.internal, 128);
    return true;
}

Comm* Comm::create(const uint8_t id128[128], int rank, int world, std::string* err) {
    Api* a = api(err);
    if (!a) return nullptr;
    NcclUniqueId id; std::memcpy(id.internal, id128, 128);
    Comm* c = new Comm;
    c->rank_ = rank; c->world_ = world;
    ncclComm_t comm = nullptr;
    const int rc = a->CommInitRank(&comm, world, id, rank);
    if (rc != 0) { if (err) *err = std::string("ncclCommInitRank: ") + (a->GetErrorString ? a->GetErrorString(rc) : "error"); delete c; return nullptr; }
    c->comm_ = comm;
    if (cudaMalloc(reinterpret_cast<void**>(&c->stage_), 64 * sizeof(double)) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->st_, cudaStreamNonBlocking) != cudaSuccess) {
        if (err) *err = "comm staging allocation failed"; delete c; return nullptr;
    }
    return c;
}

bool Comm::peer_export(uint8_t handle_out[kPeerHandleBytes]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == kPeerHandleBytes, "IPC handle size");
    if (world_ > 64) { err_ = "peer path supports at most 64 ranks"; return false; }
    if (!recv_) {
        const size_t bytes = ((size_t)2 * world_ * kPeerMaxDoubles + (size_t)2 * world_) * sizeof(double);
        int* flag = nullptr;
        if (cudaMalloc(reinterpret_cast<void**>(&recv_), bytes) != cudaSuccess || cudaMemset(recv_, 0, bytes) != cudaSuccess ||
            cudaHostAlloc(reinterpret_cast<void**>(&flag), sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer(reinterpret_cast<void**>(&timed_out_dev_), flag, 0) != cudaSuccess) {
            err_ = std::string("peer region allocation failed: ") + cudaGetErrorString(cudaGetLastError()); return false;
        }
        *flag = 0; timed_out_ = flag;
        cudaDeviceSynchronize();
    }
    cudaIpcMemHandle_t hnd;
    if (cudaIpcGetMemHandle(&hnd, recv_) != cudaSuccess) { err_ = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(cudaGetLastError()); return false; }
    std::memcpy(handle_out, &hnd, kPeerHandleBytes);
    return true;
}

bool Comm::peer_enable(const uint8_t* handles) {
    if (!recv_) { err_ = "peer_export must be called first"; return false; }
    double* host_ptrs[64];
    for (int r = 0; r < world_; ++r) {
        if (r == rank_) { host_ptrs[r] = recv_; continue; }
        cudaIpcMemHandle_t hnd; std::memcpy(&hnd, handles + (size_t)r * kPeerHandleBytes, kPeerHandleBytes);
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            err_ = std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(cudaGetLastError());
            return false;
        }
        opened_[r] = p; host_ptrs[r] = static_cast<double*>(p);
    }
    if (cudaMalloc(reinterpret_cast<void**>(&peers_dev_), sizeof(double*) * world_) != cudaSuccess ||
        cudaMemcpy(peers_dev_, host_ptrs, sizeof(double*) * world_, cudaMemcpyHostToDevice) != cudaSuccess) {
        err_ = "peer pointer table allocation failed"; return false;
    }
    peer_on_ = true;
    return true;
}

bool Comm::allreduce_test(double* host_buf, size_t n, bool use_peer) {
    if (n > (size_t)kPeerMaxDoubles) { err_ = "allreduce_test: too large"; return false; }
    double* d = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&d), n * sizeof(double)) != cudaSuccess) { err_ = "allreduce_test: allocation failed"; return false; }
    bool ok = cudaMemcpyAsync(d, host_buf, n * sizeof(double), cudaMemcpyHostToDevice, st_) == cudaSuccess;
    const bool saved = peer_on_;
    if (!use_peer) peer_on_ = false;
    ok = ok && allreduce_sum(d, n, st_);
    peer_on_ = saved;
    ok = ok && cudaMemcpyAsync(host_buf, d, n * sizeof(double), cudaMemcpyDeviceToHost, st_) == cudaSuccess && cudaStreamSynchronize(st_) == cudaSuccess;
    ok = ok && check_timeout();
    cudaFree(d);
    if (!ok && err_.empty()) err_ = "allreduce_test failed";
    return ok;
}

Comm::~Comm() {
    for (int r = 0; r < 64; ++r) if (opened_[r]) cudaIpcCloseMemHandle(opened_[r]);
    if (peers_dev_) cudaFree(peers_dev_);
    if (recv_) cudaFree(recv_);
    if (timed_out_) cudaFreeHost(const_cast<int*>(timed_out_));
    Api* a = api(nullptr);
    if (comm_ && a && a->CommDestroy) a->CommDestroy(comm_);
    if (stage_) cudaFree(stage_);
    if (st_) cudaStreamDestroy(st_);
}

bool Comm::allreduce_sum(double* dev_buf, size_t n, cudaStream_t st) {
    if (peer_on_ && n <= (size_t)kPeerMaxDoubles) {
        PeerArgs a; peer_args(n, &a);
        launch_peer_allreduce(dev_buf, (int)n, a, st);
        if (cudaGetLastError() != cudaSuccess) { err_ = "peer all-reduce launch failed"; return false; }
        return true;
    }
    Api* a = api(&err_);
    if (!a) return false;
    const int rc = a->AllReduce(dev_buf, dev_buf, n, kNcclFloat64, kNcclSum, comm_, st);
    if (rc != 0) { err_ = std::string("ncclAllReduce: ") + (a->GetErrorString ? a->GetErrorString(rc) : "error"); return false; }
    return true;
}

bool Comm::allreduce_host(double* host_buf, size_t n, bool is_max) {
    Api* a = api(&err_);
    if (!a || n > 64) { err_ = "allreduce_host: bad size"; return false; }
    if (cudaMemcpyAsync(stage_, host_buf, n * sizeof(double), cudaMemcpyHostToDevice, st_) != cudaSuccess) { err_ = "stage H2D failed"; return false; }
    const int rc = a->AllReduce(stage_, stage_, n, kNcclFloat64, is_max ? kNcclMax : kNcclSum, comm_, st_);
    if (rc != 0) { err_ = std::string("ncclAllReduce: ") + (a->GetErrorString ? a->GetErrorString(rc) : "error"); return false; }
    if (cudaMemcpyAsync(host_buf, stage_, n * sizeof(double), cudaMemcpyDeviceToHost, st_) != cudaSuccess ||
        cudaStreamSynchronize(st_) != cudaSuccess) { err_ = "stage D2H failed"; return false; }
    return true;
}

}  // namespace calcomm
