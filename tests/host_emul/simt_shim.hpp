// TEST INFRASTRUCTURE — a lock-step SIMT shim that lets g++ compile and run the PRODUCT's warp-level CUDA
// kernel sources (the *_kernel.cuh / *.cuh device headers of calibration_b200/csrc) on the CPU, with no GPU.
//
// Every CUDA thread of one CTA becomes an OS thread; the warp collectives (__shfl_sync, __shfl_xor_sync,
// __ballot_sync, __any_sync, __all_sync, __syncwarp) and __syncthreads are rendezvous points of the 32 lanes
// of a warp (the CTA for __syncthreads), which is exactly the guarantee the kernels rely on: every collective
// is reached by all lanes named in its (full) mask.  CTAs run one after the other; shared memory is one static
// buffer.  It is slow (two barrier phases per collective) and meant for a handful of small problems: it checks
// the warp decomposition of a kernel — ballots, shuffles, lane-partitioned sums, batch bookkeeping — against the
// oracle, which the scalar host emulations (emul.cpp, plane_emul.cpp) cannot.
//
// Include this header BEFORE the product's device headers.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

// the CUDA runtime header (pulled in by some product headers for cudaStream_t) defines the same decorations for a
// host compiler: take it first where it exists, then impose the shim's meanings
#if __has_include(<cuda_runtime.h>)
#include <cuda_runtime.h>
#endif
#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __shared__
#undef __align__
#undef __launch_bounds__
#undef __grid_constant__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
// static __shared__ arrays become function-local statics (one CTA runs at a time); a translation unit whose kernels
// declare dynamic shared memory (`extern __shared__`) leaves SIMT_SHARED_STORAGE empty instead
#ifndef SIMT_SHARED_STORAGE
#define SIMT_SHARED_STORAGE
#endif
#define __shared__ SIMT_SHARED_STORAGE
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
#define __grid_constant__

using std::isfinite;
using std::max;
using std::min;

namespace simt {

struct Dim3 { unsigned x = 1, y = 1, z = 1; };

struct Warp {
    std::barrier<> bar{32};
    uint64_t slot[32];
};
struct Cta {
    explicit Cta(int n_threads) : bar(n_threads), warps((n_threads + 31) / 32) {
        for (auto& w : warps) w = std::make_unique<Warp>();
    }
    std::barrier<> bar;
    std::vector<std::unique_ptr<Warp>> warps;
};

inline thread_local Warp* tl_warp = nullptr;
inline thread_local Cta* tl_cta = nullptr;
inline thread_local int tl_lane = 0;

template <class T>
inline uint64_t to_bits(T v) { static_assert(sizeof(T) <= 8); uint64_t b = 0; std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T>
inline T from_bits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }

// all lanes publish a value, rendezvous, read what they need, rendezvous again (so the slots can be reused)
template <class T, class F>
inline auto exchange(T mine, F&& read) {
    Warp& w = *tl_warp;
    w.slot[tl_lane] = to_bits(mine);
    w.bar.arrive_and_wait();
    auto r = read(w.slot);
    w.bar.arrive_and_wait();
    return r;
}

// ---- stand-ins for the asynchronous-copy machinery of the kernels that stage tiles with TMA (CALIB_SIMT_SHIM
// branches of k1_kernel.cuh): the bulk copy is a memcpy by the issuing thread, the mbarrier a completion counter
// every thread polls, the named barrier `bar.sync 1, n` (n = the whole CTA in those kernels) the CTA rendezvous ----
inline void mbar_init(unsigned long long* b) { __atomic_store_n(b, 0ULL, __ATOMIC_RELEASE); }
inline void bulk_copy_and_complete(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
    std::memcpy(dst, src, bytes);
    __atomic_fetch_add(b, 1ULL, __ATOMIC_RELEASE);
}
inline void mbar_wait(const unsigned long long* b, unsigned long long completions) {
    while (__atomic_load_n(b, __ATOMIC_ACQUIRE) < completions) std::this_thread::yield();
}
inline void named_barrier(int /*n_threads: the whole CTA*/) { tl_cta->bar.arrive_and_wait(); }

// Runs `kernel()` for every thread of every CTA of the grid.  The CTA size must be a multiple of 32.
inline void launch(unsigned grid, unsigned block, const std::function<void()>& kernel, unsigned grid_y = 1);

}  // namespace simt

inline thread_local simt::Dim3 threadIdx, blockIdx, blockDim, gridDim;

inline void simt::launch(unsigned grid, unsigned block, const std::function<void()>& kernel, unsigned grid_y) {
    for (unsigned by = 0; by < grid_y; ++by)
    for (unsigned b = 0; b < grid; ++b) {
        Cta cta((int)block);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block; ++t)
            th.emplace_back([&, t] {
                threadIdx.x = t; blockIdx.x = b; blockDim.x = block; gridDim.x = grid; blockIdx.y = by; gridDim.y = grid_y;
                tl_cta = &cta; tl_warp = cta.warps[t / 32].get(); tl_lane = (int)(t % 32);
                kernel();
                tl_warp->bar.arrive_and_drop();   // a lane that has left the kernel no longer takes part in rendezvous
                cta.bar.arrive_and_drop();
            });
        for (auto& x : th) x.join();
    }
}

// ---- warp collectives (full masks only, as the kernels use them) ----
inline void __syncwarp(unsigned = 0xffffffffu) { simt::tl_warp->bar.arrive_and_wait(); }
inline void __syncthreads() { simt::tl_cta->bar.arrive_and_wait(); }
template <class T>
inline T __shfl_sync(unsigned, T v, int src) {
    return simt::exchange(v, [&](const uint64_t* s) { return simt::from_bits<T>(s[src & 31]); });
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    return simt::exchange(v, [&](const uint64_t* s) { return simt::from_bits<T>(s[(simt::tl_lane ^ lane_mask) & 31]); });
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, int delta) {   // lanes whose source is out of range keep their own value
    return simt::exchange(v, [&](const uint64_t* s) { const int src = simt::tl_lane + delta; return src < 32 ? simt::from_bits<T>(s[src]) : v; });
}
inline unsigned __ballot_sync(unsigned, int pred) {
    return simt::exchange<uint64_t>(pred ? 1u : 0u, [&](const uint64_t* s) { unsigned m = 0; for (int i = 0; i < 32; ++i) m |= (unsigned)(s[i] & 1u) << i; return m; });
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }

// ---- device intrinsics ----
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T> inline T __ldcg(const T* p) { return *p; }
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
