// Host side of the refinement path: the C ABI of include/calib_b200.h, the
// problem set-up (replacing the reference's Ceres problem builders) and the
// Levenberg–Marquardt driver (replacing ceres::Solve as configured by
// solve_problem, reference src/estimation/detail/ceresutils.h:27-43).
// Damping, step acceptance, convergence tests and the covariance assembly run
// here on the host; every O(observations) or O(views) pass is a kernel of
// refine_kernels.cu.  There is no CPU fallback: a missing device is an error.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/calib_b200.h"
#include "comm.h"
#include "refine_kernels.cuh"
#include "refine_model.hpp"

using namespace calk;

namespace {

thread_local std::string g_err;
cal_status fail(cal_status s, const std::string& m) { g_err = m; return s; }

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) return fail(CAL_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// set-up temporaries of cal_refine_create: released on every exit, also the CUDA_TRY ones
struct ScopedStream {
    cudaStream_t s = nullptr;
    ~ScopedStream() { if (s) cudaStreamDestroy(s); }   // pending work still completes (CUDA defers the release)
};
struct ScopedEvent {
    cudaEvent_t e = nullptr;
    ~ScopedEvent() { if (e) cudaEventDestroy(e); }
};
template <class T> struct ScopedAsyncBuf {   // cudaMallocAsync'ed staging buffer, freed in stream order
    T* p = nullptr;
    cudaStream_t st = nullptr;               // stream the failure-path free is ordered on
    ~ScopedAsyncBuf() { if (p) cudaFreeAsync(p, st); }
    cudaError_t free_on(cudaStream_t s) { cudaError_t e = p ? cudaFreeAsync(p, s) : cudaSuccess; p = nullptr; return e; }
};

template <class T> cudaError_t dev_alloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T)); }
template <class T> cudaError_t upload(T* dst, const std::vector<T>& src, cudaStream_t st) {
    if (src.empty()) return cudaSuccess;
    return cudaMemcpyAsync(dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, st);
}

// dense lower Cholesky / solve (host, shared block only: ns <= a few hundred)
bool chol_host(std::vector<double>& A, int n) {
    for (int j = 0; j < n; ++j) {
        double s = A[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) s -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
        if (!(s > 0.0) || !std::isfinite(s)) return false;
        const double l = std::sqrt(s);
        A[(size_t)j * n + j] = l;
        for (int i = j + 1; i < n; ++i) {
            double t = A[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) t -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
            A[(size_t)i * n + j] = t / l;
        }
    }
    return true;
}
void chol_solve_host(const std::vector<double>& L, int n, double* b) {
    for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * b[k]; b[i] = s / L[(size_t)i * n + i]; }
    for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * b[k]; b[i] = s / L[(size_t)i * n + i]; }
}

void quat_plus_jacobian(const double* q, double* J) {  // QuaternionManifold::PlusJacobian, 4x3 row-major
    J[0] = -q[1]; J[1] = -q[2]; J[2] = -q[3];
    J[3] = q[0];  J[4] = q[3];  J[5] = -q[2];
    J[6] = -q[3]; J[7] = q[0];  J[8] = q[1];
    J[9] = q[2];  J[10] = -q[1]; J[11] = q[0];
}

// Page-locked blocks cost milliseconds to create and to release (cudaHostAlloc / cudaFreeHost): handles borrow theirs
// from a process-wide free list and return it on destruction; the list is never shrunk.
struct PinPool {
    std::mutex m;
    std::vector<std::pair<size_t, double*>> idle;   // (doubles, block)
    std::vector<std::pair<size_t, double*>> lent;   // blocks handed out through cal_host_borrow
    double* get(size_t n, size_t* got) {
        {
            std::lock_guard<std::mutex> lk(m);
            size_t best = idle.size();
            for (size_t i = 0; i < idle.size(); ++i)
                if (idle[i].first >= n && (best == idle.size() || idle[i].first < idle[best].first)) best = i;
            if (best < idle.size()) { double* p = idle[best].second; *got = idle[best].first; idle.erase(idle.begin() + (long)best); return p; }
        }
        const size_t cap = std::max<size_t>(n, 4096);
        double* p = nullptr;
        if (cudaHostAlloc(reinterpret_cast<void**>(&p), cap * sizeof(double), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        *got = cap;
        return p;
    }
    void put(double* p, size_t n) { if (p) { std::lock_guard<std::mutex> lk(m); idle.emplace_back(n, p); } }
};
PinPool& pin_pool() { static PinPool* pool = new PinPool; return *pool; }     // small result blocks of the handles
PinPool& host_pool() { static PinPool* pool = new PinPool; return *pool; }    // large staging blocks of callers (cal_host_borrow)

// A large result (the dense covariance of the per-view kinds: 396 MB at 2 cameras x 1 000 views) into the CALLER's
// pageable memory.  cudaMemcpy into pageable memory goes through the driver's small bounce buffers at a fraction of
// the link rate (88 ms for those 396 MB); here the device copies 32 MB chunks into two page-locked blocks of the
// process-wide pool at link rate while host threads move the previous chunk to its destination.
cudaError_t download_large(double* dst, const double* src_dev, size_t n, cudaStream_t st) {
    if (n <= ((size_t)64 << 10)) {               // small: one plain copy
        cudaError_t e = cudaMemcpyAsync(dst, src_dev, n * sizeof(double), cudaMemcpyDeviceToHost, st);
        return e != cudaSuccess ? e : cudaStreamSynchronize(st);
    }
    // chunks of 32 MB for the big matrices, a quarter of the payload (at least 2 MB) for the 10 MB-class vectors
    const size_t kChunk = std::min<size_t>((size_t)4 << 20, std::max<size_t>((size_t)256 << 10, (n + 3) / 4));
    struct Block { double* p = nullptr; size_t got = 0; cudaEvent_t ev = nullptr; ~Block() { if (ev) cudaEventDestroy(ev); host_pool().put(p, got); } } blk[2];
    for (Block& b : blk) {
        b.p = host_pool().get(kChunk, &b.got);
        if (!b.p) return cudaErrorMemoryAllocation;
        cudaError_t e = cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int n_thr = (int)std::max<size_t>(1, std::min<size_t>(std::min(8u, hw), kChunk / ((size_t)512 << 10)));   // at least 4 MB per copy thread
    auto scatter = [&](double* to, const double* from, size_t cnt) {
        std::vector<std::thread> pool;
        const size_t per = (cnt + n_thr - 1) / n_thr;
        for (int t = 1; t < n_thr; ++t) {
            const size_t a = std::min(cnt, per * t), b = std::min(cnt, per * (t + 1));
            if (b <= a) continue;
            try { pool.emplace_back([=] { std::memcpy(to + a, from + a, (b - a) * sizeof(double)); }); }
            catch (...) { std::memcpy(to + a, from + a, (b - a) * sizeof(double)); }   // no thread to be had: copy here
        }
        std::memcpy(to, from, std::min(cnt, per) * sizeof(double));
        for (std::thread& th : pool) th.join();
    };
    const size_t n_chunks = (n + kChunk - 1) / kChunk;
    for (size_t c = 0; c <= n_chunks; ++c) {
        if (c < n_chunks) {
            const size_t off = c * kChunk, cnt = std::min(kChunk, n - off);
            cudaError_t e = cudaMemcpyAsync(blk[c & 1].p, src_dev + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaEventRecord(blk[c & 1].ev, st);
            if (e != cudaSuccess) { cudaStreamSynchronize(st); return e; }   // (nothing may still be writing a block that goes back to the pool)
        }
        if (c > 0) {   // chunk c - 1 has landed (chunk c is on its way): move it out
            const size_t off = (c - 1) * kChunk, cnt = std::min(kChunk, n - off);
            cudaError_t e = cudaEventSynchronize(blk[(c - 1) & 1].ev);
            if (e != cudaSuccess) { cudaStreamSynchronize(st); return e; }
            scatter(dst + off, blk[(c - 1) & 1].p, cnt);
        }
    }
    return cudaSuccess;
}

}  // namespace

struct cal_refine_handle : calk::HostModel {
    int device = 0;
    cudaStream_t st = nullptr;
    DevLayout L{};
    EvalBuffers B{};
    ViewBuffers V{};
    std::vector<void*> allocs;
    int64_t n_blocks = 0, n_obs = 0;
    ReduceDesc R{};
    int n_syrk_cta = 1;
    std::vector<int64_t> blk_orig_host;   // device block -> original block
    std::vector<int32_t> blk_cam_host, blk_view_host;
    std::vector<int32_t> blk_len_host;    // observations per ORIGINAL residual block
    std::vector<char> view_free_host;
    std::vector<int32_t> view_blk_off_host, view_blk_idx_host;
    // host mirrors of the last Jacobian evaluation (shared block)
    std::vector<double> Hss, gs, cam_sums;
    double cost = 0;
    calcomm::Comm* comm = nullptr;  // not owned (cal_comm_create / cal_comm_destroy)
    // Small results come back through ONE page-locked block, so that their device-to-host copies are truly
    // asynchronous (a copy into pageable memory blocks the host like a synchronisation):
    //   [0, 4) red of the step  | [4, 8) red of the candidate | [8, 12) red of the norms | [12, 16) flags (int32 pairs)
    //   [16, 16 + ns) y of the reduced solve | then cam_sums [n_cams * NV]
    double* pin = nullptr; size_t pin_doubles = 0;
    double* dSm = nullptr; double* dgss = nullptr; int32_t* dinfo = nullptr;   // reduced system of one LM iteration (device)
    // second set of the per-block products and per-view blocks a Jacobian pass writes (per-view kinds): the speculative pass
    // at an LM candidate fills it while the accepted point's set stays intact (swap_view_sets)
    double *alt_Hvv = nullptr, *alt_gv = nullptr, *alt_Evc = nullptr, *alt_Evi = nullptr, *alt_Hpp = nullptr, *alt_gp = nullptr;
    void swap_view_sets() {
        std::swap(B.blk_Hvv, alt_Hvv); std::swap(B.blk_gv, alt_gv); std::swap(B.blk_Evc, alt_Evc); std::swap(B.blk_Evi, alt_Evi);
        std::swap(V.Hpp, alt_Hpp); std::swap(V.gp, alt_gp);
    }
    double* pin_red(int k) const { return pin + 4 * k; }
    int32_t* pin_flags() const { return reinterpret_cast<int32_t*>(pin + 12); }
    double* pin_y() const { return pin + 16; }
    double* pin_sums() const { return pin + 16 + std::max(ns, 1); }
    // counters
    int64_t launches = 0;
    float bench_setup_ms = 0, bench_between_ms = 0;   // breakdown of the last cal_refine_bench_pass

    // all device buffers of a handle are carved from one arena (one cudaMalloc / cudaFree)
    unsigned char* arena = nullptr; size_t arena_size = 0, arena_used = 0;
    // stream-ordered allocation from the device's default memory pool (release threshold raised in
    // cal_refine_create), so repeated create / destroy cycles reuse the same physical memory
    cudaError_t arena_reserve(size_t bytes) {
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&arena), bytes, st);
        if (e == cudaSuccess) { arena_size = bytes; arena_used = 0; }
        return e;
    }
    template <class T> cudaError_t alloc(T** p, size_t n) {
        const size_t bytes = (std::max<size_t>(n, 1) * sizeof(T) + 255) / 256 * 256;
        if (arena && arena_used + bytes <= arena_size) { *p = reinterpret_cast<T*>(arena + arena_used); arena_used += bytes; return cudaSuccess; }
        cudaError_t e = dev_alloc(p, n);
        if (e == cudaSuccess) allocs.push_back(*p);
        return e;
    }
    ~cal_refine_handle() {
        for (void* p : allocs) cudaFree(p);
        if (arena && st) { cudaFreeAsync(arena, st); cudaStreamSynchronize(st); }
        pin_pool().put(pin, pin_doubles);
        if (st) cudaStreamDestroy(st);
    }
};

namespace {

cal_status validate(const cal_problem_desc& d) {
    if (d.kind < 0 || d.kind > 2 || d.model < 0 || d.model > 1) return fail(CAL_ERR_INVALID_ARGUMENT, "unknown problem kind / camera model");
    if (d.n_cams <= 0) return fail(CAL_ERR_INVALID_ARGUMENT, "No camera intrinsics provided");            // bundle.cpp:139-141
    if (d.kind == CAL_KIND_INTRINSICS && std::max(d.n_views, d.n_views_total) < 4)
        return fail(CAL_ERR_INVALID_ARGUMENT, "Insufficient views for calibration (at least 4 required).");  // intrinsics.cpp:92-96
    if (d.n_blocks <= 0 || d.n_obs <= 0) return fail(CAL_ERR_INVALID_ARGUMENT, "No observations provided");    // bundle.cpp:142-144
    if (d.board_n < 0) return fail(CAL_ERR_INVALID_ARGUMENT, "board_n < 0");
    const bool board = d.board_n > 0;   // shared-board form: obj_x / obj_y come from one board of board_n points
    if ((board ? (!d.board_x || !d.board_y) : (!d.obj_x || !d.obj_y)) || !d.img_u || !d.img_v || !d.block_offset || !d.block_cam)
        return fail(CAL_ERR_INVALID_ARGUMENT, "null observation arrays");
    if (d.kind == CAL_KIND_INTRINSICS && (d.n_cams != 1 || d.n_blocks != d.n_views))
        return fail(CAL_ERR_INVALID_ARGUMENT, "intrinsics: one camera and one residual block per view expected");
    if (d.kind == CAL_KIND_EXTRINSICS && !d.block_view) return fail(CAL_ERR_INVALID_ARGUMENT, "extrinsics: block_view required");
    if (d.kind == CAL_KIND_BUNDLE && !d.block_b_se3_g) return fail(CAL_ERR_INVALID_ARGUMENT, "bundle: block_b_se3_g required");
    if (d.block_offset[0] != 0 || d.block_offset[d.n_blocks] != d.n_obs) return fail(CAL_ERR_INVALID_ARGUMENT, "block_offset does not span the observations");
    for (int64_t b = 0; b < d.n_blocks; ++b) {
        // *Residual::create throws on an empty view (intrinsicresidual.h:38-40, bundleresidual.h:59-61)
        if (d.block_offset[b + 1] <= d.block_offset[b]) return fail(CAL_ERR_INVALID_ARGUMENT, "No observations provided");
        if (board && d.block_offset[b + 1] - d.block_offset[b] != d.board_n)
            return fail(CAL_ERR_INVALID_ARGUMENT, "shared-board form: every residual block must hold exactly board_n observations");
        if (d.block_cam[b] < 0 || d.block_cam[b] >= d.n_cams) return fail(CAL_ERR_INVALID_ARGUMENT, "block_cam out of range");
        if (d.kind == CAL_KIND_EXTRINSICS && (d.block_view[b] < 0 || d.block_view[b] >= d.n_views))
            return fail(CAL_ERR_INVALID_ARGUMENT, "Incompatible pose vector sizes for joint optimization");  // extrinsics.cpp:163-171
    }
    return CAL_OK;
}

}  // namespace

extern "C" const char* cal_last_error(void) { return g_err.c_str(); }
// internal: lets the other translation units of the library report through cal_last_error()
extern "C" void cal_set_last_error_(const char* msg) { g_err = msg ? msg : ""; }

// Page-locked host memory for callers that stage large inputs (the C++ adapter packs std::vector<BundleObservation>
// into it): borrowed from a process-wide pool that is never shrunk, so a long-running caller page-locks once.
extern "C" cal_status cal_host_borrow(size_t bytes, void** out) {
    if (!out) return fail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (cal_device_count() <= 0) return fail(CAL_ERR_CUDA, "no CUDA device: calib_b200 has no CPU fallback");
    size_t got = 0;
    double* p = host_pool().get((bytes + 7) / 8, &got);
    if (!p) return fail(CAL_ERR_CUDA, "page-locked allocation failed");
    { std::lock_guard<std::mutex> lk(host_pool().m); host_pool().lent.emplace_back(got, p); }
    *out = p;
    return CAL_OK;
}
extern "C" void cal_host_return(void* ptr) {
    if (!ptr) return;
    PinPool& pool = host_pool();
    size_t n = 0;
    {
        std::lock_guard<std::mutex> lk(pool.m);
        for (size_t i = 0; i < pool.lent.size(); ++i)
            if (pool.lent[i].second == ptr) { n = pool.lent[i].first; pool.lent.erase(pool.lent.begin() + (long)i); break; }
    }
    if (n) pool.put(static_cast<double*>(ptr), n);
}

extern "C" int cal_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" cal_status cal_refine_create(const cal_problem_desc* dp, int device, cal_refine_handle** out) {
    if (!dp || !out) return fail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    const cal_problem_desc& d = *dp;
    if (cal_status s = validate(d)) return s;
    if (cal_device_count() <= device) return fail(CAL_ERR_CUDA, "no CUDA device: calib_b200 has no CPU fallback");
    CUDA_TRY(cudaSetDevice(device));
    std::unique_ptr<cal_refine_handle> hp(new cal_refine_handle);
    cal_refine_handle& h = *hp;
    h.device = device;
    CUDA_TRY(cudaStreamCreateWithFlags(&h.st, cudaStreamNonBlocking));
    h.init_model(d);
    ProblemShape& S = h.S;
    h.n_blocks = d.n_blocks; h.n_obs = d.n_obs;
    if (d.kind != CAL_KIND_BUNDLE && h.ns + 1 > kSyrkMaxN)
        return fail(CAL_ERR_INVALID_ARGUMENT, "shared block too large for the Schur kernel (ns + 1 > 176)");

    // ---- start the big copy first: raw SoA observations into one staging buffer on a copy
    // stream, so the transfer overlaps the host-side layout construction below ----
    const bool trace = getenv("CALIB_B200_TRACE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (trace) std::fprintf(stderr, "[calib_b200] create: %-28s %8.2f ms\n", what,
                                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
    };
    {   // keep freed blocks cached in the default pool instead of returning them to the OS at every sync
        cudaMemPool_t pool; uint64_t thr = UINT64_MAX;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    ScopedStream cst_guard, us_guard;
    ScopedEvent copied_guard;
    ScopedAsyncBuf<double> raw_guard, raw_bTg_guard;
    ScopedAsyncBuf<int64_t> dsrc_guard;
    CUDA_TRY(cudaStreamCreateWithFlags(&cst_guard.s, cudaStreamNonBlocking));
    const cudaStream_t cst = cst_guard.s;
    raw_guard.st = cst; raw_bTg_guard.st = h.st; dsrc_guard.st = h.st;
    // staging buffer: [obj_x | obj_y | img_u | img_v], n_obs each; in the shared-board form the two object
    // arrays shrink to the board_n points of the one board (half the PCIe traffic)
    const size_t n_objpts = d.board_n > 0 ? (size_t)d.board_n : (size_t)d.n_obs;
    CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&raw_guard.p), sizeof(double) * (2 * n_objpts + 2 * (size_t)d.n_obs), cst));
    double* const raw = raw_guard.p;
    double* const raw_x = raw, *const raw_y = raw + n_objpts, *const raw_u = raw + 2 * n_objpts, *const raw_v = raw_u + d.n_obs;
    {
        const size_t nb = (size_t)d.n_obs * sizeof(double), nbo = n_objpts * sizeof(double);
        // cudaMemcpyDefault: the observation arrays may be host memory or already on a device
        CUDA_TRY(cudaMemcpyAsync(raw_x, d.board_n > 0 ? d.board_x : d.obj_x, nbo, cudaMemcpyDefault, cst));
        CUDA_TRY(cudaMemcpyAsync(raw_y, d.board_n > 0 ? d.board_y : d.obj_y, nbo, cudaMemcpyDefault, cst));
        CUDA_TRY(cudaMemcpyAsync(raw_u, d.img_u, nb, cudaMemcpyDefault, cst));
        CUDA_TRY(cudaMemcpyAsync(raw_v, d.img_v, nb, cudaMemcpyDefault, cst));
    }
    lap("copies issued");
    // ---- device block order: by camera (counting sort, stable), camera groups padded to 32 ----
    std::vector<int64_t>& borig = h.blk_orig_host;
    std::vector<int32_t>&bcam = h.blk_cam_host, &bview = h.blk_view_host;
    {
        std::vector<int64_t> cnt(d.n_cams + 1, 0);
        for (int64_t b = 0; b < d.n_blocks; ++b) cnt[d.block_cam[b]]++;
        std::vector<int64_t> start(d.n_cams + 1, 0);
        for (int c = 0; c < d.n_cams; ++c) start[c + 1] = start[c] + (cnt[c] + 31) / 32 * 32;
        const int64_t tot = start[d.n_cams];
        borig.assign(tot, -1); bcam.assign(tot, 0); bview.assign(tot, -1);
        for (int c = 0; c < d.n_cams; ++c) std::fill(bcam.begin() + start[c], bcam.begin() + start[c + 1], c);
        std::vector<int64_t> cur(start.begin(), start.end() - 1);
        for (int64_t b = 0; b < d.n_blocks; ++b) {
            const int64_t pos = cur[d.block_cam[b]]++;
            borig[pos] = b;
            bview[pos] = d.kind == CAL_KIND_INTRINSICS ? (int32_t)b : (d.kind == CAL_KIND_EXTRINSICS ? d.block_view[b] : -1);
        }
    }
    const int64_t nblk = (int64_t)borig.size();
    h.L.n_blk = nblk;
    h.blk_len_host.resize(d.n_blocks);
    for (int64_t b = 0; b < d.n_blocks; ++b) h.blk_len_host[b] = (int32_t)(d.block_offset[b + 1] - d.block_offset[b]);
    // ---- segments ----
    int64_t max_len = 0;
    for (int64_t b = 0; b < d.n_blocks; ++b) max_len = std::max<int64_t>(max_len, d.block_offset[b + 1] - d.block_offset[b]);
    // Large problems: one segment per residual block and a (zero-length) segment for every padding
    // block, so that segment s == device block s and K1 can finish each block in its epilogue.
    // Small problems: blocks are cut into short segments so that they still fill the machine.
    bool fused = d.n_blocks >= 16384;
    if (const char* e = getenv("CALIB_B200_FUSED")) fused = e[0] == '1';
    h.L.fused = fused ? 1 : 0;
    const int64_t target = fused ? std::max<int64_t>(max_len, 1)
                                 : std::min<int64_t>(std::max<int64_t>(8, (d.n_obs + 65535) / 65536), std::max<int64_t>(max_len, 1));
    std::vector<int32_t> seg_len, seg_blk, seg_cam, blk_seg_off(nblk + 1, 0), blk_vfree(nblk, 0);
    std::vector<int64_t> seg_src;
    {
        const size_t guess = (size_t)(d.n_obs / target + nblk + 64);
        seg_len.reserve(guess); seg_blk.reserve(guess); seg_cam.reserve(guess); seg_src.reserve(guess);
    }
    h.view_free_host.assign(std::max(S.n_views, 0), 0);
    for (int v = 0; v < S.n_views; ++v) h.view_free_host[v] = !h.pbs[h.pb_viewq(v)].constant;
    h.L.one_seg_per_blk = 1;
    for (int64_t b = 0; b < nblk; ++b) {
        blk_seg_off[b] = (int32_t)seg_len.size();
        if (borig[b] < 0) {
            if (fused) { seg_len.push_back(0); seg_blk.push_back((int32_t)b); seg_cam.push_back(bcam[b]); seg_src.push_back(0); }
            continue;
        }
        const int64_t o0 = d.block_offset[borig[b]], len = d.block_offset[borig[b] + 1] - o0;
        const int64_t nsegb = (len + target - 1) / target, sl = (len + nsegb - 1) / nsegb;
        if (nsegb != 1) h.L.one_seg_per_blk = 0;
        for (int64_t k = 0; k < len; k += sl) {
            seg_len.push_back((int32_t)std::min<int64_t>(sl, len - k)); seg_blk.push_back((int32_t)b);
            seg_cam.push_back(bcam[b]); seg_src.push_back(o0 + k);
        }
        blk_vfree[b] = d.kind == CAL_KIND_BUNDLE ? (d.optimize_target_pose != 0) : (bview[b] >= 0 && h.view_free_host[bview[b]]);
    }
    blk_seg_off[nblk] = (int32_t)seg_len.size();
    while (seg_len.size() % 32) { seg_len.push_back(0); seg_blk.push_back(0); seg_cam.push_back(0); seg_src.push_back(0); }
    const int64_t nseg = (int64_t)seg_len.size(), ntiles = nseg / 32;
    h.L.n_seg = nseg; h.L.n_tiles = ntiles;
    std::vector<int64_t> tile_off(ntiles); std::vector<int32_t> tile_depth(ntiles);
    int64_t slices = 0;
    for (int64_t t = 0; t < ntiles; ++t) {
        int32_t dep = 0; for (int l = 0; l < 32; ++l) dep = std::max(dep, seg_len[t * 32 + l]);
        tile_off[t] = slices; tile_depth[t] = dep; slices += dep;
    }
    h.L.n_slices = slices;
    lap("host layout built");

    // ---- one arena for every device buffer of the handle ----
    {
        // fused K1 keeps no per-segment local systems and no block-indexed rows (only the cost row)
        int nvt = 0; k1_tile_value_map(S, nullptr, &nvt, nullptr);
        const size_t nbr = fused ? 1 : (size_t)(S.NV - S.NE);
        const size_t seg_doubles = fused ? 9 + 2 : 9 + S.NE + 2;
        size_t bytes = (size_t)slices * 128 * 8 + (size_t)ntiles * 12 + (size_t)nseg * (12 + 8 * seg_doubles) +
                       (size_t)nblk * (4 * 4 + 8 + 8 * (36 + 2 + nbr + 12)) + (size_t)h.n_amb * 16 + (1u << 20);
        if (fused) bytes += (size_t)ntiles * nvt * 8 + (size_t)(ntiles / 32 + S.n_cams + 1) * (nvt * 8 + 64);
        if (S.n_views > 0) bytes += (size_t)nblk * 8 * (2 * (21 + 6 + 36 + 6 * std::max(S.PI, 1)) + 6 * (6 + S.PI)) + (size_t)S.n_views * 8 * (120 + 42) + 8 * 256 +
                                    (size_t)schur_num_ctas(S.n_views) * (h.ns + 1) * (h.ns + 1) * 8 + (size_t)h.ns * h.ns * 8 + (1u << 20);
        bytes += 256 * 64;  // alignment slack
        CUDA_TRY(h.arena_reserve(bytes));
    }
    lap("arena allocated");
    // ---- uploads ----
    DevLayout& L = h.L;
    CUDA_TRY(h.alloc(&L.obs, (size_t)slices * 128));
    CUDA_TRY(h.alloc(&L.tile_off, ntiles)); CUDA_TRY(h.alloc(&L.tile_depth, ntiles));
    CUDA_TRY(h.alloc(&L.seg_len, nseg)); CUDA_TRY(h.alloc(&L.seg_blk, nseg)); CUDA_TRY(h.alloc(&L.seg_cam, nseg));
    CUDA_TRY(h.alloc(&L.blk_cam, nblk)); CUDA_TRY(h.alloc(&L.blk_view, nblk)); CUDA_TRY(h.alloc(&L.blk_orig, nblk));
    CUDA_TRY(h.alloc(&L.blk_seg_off, nblk + 1)); CUDA_TRY(h.alloc(&L.blk_vfree, nblk));
    CUDA_TRY(upload(L.tile_off, tile_off, h.st)); CUDA_TRY(upload(L.tile_depth, tile_depth, h.st));
    CUDA_TRY(upload(L.seg_len, seg_len, h.st)); CUDA_TRY(upload(L.seg_blk, seg_blk, h.st)); CUDA_TRY(upload(L.seg_cam, seg_cam, h.st));
    CUDA_TRY(upload(L.blk_cam, bcam, h.st)); CUDA_TRY(upload(L.blk_view, bview, h.st)); CUDA_TRY(upload(L.blk_orig, borig, h.st));
    CUDA_TRY(upload(L.blk_seg_off, blk_seg_off, h.st)); CUDA_TRY(upload(L.blk_vfree, blk_vfree, h.st));
    double*& raw_bTg = raw_bTg_guard.p;
    if (d.kind == CAL_KIND_BUNDLE) {  // robot poses: upload AoS, permute + transpose on the device
        CUDA_TRY(h.alloc(&L.blk_bTg, (size_t)12 * nblk));
        CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&raw_bTg), sizeof(double) * 12 * (size_t)d.n_blocks, h.st));
        CUDA_TRY(cudaMemcpyAsync(raw_bTg, d.block_b_se3_g, sizeof(double) * 12 * d.n_blocks, cudaMemcpyDefault, h.st));
        launch_btg_permute(L, raw_bTg, h.st);
    }
    // one-time repack into the tile-transposed layout: queued behind the raw copy ON THE DEVICE (event), the
    // host does not wait here — the rest of the set-up below (reduction chunk tables, buffer clears, view CSR)
    // runs on a second stream and overlaps the H2D transfer; both streams are joined once at the end
    CUDA_TRY(cudaStreamCreateWithFlags(&us_guard.s, cudaStreamNonBlocking));
    const cudaStream_t us = us_guard.s;
    CUDA_TRY(cudaEventCreateWithFlags(&copied_guard.e, cudaEventDisableTiming));
    const cudaEvent_t copied = copied_guard.e;
    {
        CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&dsrc_guard.p), sizeof(int64_t) * std::max<size_t>(nseg, 1), h.st));
        int64_t* const dsrc = dsrc_guard.p;
        CUDA_TRY(upload(dsrc, seg_src, h.st));
        CUDA_TRY(cudaEventRecord(copied, cst));
        CUDA_TRY(cudaStreamWaitEvent(h.st, copied, 0));
        launch_repack(L, raw_x, raw_y, raw_u, raw_v, dsrc, d.board_n, h.st);
        CUDA_TRY(raw_guard.free_on(h.st)); CUDA_TRY(dsrc_guard.free_on(h.st));
        CUDA_TRY(raw_bTg_guard.free_on(h.st));
        lap("repack queued");
    }
    // ---- evaluation buffers ----
    EvalBuffers& B = h.B;
    {   // column chunks for the deterministic per-camera reductions (segments and blocks)
        auto make_chunks = [&](const std::vector<int32_t>& cam_of, int64_t n, std::vector<ColChunk>& ch, std::vector<int32_t>& off, int64_t kChunk) {
            off.assign(S.n_cams + 1, 0);
            int64_t i = 0;
            std::vector<std::vector<ColChunk>> per(S.n_cams);
            while (i < n) {
                const int cam = cam_of[i]; int64_t j = i;
                while (j < n && cam_of[j] == cam) ++j;
                for (int64_t k = i; k < j; k += kChunk) per[cam].push_back(ColChunk{cam, 0, k, std::min(j, k + kChunk)});
                i = j;
            }
            for (int c = 0; c < S.n_cams; ++c) { off[c] = (int32_t)ch.size(); ch.insert(ch.end(), per[c].begin(), per[c].end()); }
            off[S.n_cams] = (int32_t)ch.size();
        };
        std::vector<ColChunk> sc, bc; std::vector<int32_t> so, bo;
        std::vector<int32_t> seg_cam_trim(seg_cam.begin(), seg_cam.begin() + blk_seg_off[nblk]);
        make_chunks(seg_cam_trim, (int64_t)seg_cam_trim.size(), sc, so, 8192);
        make_chunks(bcam, nblk, bc, bo, 8192);
        if (fused) {  // tiles are camera-pure (camera groups are padded to 32 blocks = one tile)
            std::vector<int32_t> tile_cam(ntiles);
            for (int64_t t = 0; t < ntiles; ++t) tile_cam[t] = bcam[t * 32];
            std::vector<ColChunk> tc; std::vector<int32_t> to;
            make_chunks(tile_cam, ntiles, tc, to, 32);
            std::vector<int32_t> vmap; int nvt = 0; k1_tile_value_map(S, nullptr, &nvt, &vmap);
            h.R.n_tile_chunks = (int)tc.size(); h.R.nvt = nvt;
            CUDA_TRY(h.alloc(&h.R.tile_chunks, tc.size())); CUDA_TRY(h.alloc(&h.R.tile_cam_chunk_off, to.size()));
            CUDA_TRY(upload(h.R.tile_chunks, tc, us)); CUDA_TRY(upload(h.R.tile_cam_chunk_off, to, us));
            CUDA_TRY(h.alloc(&h.R.tile_tickets, S.n_cams + 1)); CUDA_TRY(cudaMemsetAsync(h.R.tile_tickets, 0, sizeof(unsigned) * (S.n_cams + 1), us));
            h.R.n_active_cams = 0;
            for (int c = 0; c < S.n_cams; ++c) h.R.n_active_cams += to[c + 1] > to[c] ? 1 : 0;
            CUDA_TRY(h.alloc(&h.B.tile_vmap, vmap.size())); CUDA_TRY(upload(h.B.tile_vmap, vmap, us));
            CUDA_TRY(h.alloc(&h.B.tile_vals, (size_t)ntiles * nvt)); CUDA_TRY(h.alloc(&h.B.partial_tile, tc.size() * (size_t)nvt));
            CUDA_TRY(cudaStreamSynchronize(us));  // the staging vectors above go out of scope
        }
        h.R.n_seg_chunks = (int)sc.size(); h.R.n_blk_chunks = (int)bc.size();
        CUDA_TRY(h.alloc(&h.R.seg_chunks, sc.size())); CUDA_TRY(h.alloc(&h.R.blk_chunks, bc.size()));
        CUDA_TRY(h.alloc(&h.R.seg_cam_chunk_off, so.size())); CUDA_TRY(h.alloc(&h.R.blk_cam_chunk_off, bo.size()));
        CUDA_TRY(upload(h.R.seg_chunks, sc, us)); CUDA_TRY(upload(h.R.blk_chunks, bc, us));
        CUDA_TRY(upload(h.R.seg_cam_chunk_off, so, us)); CUDA_TRY(upload(h.R.blk_cam_chunk_off, bo, us));
        CUDA_TRY(cudaStreamSynchronize(us));
    }
    const int n_brows = fused ? 1 : S.NV - S.NE;
    CUDA_TRY(h.alloc(&B.x, h.n_amb)); CUDA_TRY(h.alloc(&B.camc, S.n_cams)); CUDA_TRY(h.alloc(&B.camT, (size_t)S.n_cams * 36));
    CUDA_TRY(h.alloc(&B.seg_frame, (size_t)9 * nseg)); CUDA_TRY(h.alloc(&B.blk_Tv, (size_t)36 * nblk));
    if (!fused) CUDA_TRY(h.alloc(&B.segN, (size_t)S.NE * nseg));
    CUDA_TRY(h.alloc(&B.seg_ssr, nseg)); CUDA_TRY(h.alloc(&B.blk_ssr, nblk));
    CUDA_TRY(h.alloc(&B.partial, (size_t)std::max(h.R.n_seg_chunks, 1) * S.NE)); CUDA_TRY(h.alloc(&B.partial_blk, (size_t)std::max(h.R.n_blk_chunks, 1) * n_brows));
    CUDA_TRY(h.alloc(&B.cam_sums, (size_t)S.n_cams * S.NV));
    CUDA_TRY(cudaMemsetAsync(B.cam_sums, 0, sizeof(double) * S.n_cams * S.NV, us));
    CUDA_TRY(h.alloc(&B.blk_w, nblk)); CUDA_TRY(h.alloc(&B.seg_w, nseg)); CUDA_TRY(h.alloc(&B.blk_rows, (size_t)n_brows * nblk));
    CUDA_TRY(cudaMemsetAsync(B.seg_w, 0, sizeof(double) * nseg, us));
    CUDA_TRY(cudaMemsetAsync(B.blk_rows, 0, sizeof(double) * n_brows * nblk, us));
    CUDA_TRY(cudaMemsetAsync(B.seg_frame, 0, sizeof(double) * 9 * nseg, us));
    CUDA_TRY(cudaMemsetAsync(B.seg_ssr, 0, sizeof(double) * nseg, us));
    if (S.n_views > 0) {
        CUDA_TRY(h.alloc(&B.blk_Hvv, (size_t)21 * nblk)); CUDA_TRY(h.alloc(&B.blk_gv, (size_t)6 * nblk));
        CUDA_TRY(h.alloc(&B.blk_Evc, (size_t)36 * nblk)); CUDA_TRY(h.alloc(&B.blk_Evi, (size_t)6 * std::max(S.PI, 1) * nblk));
        CUDA_TRY(cudaMemsetAsync(B.blk_Hvv, 0, sizeof(double) * 21 * nblk, us)); CUDA_TRY(cudaMemsetAsync(B.blk_gv, 0, sizeof(double) * 6 * nblk, us));
        CUDA_TRY(cudaMemsetAsync(B.blk_Evc, 0, sizeof(double) * 36 * nblk, us));
        CUDA_TRY(cudaMemsetAsync(B.blk_Evi, 0, sizeof(double) * 6 * std::max(S.PI, 1) * nblk, us));
        CUDA_TRY(h.alloc(&h.alt_Hvv, (size_t)21 * nblk)); CUDA_TRY(h.alloc(&h.alt_gv, (size_t)6 * nblk));
        CUDA_TRY(h.alloc(&h.alt_Evc, (size_t)36 * nblk)); CUDA_TRY(h.alloc(&h.alt_Evi, (size_t)6 * std::max(S.PI, 1) * nblk));
        CUDA_TRY(cudaMemsetAsync(h.alt_Hvv, 0, sizeof(double) * 21 * nblk, us)); CUDA_TRY(cudaMemsetAsync(h.alt_gv, 0, sizeof(double) * 6 * nblk, us));
        CUDA_TRY(cudaMemsetAsync(h.alt_Evc, 0, sizeof(double) * 36 * nblk, us));
        CUDA_TRY(cudaMemsetAsync(h.alt_Evi, 0, sizeof(double) * 6 * std::max(S.PI, 1) * nblk, us));
        // view CSR over device blocks
        ViewBuffers& V = h.V;
        std::vector<int32_t>& off = h.view_blk_off_host; std::vector<int32_t>& idx = h.view_blk_idx_host;
        off.assign(S.n_views + 1, 0);
        for (int64_t b = 0; b < nblk; ++b) if (bview[b] >= 0) off[bview[b] + 1]++;
        for (int v = 0; v < S.n_views; ++v) off[v + 1] += off[v];
        idx.assign(off[S.n_views], 0);
        { std::vector<int32_t> cur(off.begin(), off.end() - 1); for (int64_t b = 0; b < nblk; ++b) if (bview[b] >= 0) idx[cur[bview[b]]++] = (int32_t)b; }
        {   // one residual block per (view, camera): the Schur kernels place a view's rows per camera and would overwrite a duplicate
            std::vector<char> seen(S.n_cams, 0);
            for (int v = 0; v < S.n_views; ++v) {
                for (int k = off[v]; k < off[v + 1]; ++k) {
                    char& s = seen[bcam[idx[k]]];
                    if (s) return fail(CAL_ERR_INVALID_ARGUMENT, "a (view, camera) pair appears in more than one residual block");
                    s = 1;
                }
                for (int k = off[v]; k < off[v + 1]; ++k) seen[bcam[idx[k]]] = 0;
            }
        }
        std::vector<int32_t> vfree(S.n_views), cq(S.n_cams), ct(S.n_cams), ci(S.n_cams);
        for (int v = 0; v < S.n_views; ++v) vfree[v] = h.view_free_host[v];
        for (int c = 0; c < S.n_cams; ++c) {
            cq[c] = d.kind == CAL_KIND_EXTRINSICS ? h.pbs[h.pb_camq(c)].toff : -1;
            ct[c] = d.kind == CAL_KIND_EXTRINSICS ? h.pbs[h.pb_camt(c)].toff : -1;
            ci[c] = h.pbs[h.pb_intr(c)].toff;
        }
        const int nv = S.n_views, ns = h.ns;
        h.n_syrk_cta = schur_num_ctas(nv);
        CUDA_TRY(h.alloc(&V.view_blk_off, nv + 1)); CUDA_TRY(h.alloc(&V.view_blk_idx, idx.size())); CUDA_TRY(h.alloc(&V.view_free, nv));
        CUDA_TRY(h.alloc(&V.cam_col_q, S.n_cams)); CUDA_TRY(h.alloc(&V.cam_col_t, S.n_cams)); CUDA_TRY(h.alloc(&V.cam_col_i, S.n_cams));
        CUDA_TRY(upload(V.view_blk_off, off, us)); CUDA_TRY(upload(V.view_blk_idx, idx, us)); CUDA_TRY(upload(V.view_free, vfree, us));
        CUDA_TRY(upload(V.cam_col_q, cq, us)); CUDA_TRY(upload(V.cam_col_t, ct, us)); CUDA_TRY(upload(V.cam_col_i, ci, us));
        CUDA_TRY(h.alloc(&V.Hpp, (size_t)nv * 36)); CUDA_TRY(h.alloc(&V.gp, (size_t)nv * 6)); CUDA_TRY(h.alloc(&V.sp, (size_t)nv * 6));
        CUDA_TRY(h.alloc(&h.alt_Hpp, (size_t)nv * 36)); CUDA_TRY(h.alloc(&h.alt_gp, (size_t)nv * 6));
        CUDA_TRY(h.alloc(&V.dp, (size_t)nv * 6)); CUDA_TRY(h.alloc(&V.Lp, (size_t)nv * 36)); CUDA_TRY(h.alloc(&V.Linv, (size_t)nv * 6)); CUDA_TRY(h.alloc(&V.view_f, (size_t)nv * 6));
        V.ns = ns; V.ncp = (ns + 1 + kSyrkTile - 1) / kSyrkTile * kSyrkTile;
        CUDA_TRY(h.alloc(&V.Fd, (size_t)nv * 6 * V.ncp)); CUDA_TRY(h.alloc(&V.delta_p, (size_t)nv * 6));
        CUDA_TRY(h.alloc(&V.s_shared, ns)); CUDA_TRY(h.alloc(&V.y_shared, ns)); CUDA_TRY(h.alloc(&V.C, (size_t)ns * ns)); CUDA_TRY(h.alloc(&V.c, ns));
        CUDA_TRY(h.alloc(&V.partialC, (size_t)h.n_syrk_cta * (ns + 1) * (ns + 1))); CUDA_TRY(h.alloc(&V.red, (size_t)nv * 4));
        CUDA_TRY(h.alloc(&V.red_out, 4)); CUDA_TRY(h.alloc(&V.fail, 1));
        CUDA_TRY(h.alloc(&V.red_part, (size_t)kReduceViewsCtas * 4)); CUDA_TRY(h.alloc(&V.red_ticket, 1));
        CUDA_TRY(cudaMemsetAsync(V.red_ticket, 0, sizeof(unsigned), us));
        CUDA_TRY(cudaMemsetAsync(V.red, 0, sizeof(double) * nv * 4, us)); CUDA_TRY(cudaMemsetAsync(V.fail, 0, sizeof(int32_t), us));
        CUDA_TRY(cudaMemsetAsync(V.Fd, 0, sizeof(double) * nv * 6 * V.ncp, us));
        CUDA_TRY(cudaMemsetAsync(V.delta_p, 0, sizeof(double) * nv * 6, us));
        CUDA_TRY(cudaMemsetAsync(V.sp, 0, sizeof(double) * nv * 6, us));
    }
    CUDA_TRY(h.alloc(&h.V.x_cand, h.n_amb));
    CUDA_TRY(h.alloc(&h.dSm, (size_t)std::max(h.ns, 1) * std::max(h.ns, 1))); CUDA_TRY(h.alloc(&h.dgss, std::max(h.ns, 1))); CUDA_TRY(h.alloc(&h.dinfo, 2));
    h.pin = pin_pool().get(16 + (size_t)std::max(h.ns, 1) + (size_t)S.n_cams * S.NV, &h.pin_doubles);
    if (!h.pin) return fail(CAL_ERR_CUDA, "page-locked result block: allocation failed");
    std::memset(h.pin, 0, h.pin_doubles * sizeof(double));
    lap("host set-up done");
    CUDA_TRY(cudaStreamSynchronize(us));
    CUDA_TRY(cudaStreamSynchronize(h.st));
    CUDA_TRY(cudaGetLastError());
    lap("done");   // the scoped streams / event are released on return
    if (trace) std::fprintf(stderr, "[calib_b200] create: arena %.1f MB reserved, %.1f MB used, %zu extra allocations\n",
                            h.arena_size / 1048576.0, h.arena_used / 1048576.0, h.allocs.size());
    *out = hp.release();
    return CAL_OK;
}

extern "C" void cal_refine_destroy(cal_refine_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    delete h;
}
extern "C" int64_t cal_refine_param_count(const cal_refine_handle* h) { return h ? h->n_amb : 0; }
extern "C" int64_t cal_refine_tangent_count(const cal_refine_handle* h) { return h ? h->n_tan : 0; }

namespace {

// ---- device passes -----------------------------------------------------------
// Per-camera sums of a pass and, with a communicator, their all-reduce over the ranks.  In the fused layout the
// all-reduce runs inside the reduction kernel over NVLink peer memory (k_tile_reduce); otherwise it is its own call.
cal_status reduce_pass(cal_refine_handle& h, const EvalBuffers& B, bool jac) {
    const ProblemShape& S = h.S;
    const size_t n = (size_t)S.n_cams * (jac ? S.NV : 1);
    calcomm::PeerArgs pa;
    const bool inline_peer = h.comm && jac && h.L.fused && h.R.n_tile_chunks > 0 && h.comm->peer_args(n, &pa);
    bool done = false;
    h.launches += 2 + launch_assemble(S, h.L, B, h.R, jac ? 1 : 0, inline_peer ? &pa : nullptr, &done, h.st);
    if (h.comm && !done && !h.comm->allreduce_sum(B.cam_sums, n, h.st)) return fail(CAL_ERR_COMM, h.comm->error());
    return CAL_OK;
}

// x_dev must already hold the parameters.  After a Jacobian pass the host
// mirrors h.Hss / h.gs / h.cost describe the shared block; per-view blocks stay
// on the device (h.V.Hpp, h.V.gp, B.blk_E*).
cal_status device_pass(cal_refine_handle& h, double* x_dev, bool jac, const double* x_host, bool with_norms = false) {
    EvalBuffers B = h.B; B.x = x_dev;
    const ProblemShape& S = h.S;
    launch_setup(S, h.L, B, h.st);
    if (jac) launch_k1(S, h.L, B, h.st); else launch_cost(S, h.L, B, h.st);
    const int NV = jac ? S.NV : 1;
    if (cal_status st = reduce_pass(h, B, jac)) return st;
    const size_t n_sums = (size_t)S.n_cams * NV;
    CUDA_TRY(cudaMemcpyAsync(h.pin_sums(), B.cam_sums, n_sums * sizeof(double), cudaMemcpyDeviceToHost, h.st));
    if (jac && S.n_views > 0) {
        launch_view_gather(S, h.L, B, h.V, h.st); h.launches++;
        if (with_norms) {   // |x|^2 and the gradient max-norm of the per-view blocks at this point (red slot 2), same synchronisation
            launch_view_norms(S, B, h.V, h.st); launch_reduce_views(h.V, S.n_views, h.st); h.launches += 2;
            CUDA_TRY(cudaMemcpyAsync(h.pin_red(2), h.V.red_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, h.st));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(h.st));
    CUDA_TRY(cudaGetLastError());
    if (h.comm && !h.comm->check_timeout()) return fail(CAL_ERR_COMM, h.comm->error());
    h.cam_sums.assign(h.pin_sums(), h.pin_sums() + n_sums);
    double cost = 0;
    for (int c = 0; c < S.n_cams; ++c) cost += h.cam_sums[(size_t)c * NV + (jac ? S.NE : 0)];
    h.cost = cost;
    if (!jac) return CAL_OK;
    h.assemble_shared(h.cam_sums.data(), x_host, h.Hss, h.gs);
    return CAL_OK;
}

// x [+] delta for the shared parameter blocks (host); per-view blocks are copied.
void plus_shared(const cal_refine_handle& h, const double* x, const double* delta_shared, double t, double* xp) {
    // only the shared blocks: the per-view part of xp is never read on the host (the device holds and updates it), and
    // walking 200 000 view blocks here twice per LM iteration was a third of the iteration at 100 000 views
    const int n_shared_pb = h.S.kind == CAL_KIND_BUNDLE ? (int)h.pbs.size() : h.pb_viewq(0);
    for (int i = 0; i < n_shared_pb; ++i) {
        const PB& pb = h.pbs[i];
        if (pb.constant) { for (int j = 0; j < pb.size; ++j) xp[pb.off + j] = x[pb.off + j]; continue; }
        double dl[12];
        for (int k = 0; k < pb.tsize; ++k) dl[k] = t * delta_shared[pb.toff + k];
        if (pb.type == PB_QUAT) quat_plus(x + pb.off, dl, xp + pb.off);
        else if (pb.type == PB_INTR) {
            int k = 0;
            for (int j = 0; j < pb.size; ++j) {
                if (pb.tsize == pb.size - 1 && j == 4) { xp[pb.off + j] = x[pb.off + j]; continue; }
                xp[pb.off + j] = x[pb.off + j] + dl[k++];
            }
            // SetParameterLowerBound(fx, 0) / (fy, 0): projection inside Program::Plus
            xp[pb.off + 0] = std::max(xp[pb.off + 0], 0.0); xp[pb.off + 1] = std::max(xp[pb.off + 1], 0.0);
        } else for (int j = 0; j < pb.size; ++j) xp[pb.off + j] = x[pb.off + j] + dl[j];
    }
}

struct LMState {
    std::vector<double> x, xp;            // host ambient (view part only meaningful at start/end)
    std::vector<double> s, diag, y, step, delta, gsv;
};

// Build the dense tangent-space system (canonical order) from the last Jacobian pass.
cal_status dense_system(cal_refine_handle& h, std::vector<double>& H, std::vector<double>& g) {
    const ProblemShape& S = h.S;
    if (S.n_views == 0) { h.assemble_dense(h.Hss, h.gs, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, H, g); return CAL_OK; }
    const int nv = S.n_views, PI = S.PI; const int64_t nblk = h.L.n_blk;
    std::vector<double> Hpp((size_t)nv * 36), gp((size_t)nv * 6), Evc((size_t)36 * nblk), Evi((size_t)6 * std::max(PI, 1) * nblk);
    CUDA_TRY(cudaMemcpy(Hpp.data(), h.V.Hpp, Hpp.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(gp.data(), h.V.gp, gp.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(Evc.data(), h.B.blk_Evc, Evc.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(Evi.data(), h.B.blk_Evi, Evi.size() * sizeof(double), cudaMemcpyDeviceToHost));
    h.assemble_dense(h.Hss, h.gs, Hpp.data(), gp.data(), Evc.data(), Evi.data(), nblk, h.view_free_host.data(), h.view_blk_off_host.data(),
                     h.view_blk_idx_host.data(), h.blk_cam_host.data(), H, g);
    return CAL_OK;
}

}  // namespace

extern "C" cal_status cal_refine_eval(cal_refine_handle* h, const double* x, double* cost, double* g, double* H) {
    if (!h || !x) return fail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(h->B.x, x, sizeof(double) * h->n_amb, cudaMemcpyHostToDevice, h->st));
    const bool jac = g || H;
    if (cal_status s = device_pass(*h, h->B.x, jac, x)) return s;
    if (cost) *cost = h->cost;
    if (jac) {
        if (h->n_tan > 8192) return fail(CAL_ERR_INVALID_ARGUMENT, "dense H is only produced for n_tan <= 8192");
        std::vector<double> Hd, gd;
        if (cal_status s = dense_system(*h, Hd, gd)) return s;
        if (g) std::memcpy(g, gd.data(), gd.size() * sizeof(double));
        if (H) std::memcpy(H, Hd.data(), Hd.size() * sizeof(double));
    }
    return CAL_OK;
}

extern "C" cal_status cal_refine_cost(cal_refine_handle* h, const double* x, double* cost, double* block_ssr) {
    if (!h || !x) return fail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(h->B.x, x, sizeof(double) * h->n_amb, cudaMemcpyHostToDevice, h->st));
    if (cal_status s = device_pass(*h, h->B.x, false, x)) return s;
    if (cost) *cost = h->cost;
    if (block_ssr) {
        std::vector<double> tmp(h->L.n_blk);
        CUDA_TRY(cudaMemcpy(tmp.data(), h->B.blk_ssr, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int64_t b = 0; b < h->L.n_blk; ++b) if (h->blk_orig_host[b] >= 0) block_ssr[h->blk_orig_host[b]] = tmp[b];
    }
    return CAL_OK;
}

// Per-view diagnostics after a solve: view_errors as compute_per_view_errors fills
// them (src/estimation/optim/intrinsicssemidlt.cpp:137-151: sqrt(sum r^2 / (2 points)) per view) and the
// global RMS of src/pipeline/reports/intrinsics.cpp:12-31, from one residual-only pass on the device.
extern "C" cal_status cal_refine_view_errors(cal_refine_handle* h, const double* x, double* block_rms, double* global_rms) {
    if (!h || !x) return fail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(h->B.x, x, sizeof(double) * h->n_amb, cudaMemcpyHostToDevice, h->st));
    if (cal_status s = device_pass(*h, h->B.x, false, x)) return s;
    std::vector<double> ssr(h->L.n_blk);
    CUDA_TRY(cudaMemcpy(ssr.data(), h->B.blk_ssr, ssr.size() * sizeof(double), cudaMemcpyDeviceToHost));
    double sum_sq = 0.0, meas = 0.0;
    for (int64_t b = 0; b < h->L.n_blk; ++b) {
        const int64_t o = h->blk_orig_host[b];
        if (o < 0) continue;
        const double m = 2.0 * (double)h->blk_len_host[o];
        const double rms = m > 0 ? std::sqrt(ssr[b] / m) : 0.0;
        if (block_rms) block_rms[o] = rms;
        sum_sq += rms * rms * m; meas += m;
    }
    if (global_rms) *global_rms = meas > 0 ? std::sqrt(sum_sq / meas) : 0.0;
    return CAL_OK;
}

extern "C" cal_status cal_refine_bench_pass(cal_refine_handle* h, const double* x, int reps, int jacobian, float* ms_total,
                                            float* ms_k1, double* cost) {
    if (!h || !x || reps <= 0) return fail(CAL_ERR_INVALID_ARGUMENT, "bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(h->B.x, x, sizeof(double) * h->n_amb, cudaMemcpyHostToDevice, h->st));
    // events: [0] start, [1] end, then per repetition: before K1, after K1 (the set-up kernel runs before the first, the
    // reduction — and the all-reduce it carries on several GPUs — after the second)
    std::vector<cudaEvent_t> ev(2 * (size_t)reps + 2);
    for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaStreamSynchronize(h->st));
    CUDA_TRY(cudaEventRecord(ev[0], h->st));
    for (int r = 0; r < reps; ++r) {
        launch_setup(h->S, h->L, h->B, h->st);
        CUDA_TRY(cudaEventRecord(ev[2 + 2 * r], h->st));
        if (jacobian) launch_k1(h->S, h->L, h->B, h->st); else launch_cost(h->S, h->L, h->B, h->st);
        CUDA_TRY(cudaEventRecord(ev[3 + 2 * r], h->st));
        if (cal_status st = reduce_pass(*h, h->B, jacobian != 0)) return st;
        if (jacobian && h->S.n_views > 0) { launch_view_gather(h->S, h->L, h->B, h->V, h->st); h->launches++; }
    }
    CUDA_TRY(cudaEventRecord(ev[1], h->st));
    CUDA_TRY(cudaEventSynchronize(ev[1]));
    CUDA_TRY(cudaGetLastError());
    if (h->comm && !h->comm->check_timeout()) return fail(CAL_ERR_COMM, h->comm->error());
    float ms = 0, k1 = 0, setup = 0, tail = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, ev[0], ev[1]));
    for (int r = 0; r < reps; ++r) {
        float t = 0;
        CUDA_TRY(cudaEventElapsedTime(&t, ev[2 + 2 * r], ev[3 + 2 * r])); k1 += t;
        CUDA_TRY(cudaEventElapsedTime(&t, r == 0 ? ev[0] : ev[1 + 2 * r], ev[2 + 2 * r])); if (r > 0) tail += t; else setup += t;
    }
    // (between "after K1" of repetition r - 1 and "before K1" of repetition r lie reduction r - 1 and set-up r: tail holds their
    // sum over reps - 1 repetitions, setup the first set-up alone)
    h->bench_setup_ms = setup; h->bench_between_ms = reps > 1 ? tail / (reps - 1) : 0.0f;
    for (auto& e : ev) cudaEventDestroy(e);
    if (ms_total) *ms_total = ms;
    if (ms_k1) *ms_k1 = k1;
    if (cost) {
        const int NV = jacobian ? h->S.NV : 1;
        std::vector<double> cs((size_t)h->S.n_cams * NV);
        CUDA_TRY(cudaMemcpy(cs.data(), h->B.cam_sums, cs.size() * sizeof(double), cudaMemcpyDeviceToHost));
        double c = 0; for (int k = 0; k < h->S.n_cams; ++k) c += cs[(size_t)k * NV + (jacobian ? h->S.NE : 0)];
        *cost = c;
    }
    return CAL_OK;
}
// of the last cal_refine_bench_pass: device time of the first set-up kernel, and of [reduction (+ all-reduce) of one pass + set-up of
// the next] averaged over the repetitions — what a pass costs besides K1
extern "C" cal_status cal_refine_bench_breakdown(const cal_refine_handle* h, float* ms_first_setup, float* ms_reduce_plus_setup) {
    if (!h) return fail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    if (ms_first_setup) *ms_first_setup = h->bench_setup_ms;
    if (ms_reduce_plus_setup) *ms_reduce_plus_setup = h->bench_between_ms;
    return CAL_OK;
}

extern "C" int64_t cal_refine_launch_count(const cal_refine_handle* h) { return h ? h->launches : 0; }

extern "C" cal_status cal_refine_layout_info(const cal_refine_handle* h, int64_t* n_segments, int64_t* n_tiles,
                                             int64_t* obs_bytes, int32_t* local_entries, int32_t* k1_passes) {
    if (!h) return fail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    if (n_segments) *n_segments = h->L.n_seg;
    if (n_tiles) *n_tiles = h->L.n_tiles;
    if (obs_bytes) *obs_bytes = h->L.n_slices * 128 * (int64_t)sizeof(double);
    if (local_entries) *local_entries = h->S.NE;
    if (k1_passes) *k1_passes = k1_num_passes(h->S);
    return CAL_OK;
}

extern "C" cal_status cal_fp64_peak_tflops(int device, double* tflops) {
    if (!tflops) return fail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    if (cal_device_count() <= device) return fail(CAL_ERR_CUDA, "no CUDA device: calib_b200 has no CPU fallback");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop; CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    double* scratch; CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&scratch), sizeof(double) * blocks * threads));
    cudaStream_t st; CUDA_TRY(cudaStreamCreate(&st));
    const float ms = dfma_peak_ms(scratch, blocks, threads, iters, st);
    cudaStreamDestroy(st); cudaFree(scratch);
    CUDA_TRY(cudaGetLastError());
    *tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    return CAL_OK;
}

// ---------------------------------------------------------------------------
// Levenberg–Marquardt (Ceres 2.2 trust-region semantics)
// ---------------------------------------------------------------------------
extern "C" cal_status cal_refine_solve(cal_refine_handle* hp, const cal_optim_options* o, double* x_inout,
                                       cal_optim_result* res, double* cov) {
    if (!hp || !o || !x_inout || !res) return fail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    cal_refine_handle& h = *hp;
    const auto t_solve_start = std::chrono::steady_clock::now();
    CUDA_TRY(cudaSetDevice(h.device));
    std::memset(res, 0, sizeof *res);
    const ProblemShape& S = h.S;
    const int ns = h.ns, nv = S.n_views, na = h.n_amb;
    const bool views = nv > 0;
    const double eps = o->epsilon;
    const double min_relative_decrease = 1e-3, min_diag = 1e-6, max_diag = 1e32, max_radius = 1e16, min_radius = 1e-32;
    std::vector<double> x(x_inout, x_inout + na), xp(x);   // (xp: the per-view part is never touched on the host)
    const int n_shared_amb = views ? S.off_viewq : na;  // shared blocks are the head of x for the per-view kinds
    if (h.constrained) {  // project the start onto the feasible set (fx, fy >= 0)
        std::vector<double> z(std::max(ns, 1), 0.0);
        plus_shared(h, x.data(), z.data(), 0.0, xp.data());
        std::copy(xp.begin(), xp.begin() + n_shared_amb, x.begin());
    }
    double* xd = h.B.x; double* xc = h.V.x_cand;
    CUDA_TRY(cudaMemcpyAsync(xd, x.data(), sizeof(double) * na, cudaMemcpyHostToDevice, h.st));
    CUDA_TRY(cudaMemcpyAsync(xc, xd, sizeof(double) * na, cudaMemcpyDeviceToDevice, h.st));   // (not a second trip over PCIe: 11 MB at 100 k views)
    ViewBuffers V = h.V;
    EvalBuffers B = h.B;
    std::vector<double> s(ns, 1.0), diag(ns), y(ns), step(ns), delta(ns), gss(ns), Cs((size_t)ns * ns), cs(ns), Sm((size_t)ns * ns);
    double red[4] = {0, 0, 0, 0};
    int term = CAL_TERM_NO_CONVERGENCE, iter = 0, n_invalid = 0, jev = 0, cev = 0;
    std::string msg;

    // the Jacobian passes of this solve carry the per-view norms along (device_pass(..., with_norms)): one synchronisation
    auto norms = [&](double& x_norm, double& gmax) -> cal_status {
        // |x|_2 and |x - Plus(x, -g)|_inf (ambient), shared part on the host, view part from the pass (red slot 2)
        double x2 = 0, gm = 0;
        std::vector<double> ng(std::max(ns, 1)); for (int i = 0; i < ns; ++i) ng[i] = -h.gs[i];
        plus_shared(h, x.data(), ng.data(), 1.0, xp.data());
        for (int i = 0; i < n_shared_amb; ++i) { x2 += x[i] * x[i]; gm = std::max(gm, std::fabs(x[i] - xp[i])); }
        if (views) {
            std::memcpy(red, h.pin_red(2), sizeof red);
            if (h.comm) {  // the views of the other ranks
                if (!h.comm->allreduce_host(&red[0], 1, false) || !h.comm->allreduce_host(&red[3], 1, true)) return fail(CAL_ERR_COMM, h.comm->error());
            }
            x2 += red[0]; gm = std::max(gm, red[3]);
        }
        x_norm = std::sqrt(x2); gmax = gm;
        return CAL_OK;
    };
    auto eval_cost_at = [&](double* xdev, const double* xhost, double& c) -> cal_status {
        if (cal_status st = device_pass(h, xdev, false, xhost)) return st;
        ++cev; c = std::isfinite(h.cost) ? h.cost : std::numeric_limits<double>::max();
        return CAL_OK;
    };
    // candidate x [+] t*delta: shared part on the host (uploaded), view part on the device
    // (queued only: the evaluation of the candidate that follows synchronises once for both; finish_candidate then adds
    // the per-view part of |dx|^2 from red slot 1)
    double cand_shared_norm2 = 0;
    auto make_candidate = [&](double t) -> cal_status {
        plus_shared(h, x.data(), delta.data(), t, xp.data());
        double sn = 0; for (int i = 0; i < n_shared_amb; ++i) { const double dd = x[i] - xp[i]; sn += dd * dd; }
        cand_shared_norm2 = sn;
        CUDA_TRY(cudaMemcpyAsync(xc, xp.data(), sizeof(double) * n_shared_amb, cudaMemcpyHostToDevice, h.st));
        if (views) {
            EvalBuffers Bx = B; Bx.x = xd; ViewBuffers Vx = V; Vx.x_cand = xc;
            launch_view_plus(S, Bx, Vx, t, h.st); launch_reduce_views(V, nv, h.st); h.launches += 2;
            CUDA_TRY(cudaMemcpyAsync(h.pin_red(1), V.red_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, h.st));
        }
        return CAL_OK;
    };
    auto finish_candidate = [&](double& step_norm2) -> cal_status {   // after the synchronisation of the candidate's pass
        double sn = cand_shared_norm2;
        if (views) {
            double r2 = h.pin_red(1)[2];
            if (h.comm && !h.comm->allreduce_host(&r2, 1, false)) return fail(CAL_ERR_COMM, h.comm->error());
            sn += r2;
        }
        step_norm2 = sn;
        return CAL_OK;
    };

    if (cal_status st = device_pass(h, xd, true, x.data(), true)) return st;
    ++jev;
    double cost = h.cost;
    res->initial_cost = cost;
    // Speculative Jacobian (kinds without per-view unknowns): a candidate that follows an accepted step is evaluated with
    // the FULL fused pass — K1 yields the cost as well — so an accepted step costs one pass over the observations, not a
    // residual-only pass plus a Jacobian pass at the same point.  The accepted point's system (host mirrors) is kept
    // aside and comes back when the candidate is rejected or the solve stops on it; after a rejection the next
    // candidate is evaluated cost-only (rejections cluster).  Counters keep the reference's meaning: one cost
    // evaluation per candidate, one Jacobian evaluation per accepted point.
    // Per-view kinds: the pass also writes the per-block products and the per-view blocks on the device; the candidate's go
    // to the handle's second set (swap_view_sets), so the accepted point's stay intact for a rejected candidate's next solve.
    const bool may_speculate = getenv("CALIB_B200_NO_SPECULATION") == nullptr;
    bool speculate = may_speculate;
    bool spec_live = false;                       // the host mirrors describe the candidate, not the accepted point
    std::vector<double> keep_H, keep_g, keep_sums; double keep_cost = 0;
    auto sync_view_sets = [&]() {                 // the loop's copies of the buffer tables follow the handle's
        B.blk_Hvv = h.B.blk_Hvv; B.blk_gv = h.B.blk_gv; B.blk_Evc = h.B.blk_Evc; B.blk_Evi = h.B.blk_Evi; V.Hpp = h.V.Hpp; V.gp = h.V.gp;
    };
    auto spec_pass = [&](double* xdev, const double* xhost) -> cal_status {   // full fused pass at a candidate
        keep_H = h.Hss; keep_g = h.gs; keep_sums = h.cam_sums; keep_cost = h.cost;
        if (views) { h.swap_view_sets(); sync_view_sets(); }
        spec_live = true;
        return device_pass(h, xdev, true, xhost, views);
    };
    auto spec_restore = [&]() {
        if (!spec_live) return;
        h.Hss.swap(keep_H); h.gs.swap(keep_g); h.cam_sums.swap(keep_sums); h.cost = keep_cost; spec_live = false;
        if (views) { h.swap_view_sets(); sync_view_sets(); }
    };
    for (int i = 0; i < ns; ++i) s[i] = 1.0 / (1.0 + std::sqrt(h.Hss[(size_t)i * ns + i]));  // jacobi_scaling, once
    if (views) {
        CUDA_TRY(cudaMemcpyAsync(V.s_shared, s.data(), sizeof(double) * ns, cudaMemcpyHostToDevice, h.st));
        launch_view_scale(S, V, 1, h.st); h.launches++;
    }
    double x_norm = 0, gmax = 0;
    if (cal_status st = norms(x_norm, gmax)) return st;
    double radius = 1e4, decrease_factor = 2.0;
    bool reuse_diag = false;
    if (o->verbose) std::printf("iter      cost      cost_change  |gradient|   tr_radius\n%4d % .6e %.2e %.2e %.2e\n", 0, cost, 0.0, gmax, radius);

    for (;;) {
        if (iter >= o->max_iterations) { term = CAL_TERM_NO_CONVERGENCE; msg = "Maximum number of iterations reached."; break; }
        if (gmax <= eps) { term = CAL_TERM_CONVERGENCE; msg = "Gradient tolerance reached."; break; }
        if (radius <= min_radius) { term = CAL_TERM_CONVERGENCE; msg = "Minimum trust region radius reached."; break; }
        ++iter;
        if (!reuse_diag) {
            for (int i = 0; i < ns; ++i) diag[i] = std::min(std::max(h.Hss[(size_t)i * ns + i] * s[i] * s[i], min_diag), max_diag);
            if (views) { launch_view_scale(S, V, 0, h.st); h.launches++; }
        }
        // ---- solve (S H S + D^2) y = S g through the Schur complement ----
        bool ok = true;
        for (int i = 0; i < ns; ++i) {
            for (int j = 0; j < ns; ++j) Sm[(size_t)i * ns + j] = h.Hss[(size_t)i * ns + j] * s[i] * s[j];
            gss[i] = h.gs[i] * s[i]; y[i] = gss[i];
        }
        std::vector<double> Hs_scaled = Sm;  // undamped, for the model cost change
        for (int i = 0; i < ns; ++i) Sm[(size_t)i * ns + i] += diag[i] / radius;
        bool solved_on_device = false;   // per-view kinds: y of the reduced system and the per-view scalars are already here
        if (views) {
            CUDA_TRY(cudaMemsetAsync(V.fail, 0, sizeof(int32_t), h.st));
            launch_schur(S, h.L, B, V, ns, radius, h.st); h.launches += 4;
            if (h.comm) {
                if (!h.comm->allreduce_sum(V.C, (size_t)ns * ns, h.st) || !h.comm->allreduce_sum(V.c, ns, h.st)) return fail(CAL_ERR_COMM, h.comm->error());
            }
            int32_t* flags = h.pin_flags();
            if (ns > 0 && ns <= kReducedMaxN) {
                // the reduced system is solved where the Schur complement lies and the back-substitution follows in
                // stream order: ONE synchronisation per solve (y, the per-view scalars and the two flags come back together)
                CUDA_TRY(cudaMemcpyAsync(h.dSm, Sm.data(), sizeof(double) * ns * ns, cudaMemcpyHostToDevice, h.st));
                CUDA_TRY(cudaMemcpyAsync(h.dgss, gss.data(), sizeof(double) * ns, cudaMemcpyHostToDevice, h.st));
                launch_reduced_solve(h.dSm, h.dgss, V, ns, h.dinfo, h.st);
                launch_backsub(S, h.L, V, ns, h.st); launch_reduce_views(V, nv, h.st); h.launches += 3;
                CUDA_TRY(cudaMemcpyAsync(h.pin_y(), V.y_shared, sizeof(double) * ns, cudaMemcpyDeviceToHost, h.st));
                CUDA_TRY(cudaMemcpyAsync(h.pin_red(0), V.red_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, h.st));
                CUDA_TRY(cudaMemcpyAsync(&flags[0], V.fail, sizeof(int32_t), cudaMemcpyDeviceToHost, h.st));
                CUDA_TRY(cudaMemcpyAsync(&flags[1], h.dinfo, sizeof(int32_t), cudaMemcpyDeviceToHost, h.st));
                solved_on_device = true;
            } else {   // a shared block too wide for one CTA's shared memory: the Schur complement comes to the host
                CUDA_TRY(cudaMemcpyAsync(Cs.data(), V.C, sizeof(double) * ns * ns, cudaMemcpyDeviceToHost, h.st));
                CUDA_TRY(cudaMemcpyAsync(cs.data(), V.c, sizeof(double) * ns, cudaMemcpyDeviceToHost, h.st));
                CUDA_TRY(cudaMemcpyAsync(&flags[0], V.fail, sizeof(int32_t), cudaMemcpyDeviceToHost, h.st));
                flags[1] = 0;
            }
            CUDA_TRY(cudaStreamSynchronize(h.st));
            CUDA_TRY(cudaGetLastError());
            if (h.comm && !h.comm->check_timeout()) return fail(CAL_ERR_COMM, h.comm->error());
            int32_t failed = flags[0];
            if (h.comm) {  // a failed view factorisation on any rank invalidates the step on every rank
                double f = failed ? 1.0 : 0.0;
                if (!h.comm->allreduce_host(&f, 1, true)) return fail(CAL_ERR_COMM, h.comm->error());
                failed = f != 0.0;
            }
            if (failed) ok = false;
            if (solved_on_device) { if (flags[1]) ok = false; for (int i = 0; i < ns; ++i) y[i] = h.pin_y()[i]; }
            else for (int i = 0; i < ns; ++i) { y[i] -= cs[i]; for (int j = 0; j < ns; ++j) Sm[(size_t)i * ns + j] -= Cs[(size_t)i * ns + j]; }
        }
        if (ok && ns > 0 && !solved_on_device) { ok = chol_host(Sm, ns); if (ok) chol_solve_host(Sm, ns, y.data()); }
        reuse_diag = true;
        double model_cost_change = 0;
        if (ok) {
            double sg = 0, quad = 0;
            for (int i = 0; i < ns; ++i) { if (!std::isfinite(y[i])) ok = false; step[i] = -y[i]; sg += step[i] * gss[i]; }
            for (int i = 0; i < ns; ++i) { double row = 0; for (int j = 0; j < ns; ++j) row += Hs_scaled[(size_t)i * ns + j] * step[j]; quad += row * step[i]; }
            if (views && ok) {
                if (solved_on_device) std::memcpy(red, h.pin_red(0), sizeof red);
                else {
                    CUDA_TRY(cudaMemcpyAsync(V.y_shared, y.data(), sizeof(double) * ns, cudaMemcpyHostToDevice, h.st));
                    launch_backsub(S, h.L, V, ns, h.st); launch_reduce_views(V, nv, h.st); h.launches += 2;
                    CUDA_TRY(cudaMemcpyAsync(h.pin_red(0), V.red_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, h.st));
                    CUDA_TRY(cudaStreamSynchronize(h.st));
                    std::memcpy(red, h.pin_red(0), sizeof red);
                }
                if (h.comm) { double r2[2] = {red[0], red[1]}; if (!h.comm->allreduce_host(r2, 2)) return fail(CAL_ERR_COMM, h.comm->error()); red[0] = r2[0]; red[1] = r2[1]; }
                sg += red[0]; quad += red[1];
                if (!std::isfinite(sg) || !std::isfinite(quad)) ok = false;
            }
            model_cost_change = -(sg + 0.5 * quad);
            ok = ok && model_cost_change > 0.0;
        }
        if (!ok) {  // HandleInvalidStep
            if (++n_invalid >= 5) { term = CAL_TERM_FAILURE; msg = "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps: 5"; break; }
            radius /= decrease_factor; decrease_factor *= 2.0; reuse_diag = true;
            continue;
        }
        n_invalid = 0;
        for (int i = 0; i < ns; ++i) delta[i] = step[i] * s[i];
        double cand_cost = 0, step_norm2 = 0, t = 1.0;
        bool have_cand = false;
        if (h.constrained) {
            // Armijo projected line search; accepted at t = 1 in the normal case
            double g0 = 0; for (int i = 0; i < ns; ++i) g0 += h.gs[i] * delta[i];
            if (views) {  // per-view part of g . delta = sum step_p . (sp o gp) already in red[0]
                g0 += red[0];
            }
            for (int ls = 0; ls < 20; ++ls) {
                if (cal_status st = make_candidate(t)) return st;
                double c;
                if (speculate && ls == 0) {   // the full step: almost always accepted, so take the Jacobian along
                    if (cal_status st = spec_pass(xc, xp.data())) return st;
                    ++cev;
                    c = std::isfinite(h.cost) ? h.cost : std::numeric_limits<double>::max();
                } else if (cal_status st = eval_cost_at(xc, xp.data(), c)) return st;
                if (cal_status st = finish_candidate(step_norm2)) return st;
                const bool v = c < std::numeric_limits<double>::max();
                if (v && c <= cost + 1e-4 * g0 * t) { cand_cost = c; have_cand = true; break; }
                spec_restore();   // the search goes on with shorter steps, cost only
                double tn = 0.5 * t;
                if (v) { const double denom = 2.0 * (c - cost - g0 * t); if (denom > 0) tn = -g0 * t * t / denom; }
                t = std::min(std::max(tn, 1e-3 * t), 0.6 * t);
            }
            if (!have_cand) t = 1.0;
        }
        if (!have_cand) {
            if (cal_status st = make_candidate(t)) return st;
            if (speculate && !h.constrained) {
                if (cal_status st = spec_pass(xc, xp.data())) return st;
                ++cev;
                cand_cost = std::isfinite(h.cost) ? h.cost : std::numeric_limits<double>::max();
            } else if (cal_status st = eval_cost_at(xc, xp.data(), cand_cost)) return st;
            if (cal_status st = finish_candidate(step_norm2)) return st;
        }
        // ParameterToleranceReached / FunctionToleranceReached: tested before acceptance (B.3-5)
        if (std::sqrt(step_norm2) <= eps * (x_norm + eps)) { spec_restore(); term = CAL_TERM_CONVERGENCE; msg = "Parameter tolerance reached."; break; }
        const double cost_change = cost - cand_cost;
        if (std::fabs(cost_change) <= eps * cost) { spec_restore(); term = CAL_TERM_CONVERGENCE; msg = "Function tolerance reached."; break; }
        const double rho = cost_change / model_cost_change;
        if (rho > min_relative_decrease) {  // HandleSuccessfulStep
            std::swap(xd, xc);
            for (int i = 0; i < n_shared_amb; ++i) x[i] = xp[i];
            if (spec_live) spec_live = false;   // the candidate's system IS the new point's
            else if (cal_status st = device_pass(h, xd, true, x.data(), true)) return st;
            speculate = may_speculate;
            ++jev; cost = h.cost;
            if (cal_status st = norms(x_norm, gmax)) return st;
            radius = std::min(max_radius, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
            decrease_factor = 2.0; reuse_diag = false;
        } else {  // HandleUnsuccessfulStep
            spec_restore(); speculate = false;
            radius /= decrease_factor; decrease_factor *= 2.0; reuse_diag = true;
        }
        if (o->verbose) std::printf("%4d % .6e %.2e %.2e %.2e rho=%.2e\n", iter, cost, cost_change, gmax, radius, rho);
    }
    const bool trace = getenv("CALIB_B200_TRACE") != nullptr;
    const auto t_lm_end = std::chrono::steady_clock::now();
    if (trace) std::fprintf(stderr, "[calib_b200] solve: LM loop %8.2f ms (%d iterations, %d Jacobian + %d cost evaluations, %lld launches)\n",
                            std::chrono::duration<double, std::milli>(t_lm_end - t_solve_start).count(), iter, jev, cev, (long long)h.launches);
    // download the final parameters (last accepted point)
    CUDA_TRY(download_large(x_inout, xd, (size_t)na, h.st));
    if (xd != h.B.x) {  // keep the handle's buffers in their canonical roles
        CUDA_TRY(cudaMemcpyAsync(h.B.x, xd, sizeof(double) * na, cudaMemcpyDeviceToDevice, h.st));
        CUDA_TRY(cudaStreamSynchronize(h.st));
    }
    res->success = term == CAL_TERM_CONVERGENCE; res->iterations = iter; res->termination = term;
    res->num_jac_evals = jev; res->num_cost_evals = cev; res->final_cost = cost;
    static const char* names[] = {"CONVERGENCE", "NO_CONVERGENCE", "FAILURE"};
    std::snprintf(res->report, sizeof res->report, "Ceres Solver Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s",
                  iter, res->initial_cost, cost, names[term]);
    (void)msg;
    // ---- covariance (compute_covariance, ceresutils.h:69-126) ----
    if (cov && o->compute_covariance) {
        const int n = h.n_tan;
        if ((double)na * na * 8.0 > 8e9) return CAL_OK;  // would not fit; covariance_ok stays 0
        // The normal equations of the last accepted point are still on the device / in the host mirrors: the LM
        // evaluates the Jacobian only at accepted points and x only moves on acceptance, so no extra pass.
        std::vector<double> xf(x_inout, x_inout + na);
        if (views) {
            // Sharded by views: the shared block's covariance W comes from the ALL-REDUCED Schur complement, so it is the
            // global one on every rank; the view blocks this rank returns are those of ITS views (cov is [na][na] over the
            // rank's own parameter vector: shared blocks + local views).  Blocks between views of different ranks are
            // Z_v W Z_w^T with Z of both ranks and are not formed — the reference's dense matrix is only defined for a few
            // thousand views, which one GPU holds.
            // ---- block-structured covariance (refine_kernels.cu, k_cov_*): shared block on the host, view blocks on the device ----
            CUDA_TRY(cudaMemsetAsync(V.fail, 0, sizeof(int32_t), h.st));
            launch_schur(S, h.L, h.B, V, ns, std::numeric_limits<double>::infinity(), h.st); h.launches += 4;
            if (h.comm && !h.comm->allreduce_sum(V.C, (size_t)ns * ns, h.st)) return fail(CAL_ERR_COMM, h.comm->error());
            int32_t failed = 0;
            CUDA_TRY(cudaMemcpyAsync(Cs.data(), V.C, sizeof(double) * ns * ns, cudaMemcpyDeviceToHost, h.st));
            CUDA_TRY(cudaMemcpyAsync(&failed, V.fail, sizeof failed, cudaMemcpyDeviceToHost, h.st));
            CUDA_TRY(cudaStreamSynchronize(h.st));
            if (h.comm) {
                if (!h.comm->check_timeout()) return fail(CAL_ERR_COMM, h.comm->error());
                double f = failed ? 1.0 : 0.0;
                if (!h.comm->allreduce_host(&f, 1, true)) return fail(CAL_ERR_COMM, h.comm->error());
                failed = f != 0.0;
            }
            if (failed) { if (trace) std::fprintf(stderr, "[calib_b200] covariance: a view block is rank deficient\n"); return CAL_OK; }  // covariance stays empty (ceresutils.h:86-88)
            for (int i = 0; i < ns; ++i) for (int j = 0; j < ns; ++j) Sm[(size_t)i * ns + j] = h.Hss[(size_t)i * ns + j] * s[i] * s[j] - Cs[(size_t)i * ns + j];
            if (ns > 0 && !chol_host(Sm, ns)) { if (trace) std::fprintf(stderr, "[calib_b200] covariance: the reduced shared block is rank deficient\n"); return CAL_OK; }
            std::vector<double> W((size_t)ns * ns), e(std::max(ns, 1));
            for (int j = 0; j < ns; ++j) { std::fill(e.begin(), e.end(), 0.0); e[j] = 1.0; chol_solve_host(Sm, ns, e.data()); for (int i = 0; i < ns; ++i) W[(size_t)i * ns + j] = e[i]; }
            double *dW = nullptr, *dZ = nullptr, *dG = nullptr, *dAinv = nullptr, *dcov = nullptr;
            // stream-ordered allocations from the device's pool (kept warm by cal_refine_create): no cudaMalloc / cudaFree on this path
            struct Free { std::vector<double**> p; cudaStream_t st; ~Free() { for (auto q : p) if (*q) cudaFreeAsync(*q, st); } } guard{{&dW, &dZ, &dG, &dAinv, &dcov}, h.st};
            auto take = [&](double** p, size_t n) { return cudaMallocAsync(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(double), h.st); };
            CUDA_TRY(take(&dW, (size_t)ns * ns)); CUDA_TRY(take(&dZ, (size_t)nv * 6 * ns)); CUDA_TRY(take(&dG, (size_t)nv * 6 * ns));
            CUDA_TRY(take(&dAinv, (size_t)nv * 36)); CUDA_TRY(take(&dcov, (size_t)na * na));
            CUDA_TRY(cudaMemcpyAsync(dW, W.data(), sizeof(double) * ns * ns, cudaMemcpyHostToDevice, h.st));
            CUDA_TRY(cudaMemsetAsync(dcov, 0, sizeof(double) * na * na, h.st));
            CUDA_TRY(cudaMemsetAsync(dG, 0, sizeof(double) * nv * 6 * ns, h.st));
            launch_cov_views(S, h.L, V, h.B.x, ns, dW, dZ, dG, dAinv, dcov, na, h.st); h.launches += 2;
            std::vector<double> Gh((size_t)nv * 6 * ns), sph((size_t)nv * 6);
            CUDA_TRY(download_large(cov, dcov, (size_t)na * na, h.st));
            CUDA_TRY(cudaMemcpyAsync(Gh.data(), dG, Gh.size() * sizeof(double), cudaMemcpyDeviceToHost, h.st));
            CUDA_TRY(cudaMemcpyAsync(sph.data(), V.sp, sph.size() * sizeof(double), cudaMemcpyDeviceToHost, h.st));
            CUDA_TRY(cudaStreamSynchronize(h.st));
            CUDA_TRY(cudaGetLastError());
            // plus-Jacobians of the shared blocks
            const int n_shared_pb = h.pb_viewq(0);
            std::vector<std::vector<double>> Pj(n_shared_pb);
            for (int bi = 0; bi < n_shared_pb; ++bi) {
                const PB& pb = h.pbs[bi]; if (pb.constant) continue;
                std::vector<double>& Pm = Pj[bi]; Pm.assign((size_t)pb.size * pb.tsize, 0.0);
                if (pb.type == PB_QUAT) quat_plus_jacobian(xf.data() + pb.off, Pm.data());
                else if (pb.type == PB_INTR && pb.tsize == pb.size - 1) { int k = 0; for (int j = 0; j < pb.size; ++j) { if (j == 4) continue; Pm[(size_t)j * pb.tsize + k] = 1.0; ++k; } }
                else for (int j = 0; j < pb.size; ++j) Pm[(size_t)j * pb.tsize + j] = 1.0;
            }
            for (int bi = 0; bi < n_shared_pb; ++bi) {
                const PB& a = h.pbs[bi]; if (a.constant) continue;
                // shared x shared: P_a (s s^T o W) P_b^T
                for (int bj = 0; bj < n_shared_pb; ++bj) {
                    const PB& b = h.pbs[bj]; if (b.constant) continue;
                    for (int i = 0; i < a.size; ++i) for (int j = 0; j < b.size; ++j) {
                        double acc = 0;
                        for (int k = 0; k < a.tsize; ++k) { const double pik = Pj[bi][(size_t)i * a.tsize + k]; if (pik == 0.0) continue;
                            for (int l = 0; l < b.tsize; ++l) acc += pik * s[a.toff + k] * s[b.toff + l] * W[(size_t)(a.toff + k) * ns + b.toff + l] * Pj[bj][(size_t)j * b.tsize + l]; }
                        cov[(size_t)(a.off + i) * na + b.off + j] = acc;
                    }
                }
                // shared x view: C_sv = -(Z_v W)^T, scaled and lifted on both sides
                for (int v = 0; v < nv; ++v) {
                    if (!h.view_free_host[v]) continue;
                    const double* q = xf.data() + S.off_viewq + 4 * (size_t)v;
                    double Pq[12]; quat_plus_jacobian(q, Pq);
                    double LA[12][6];
                    for (int i = 0; i < a.size; ++i) for (int m = 0; m < 6; ++m) {
                        double acc = 0;
                        for (int k = 0; k < a.tsize; ++k) acc -= Pj[bi][(size_t)i * a.tsize + k] * s[a.toff + k] * sph[(size_t)v * 6 + m] * Gh[((size_t)v * 6 + m) * ns + a.toff + k];
                        LA[i][m] = acc;
                    }
                    for (int i = 0; i < a.size; ++i) {
                        for (int bq = 0; bq < 4; ++bq) {
                            const double val = LA[i][0] * Pq[3 * bq] + LA[i][1] * Pq[3 * bq + 1] + LA[i][2] * Pq[3 * bq + 2];
                            cov[(size_t)(a.off + i) * na + S.off_viewq + 4 * (size_t)v + bq] = val;
                            cov[(size_t)(S.off_viewq + 4 * (size_t)v + bq) * na + a.off + i] = val;
                        }
                        for (int bt = 0; bt < 3; ++bt) {
                            cov[(size_t)(a.off + i) * na + S.off_viewt + 3 * (size_t)v + bt] = LA[i][3 + bt];
                            cov[(size_t)(S.off_viewt + 3 * (size_t)v + bt) * na + a.off + i] = LA[i][3 + bt];
                        }
                    }
                }
            }
            res->covariance_ok = 1;
            if (trace) std::fprintf(stderr, "[calib_b200] solve: covariance (block-structured, %d x %d) %8.2f ms\n", na, na,
                                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_lm_end).count());
            return CAL_OK;
        }
        std::vector<double> Hd, gd;
        if (n > 8192) return CAL_OK;
        if (cal_status st = dense_system(h, Hd, gd)) return st;
        std::vector<double> Lc = Hd;
        if (n > 0 && !chol_host(Lc, n)) return CAL_OK;  // rank deficient: covariance stays empty (ceresutils.h:86-88)
        std::vector<double> Ct((size_t)n * n, 0.0), e(n);
        for (int j = 0; j < n; ++j) { std::fill(e.begin(), e.end(), 0.0); e[j] = 1.0; chol_solve_host(Lc, n, e.data()); for (int i = 0; i < n; ++i) Ct[(size_t)i * n + j] = e[i]; }
        // lift to ambient coordinates block by block: cov(I,J) = P_I C(I,J) P_J^T
        std::fill(cov, cov + (size_t)na * na, 0.0);
        std::vector<std::vector<double>> Pj(h.pbs.size());
        for (size_t bi = 0; bi < h.pbs.size(); ++bi) {
            const PB& pb = h.pbs[bi]; if (pb.constant) continue;
            std::vector<double>& Pm = Pj[bi]; Pm.assign((size_t)pb.size * pb.tsize, 0.0);
            if (pb.type == PB_QUAT) quat_plus_jacobian(xf.data() + pb.off, Pm.data());
            else if (pb.type == PB_INTR && pb.tsize == pb.size - 1) { int k = 0; for (int j = 0; j < pb.size; ++j) { if (j == 4) continue; Pm[(size_t)j * pb.tsize + k] = 1.0; ++k; } }
            else for (int j = 0; j < pb.size; ++j) Pm[(size_t)j * pb.tsize + j] = 1.0;
        }
        for (size_t bi = 0; bi < h.pbs.size(); ++bi) {
            const PB& a = h.pbs[bi]; if (a.constant) continue;
            for (size_t bj = 0; bj < h.pbs.size(); ++bj) {
                const PB& b = h.pbs[bj]; if (b.constant) continue;
                for (int i = 0; i < a.size; ++i) for (int j = 0; j < b.size; ++j) {
                    double acc = 0;
                    for (int k = 0; k < a.tsize; ++k) { const double pik = Pj[bi][(size_t)i * a.tsize + k]; if (pik == 0.0) continue;
                        for (int l = 0; l < b.tsize; ++l) acc += pik * Ct[(size_t)(a.toff + k) * n + b.toff + l] * Pj[bj][(size_t)j * b.tsize + l]; }
                    cov[(size_t)(a.off + i) * na + b.off + j] = acc;
                }
            }
        }
        res->covariance_ok = 1;
    }
    return CAL_OK;
}

extern "C" cal_status cal_comm_create(const uint8_t unique_id[128], int rank, int world_size, int device, cal_comm** out) {
    if (!unique_id || !out) return fail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (cal_device_count() <= device) return fail(CAL_ERR_CUDA, "no CUDA device: calib_b200 has no CPU fallback");
    CUDA_TRY(cudaSetDevice(device));
    std::string err;
    calcomm::Comm* c = calcomm::Comm::create(unique_id, rank, world_size, &err);
    if (!c) return fail(CAL_ERR_COMM, err);
    *out = new cal_comm{c};
    return CAL_OK;
}
extern "C" void cal_comm_destroy(cal_comm* c) { if (c) { delete c->c; delete c; } }
extern "C" cal_status cal_refine_attach_comm(cal_refine_handle* h, cal_comm* c) {
    if (!h) return fail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    h->comm = c ? c->c : nullptr;
    return CAL_OK;
}
extern "C" cal_status cal_comm_peer_export(cal_comm* c, uint8_t handle_out[64]) {
    if (!c || !c->c || !handle_out) return fail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    return c->c->peer_export(handle_out) ? CAL_OK : fail(CAL_ERR_COMM, c->c->error());
}
extern "C" cal_status cal_comm_peer_enable(cal_comm* c, const uint8_t* handles) {
    if (!c || !c->c || !handles) return fail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    return c->c->peer_enable(handles) ? CAL_OK : fail(CAL_ERR_COMM, c->c->error());
}
extern "C" void cal_comm_peer_disable(cal_comm* c) { if (c && c->c) c->c->peer_disable(); }
extern "C" cal_status cal_comm_allreduce_test(cal_comm* c, double* host_inout, int32_t n, int use_peer) {
    if (!c || !c->c || !host_inout || n <= 0) return fail(CAL_ERR_INVALID_ARGUMENT, "bad argument");
    return c->c->allreduce_test(host_inout, (size_t)n, use_peer != 0) ? CAL_OK : fail(CAL_ERR_COMM, c->c->error());
}
extern "C" cal_status cal_comm_unique_id(uint8_t out128[128]) {
    std::string err;
    if (!calcomm::Comm::unique_id(out128, &err)) return fail(CAL_ERR_COMM, err);
    return CAL_OK;
}
