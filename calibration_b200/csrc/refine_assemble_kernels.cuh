// Device code of the segment-layout assembly (small problems: blocks cut into short segments, K1 stores the local
// system per segment): per-block Huber weight, the chain rule of the view-type pose block, and the deterministic
// per-camera column sums (see refine_kernels.cu for the description and the launchers).  Kept in a header so that
// tests/host_emul can run this very source on the CPU under the lock-step SIMT shim (test-only).
#pragma once
#include "k1_math.cuh"
#include "refine_kernels.cuh"

namespace calk {

template <int JAC>
__global__ void k_block_weight(ProblemShape S, DevLayout L, EvalBuffers B, int rr_row) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= L.n_blk) return;
    const int s0 = L.blk_seg_off[b], s1 = L.blk_seg_off[b + 1];
    double ssr = 0.0;
    for (int s = s0; s < s1; ++s) ssr += JAC ? B.segN[(int64_t)rr_row * L.n_seg + s] : B.seg_ssr[s];
    double rho, w; huber_weight(S.huber_delta, ssr, rho, w);
    B.blk_ssr[b] = ssr;
    B.blk_w[b] = w;
    B.blk_rows[b] = 0.5 * rho;  // row 0 of the block-indexed matrix: cost
    if (JAC) for (int s = s0; s < s1; ++s) B.seg_w[s] = w;
}

// ONE = every block of the launch has exactly one segment (the large-problem case): all
// input loads are then independent straight-line loads the compiler batches (memory-level
// parallelism); otherwise the per-block segment loop is kept.
template <int MODEL, int IMODE, bool ONE>
__device__ __forceinline__ void view_part_body(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, int64_t b) {
    using LT = Local<MODEL, IMODE>;
    constexpr int PI = LT::PI, NC = LT::NC;
    const int s0 = L.blk_seg_off[b], s1 = ONE ? s0 + 1 : L.blk_seg_off[b + 1];
    const double w = B.blk_w[b];
    const bool bundle = S.kind == 2;
    const double* __restrict__ segN = B.segN;
    auto load = [&](int a, int c) {
        const double* __restrict__ p = segN + (int64_t)LT::idx(a, c) * L.n_seg;
        double v = p[s0];
        if (!ONE) for (int s = s0 + 1; s < s1; ++s) v += p[s];
        return v;
    };
    double T[36], N[21], nr[6], ni[PI > 0 ? 6 * PI : 1];
#pragma unroll
    for (int i = 0; i < 36; ++i) T[i] = B.blk_Tv[(int64_t)i * L.n_blk + b];
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int c = a; c < 6; ++c) N[a * 6 - a * (a - 1) / 2 + (c - a)] = load(a, c);
#pragma unroll
    for (int k = 0; k < 6; ++k) nr[k] = load(k, NC);
#pragma unroll
    for (int j = 0; j < PI; ++j)
#pragma unroll
        for (int k = 0; k < 6; ++k) ni[6 * j + k] = load(k, 6 + j);
    double Q[36];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) a = fma(T[6 * k + i], N[k <= j ? k * 6 - k * (k - 1) / 2 + (j - k) : j * 6 - j * (j - 1) / 2 + (k - j)], a);
            Q[6 * i + j] = a * w;
        }
    // rows of the block-indexed matrix: 0 cost | 1..21 Hvv | 22..27 gv | 28..63 Q or Evc | 64.. Evi
    double* __restrict__ rows = B.blk_rows;
    const int64_t nb = L.n_blk;
    {
        int o = 1;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) {
                double a = 0.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) a = fma(Q[6 * i + k], T[6 * k + j], a);
                if (bundle) rows[(int64_t)o * nb + b] = a; else B.blk_Hvv[(int64_t)(o - 1) * nb + b] = a;
                ++o;
            }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) a = fma(T[6 * k + i], nr[k], a);
        a *= w;
        if (bundle) rows[(int64_t)(22 + i) * nb + b] = a; else B.blk_gv[(int64_t)i * nb + b] = a;
    }
    if (bundle) {
#pragma unroll
        for (int i = 0; i < 36; ++i) rows[(int64_t)(28 + i) * nb + b] = Q[i];
    } else {
        const double* __restrict__ Tc = B.camT + (int64_t)L.blk_cam[b] * 36;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                double a = 0.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) a = fma(Q[6 * i + k], Tc[6 * k + j], a);
                B.blk_Evc[(int64_t)(6 * i + j) * nb + b] = a;
            }
    }
#pragma unroll
    for (int j = 0; j < PI; ++j)
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) a = fma(T[6 * k + i], ni[6 * j + k], a);
            a *= w;
            if (bundle) rows[(int64_t)(64 + PI * i + j) * nb + b] = a; else B.blk_Evi[(int64_t)(PI * i + j) * nb + b] = a;
        }
}

template <int MODEL, int IMODE, bool ONE>
__global__ void __launch_bounds__(128) k_view_part(ProblemShape S, DevLayout L, EvalBuffers B) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= L.n_blk) return;
    if (!L.blk_vfree[b]) return;  // outputs stay zero (buffers are cleared once at create)
    view_part_body<MODEL, IMODE, ONE>(S, L, B, b);
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(256) k_colsum(const double* __restrict__ M, int64_t ld, const double* __restrict__ w,
                                                const ColChunk* __restrict__ chunks, double* __restrict__ partial,
                                                int n_rows) {
    __shared__ double sm[8];
    const ColChunk c = chunks[blockIdx.x];
    const int row = blockIdx.y;
    const double* __restrict__ p = M + (int64_t)row * ld;
    double a = 0.0;
    for (int64_t i = c.begin + threadIdx.x; i < c.end; i += 256) a += WEIGHTED ? p[i] * w[i] : p[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sm[k];
        partial[(int64_t)blockIdx.x * n_rows + row] = t;
    }
}

__global__ void k_final_reduce(const double* __restrict__ partial, const int32_t* __restrict__ cam_chunk_off, int n_cams,
                               int n_rows, double* __restrict__ cam_sums, int NV, int row_base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cams * n_rows) return;
    const int cam = i / n_rows, row = i % n_rows;
    double a = 0.0;
    for (int c = cam_chunk_off[cam]; c < cam_chunk_off[cam + 1]; ++c) a += partial[(int64_t)c * n_rows + row];
    cam_sums[(int64_t)cam * NV + row_base + row] = a;
}

}  // namespace calk
