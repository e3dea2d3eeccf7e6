// Microbenchmarks that decide the K1 design: does the FP64 pipe skip inactive
// half-warps, what does a 64-bit shuffle / shared-memory exchange cost next to a DFMA stream.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>  // 0 full warp, 1 lanes 0-15, 2 even lanes, 3 divergent halves (both execute, different code)
__global__ void k_dfma(double* out, int iters) {
    const int lane = threadIdx.x & 31;
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    bool act = MODE == 0 || (MODE == 1 && lane < 16) || (MODE == 2 && !(lane & 1)) || MODE == 3;
    if (act) {
        if (MODE == 3 && lane >= 16) {
            for (int i = 0; i < iters; ++i) {
                a0 = fma(a0, c, m); a1 = fma(a1, c, m); a2 = fma(a2, c, m); a3 = fma(a3, c, m);
                a4 = fma(a4, c, m); a5 = fma(a5, c, m); a6 = fma(a6, c, m); a7 = fma(a7, c, m);
            }
        } else {
            for (int i = 0; i < iters; ++i) {
                a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
                a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// DFMA stream with NS 64-bit shuffles (2 SHFL each) per 32 DFMAs
template <int NS>
__global__ void k_dfma_shfl(double* out, int iters) {
    double a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, c = 1e-7;
    double s = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
#pragma unroll
        for (int k = 0; k < NS; ++k) s += __shfl_xor_sync(0xffffffffu, a[k % 16], 16);
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, s);
    }
    double r = s; for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// DFMA stream with NS shared-memory round trips (STS.64 + LDS.64 of the partner warp's slot, named barrier) per 32 DFMAs
template <int NS>
__global__ void k_dfma_smem(double* out, int iters) {
    __shared__ double buf[2][8][NS > 8 ? 8 : (NS > 0 ? NS : 1)][32];
    double a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, c = 1e-7;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, pair = warp >> 1;
    double s = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
#pragma unroll
        for (int k = 0; k < NS; ++k) buf[it & 1][warp][k % 8][lane] = a[k % 16];
        asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
#pragma unroll
        for (int k = 0; k < NS; ++k) s += buf[it & 1][warp ^ 1][k % 8][lane];
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, s);
    }
    double r = s; for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <class F> float time_it(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 2, threads = 256, iters = 1 << 15;
    double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
    const double n = 8.0 * iters * (double)blocks * threads;
    float t;
    t = time_it([&] { k_dfma<0><<<blocks, threads>>>(out, iters); }); printf("dfma full        %.3f ms  %.2f T lane-DFMA/s (all lanes counted)\n", t, n / t / 1e9);
    t = time_it([&] { k_dfma<1><<<blocks, threads>>>(out, iters); }); printf("dfma lanes 0-15  %.3f ms  %.2f T warp-slot-equivalents/s\n", t, n / t / 1e9);
    t = time_it([&] { k_dfma<2><<<blocks, threads>>>(out, iters); }); printf("dfma even lanes  %.3f ms  %.2f\n", t, n / t / 1e9);
    t = time_it([&] { k_dfma<3><<<blocks, threads>>>(out, iters); }); printf("dfma divergent halves %.3f ms  %.2f\n", t, n / t / 1e9);
    const int it2 = 1 << 13; const double n2 = 32.0 * it2 * (double)blocks * threads;
    t = time_it([&] { k_dfma_shfl<0><<<blocks, threads>>>(out, it2); }); printf("dfma32 + 0 shfl64  %.3f ms %.2f T DFMA/s\n", t, n2 / t / 1e9);
    t = time_it([&] { k_dfma_shfl<4><<<blocks, threads>>>(out, it2); }); printf("dfma32 + 4 shfl64  %.3f ms %.2f\n", t, n2 / t / 1e9);
    t = time_it([&] { k_dfma_shfl<8><<<blocks, threads>>>(out, it2); }); printf("dfma32 + 8 shfl64  %.3f ms %.2f\n", t, n2 / t / 1e9);
    t = time_it([&] { k_dfma_shfl<16><<<blocks, threads>>>(out, it2); }); printf("dfma32 + 16 shfl64 %.3f ms %.2f\n", t, n2 / t / 1e9);
    t = time_it([&] { k_dfma_smem<0><<<blocks, threads>>>(out, it2); }); printf("dfma32 + bar only  %.3f ms %.2f\n", t, n2 / t / 1e9);
    t = time_it([&] { k_dfma_smem<4><<<blocks, threads>>>(out, it2); }); printf("dfma32 + 4 smem rt %.3f ms %.2f\n", t, n2 / t / 1e9);
    t = time_it([&] { k_dfma_smem<8><<<blocks, threads>>>(out, it2); }); printf("dfma32 + 8 smem rt %.3f ms %.2f\n", t, n2 / t / 1e9);
    t = time_it([&] { k_dfma_smem<16><<<blocks, threads>>>(out, it2); }); printf("dfma32 + 16 smem rt %.3f ms %.2f\n", t, n2 / t / 1e9);
    return 0;
}
