// Are the FP64 tensor-core instructions (DMMA: mma.sync m8n8k4 / m16n8k16, f64) a second FP64 roof on B200?
// K1 and the Schur SYRK are bound by the DFMA pipe; before shaping either around DMMA this measures
//   (1) DMMA throughput alone (independent accumulator tiles, 1..8 warps per sub-partition),
//   (2) DFMA throughput alone with the same launch shape,
//   (3) both in the same warp (interleaved) and in different warps of a sub-partition: does the sum exceed either roof?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_dmma tools/ubench_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5, %6, %7, %8, %9, %10, %11}, {%12, %13, %14, %15}, {%0, %1, %2, %3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// MODE 0: DMMA m8n8k4 only; 1: DFMA only; 2: both interleaved in every warp; 3: even warps DMMA, odd warps DFMA; 4: DMMA m16n8k16 only
template <int MODE>
__global__ void __launch_bounds__(256) k_run(double* out, int iters) {
    constexpr int NT = 8;    // independent accumulator tiles (DMMA) per warp
    constexpr int NA = 16;   // independent DFMA accumulators per thread
    double c0[NT], c1[NT], acc[NA], c4[NT / 2][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) { c0[i] = threadIdx.x * 1e-9 + i; c1[i] = 0.5 * i; }
#pragma unroll
    for (int i = 0; i < NT / 2; ++i) for (int k = 0; k < 4; ++k) c4[i][k] = i + k * 0.25;
#pragma unroll
    for (int i = 0; i < NA; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * (threadIdx.x & 3), m = 1.0000001, c = 1e-7;
    double a8[8], b4[4];
#pragma unroll
    for (int k = 0; k < 8; ++k) a8[k] = a + k * 1e-9;
#pragma unroll
    for (int k = 0; k < 4; ++k) b4[k] = b + k * 1e-9;
    const bool warp_dmma = MODE == 0 || MODE == 2 || MODE == 4 || (MODE == 3 && ((threadIdx.x >> 5) & 1) == 0);
    const bool warp_dfma = MODE == 1 || MODE == 2 || (MODE == 3 && ((threadIdx.x >> 5) & 1) == 1);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < NT / 2; ++i) dmma16816(c4[i], a8, b4);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                dmma884(c0[i], c1[i], a, b);
                acc[2 * i] = fma(acc[2 * i], m, c);
                acc[2 * i + 1] = fma(acc[2 * i + 1], m, c);
            }
        } else {
            if (warp_dmma) {
#pragma unroll
                for (int i = 0; i < NT; ++i) dmma884(c0[i], c1[i], a, b);
            }
            if (warp_dfma) {
#pragma unroll
                for (int i = 0; i < NA; ++i) acc[i] = fma(acc[i], m, c);
            }
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) r += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < NT / 2; ++i) for (int k = 0; k < 4; ++k) r += c4[i][k];
#pragma unroll
    for (int i = 0; i < NA; ++i) r += acc[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, int sms, double* out, int warps_per_sm) {
    const int iters = 8192, NT = 8, NA = 16;
    const int threads = warps_per_sm >= 8 ? 256 : warps_per_sm * 32, blocks = sms * (warps_per_sm >= 8 ? warps_per_sm / 8 : 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_run<MODE><<<blocks, threads>>>(out, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0); k_run<MODE><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)blocks * threads / 32;
    double mma_warps = MODE == 0 || MODE == 2 || MODE == 4 ? warps : (MODE == 3 ? warps / 2 : 0), fma_warps = MODE == 1 || MODE == 2 ? warps : (MODE == 3 ? warps / 2 : 0);
    if (MODE == 3 && threads == 32) { mma_warps = warps; fma_warps = 0; }
    const double mma_flop = MODE == 4 ? mma_warps * iters * (NT / 2) * 2.0 * 16 * 8 * 16 : mma_warps * iters * NT * 2.0 * 8 * 8 * 4;
    const double fma_flop = fma_warps * iters * (MODE == 2 ? 2 * NT : NA) * 2.0 * 32;
    printf("%-34s %2d warps/SM  %8.3f ms   DMMA %6.2f TFLOP/s   DFMA %6.2f TFLOP/s   sum %6.2f\n", name, warps_per_sm, ms, mma_flop / (ms * 1e-3) / 1e12,
           fma_flop / (ms * 1e-3) / 1e12, (mma_flop + fma_flop) / (ms * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 8 * 256);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    for (int w : {4, 8, 16, 32}) {
        run<0>("DMMA m8n8k4 alone", p.multiProcessorCount, out, w);
        run<4>("DMMA m16n8k16 alone", p.multiProcessorCount, out, w);
        run<1>("DFMA alone", p.multiProcessorCount, out, w);
        run<2>("DMMA + 2 DFMA interleaved per warp", p.multiProcessorCount, out, w);
        run<3>("even warps DMMA, odd warps DFMA", p.multiProcessorCount, out, w);
    }
    return 0;
}
