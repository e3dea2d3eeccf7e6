// How much of the FP64 pipe does an ILP-rich DFMA stream reach with W warps per SM sub-partition?
// (K1 keeps 68 accumulators per thread, which leaves room for 2 warps per sub-partition: DESIGN.md §4.)
// Each thread runs NACC independent accumulators; per group of 6 DFMAs the MIX variant adds one 64-bit
// shared-memory load, the instruction mix of K1's step loop.  W is forced through the dynamic shared-memory size.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_fp64_occ tools/ubench_fp64_occ.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC, bool MIX>
__global__ void __launch_bounds__(128) k_stream(double* out, int iters) {
    extern __shared__ double sm[];
    double a[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = threadIdx.x * 1e-9 + i;
    double m = 1.0000001, c = 1e-7;
    sm[threadIdx.x] = m; sm[threadIdx.x + 128] = c;
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MIX && i % 6 == 0) m = sm[(threadIdx.x + i + it) & 255];
            a[i] = fma(a[i], m, c);
        }
    }
    double r = 0; for (int i = 0; i < NACC; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int NACC, bool MIX>
void run(int W, int sms, double* out) {
    const int iters = 4096;
    const int smem = (227 * 1024) / W - 2048;   // W CTAs of 4 warps per SM = W warps per sub-partition
    cudaFuncSetAttribute(k_stream<NACC, MIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_stream<NACC, MIX>, 128, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * W * 4;
    k_stream<NACC, MIX><<<blocks, 128, smem>>>(out, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0); k_stream<NACC, MIX><<<blocks, 128, smem>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * NACC * iters * (double)blocks * 128 / (ms * 1e-3) / 1e12;
    printf("NACC %2d mix %d  warps/SMSP %d (occupancy %d CTAs)  %.3f ms  %.2f TFLOP/s  %.1f%% of 64 DFMA/clk/SM at 1.965 GHz\n", NACC, (int)MIX, W, occ, ms, tf,
           100.0 * tf / (148 * 1.965e9 * 128 / 1e12));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 8 * 4 * 128);
    for (int W : {1, 2, 3, 4, 8}) { run<64, false>(W, p.multiProcessorCount, out); run<64, true>(W, p.multiProcessorCount, out); }
    for (int W : {1, 2, 4}) run<16, false>(W, p.multiProcessorCount, out);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
