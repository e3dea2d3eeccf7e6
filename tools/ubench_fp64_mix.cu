// What does one memory / exchange instruction cost next to a DFMA stream on B200?  (Follow-up of ubench_fp64_occ.cu:
// one LDS.64 per 6 DFMAs took the FP64 pipe from 91 % to 74 % at ANY occupancy.)  64 accumulators per thread, 2 warps per
// sub-partition; per group of G DFMAs one instruction of kind K.  Reports DFMA throughput and the cost of the extra
// instruction in DFMA issue slots.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_fp64_mix tools/ubench_fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

enum { NONE, LDS32, LDS64, LDS128, STS64, SHFL, LDS64_BCAST, LDG64, LDC64, IMAD, FFMA, LDS64_NODEP, XCHG64, SHFL64, SEL64, SEL64C, LOP2, SHFL64ACC };
__constant__ double cmem[256];

template <int K, int G>
__global__ void __launch_bounds__(128) k_mix(double* out, const double* __restrict__ gin, int iters, int zero) {
    extern __shared__ double sm[];
    constexpr int NACC = 60;
    double a[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = threadIdx.x * 1e-9 + i;
    double m = 1.0000001, c = 1e-7, sink = 0, sink2 = 0;
    float f = threadIdx.x; int n = threadIdx.x;
    for (int i = threadIdx.x; i < 1024; i += 128) sm[i] = m;
    __syncthreads();
    const float* smf = reinterpret_cast<const float*>(sm);
    const double2* sm2 = reinterpret_cast<const double2*>(sm);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (i % G == 0) {
                const int j = (threadIdx.x + i + it * zero) & 255;
                if (K == LDS32) m = smf[j];
                if (K == LDS64) m = sm[j];
                if (K == LDS64_NODEP) sink += sm[j];
                if (K == LDS128) { double2 t = sm2[j]; m = t.x; c = t.y; }
                if (K == STS64) sm[256 + j] = a[i];
                if (K == SHFL) m = __shfl_xor_sync(0xffffffffu, m, 1 + (i & 15));
                if (K == LDS64_BCAST) m = sm[(i + it * zero) & 255];
                if (K == LDG64) m = gin[j];
                if (K == LDC64) m = cmem[(i + it * zero) & 255];
                if (K == IMAD) n = n * 3 + i;
                if (K == FFMA) f = fmaf(f, 1.0001f, 0.5f);
                if (K == SHFL64) { const double rcv = __shfl_xor_sync(0xffffffffu, m + 0.0 * i, 16); n ^= __double2hiint(rcv) ^ __double2loint(rcv); }
                if (K == SHFL64ACC) { const double rcv = __shfl_xor_sync(0xffffffffu, a[i], 16); n ^= __double2hiint(rcv) ^ __double2loint(rcv); }
                if (K == SEL64) { const double t = (threadIdx.x & 16) ? a[i] : a[(i + 7) % NACC]; n ^= __double2hiint(t) ^ __double2loint(t); }
                if (K == SEL64C) { const double t = ((threadIdx.x + i) & 16) ? m : c; n ^= __double2hiint(t) ^ __double2loint(t); }
                if (K == LOP2) { n ^= (n >> 3) + i; n ^= (n << 5) + it; }
                if (K == XCHG64) {   // what the half-warp exchange of K1 does per value: conditional swap (2 SEL) + 64-bit shuffle (2 SHFL);
                                     // the result goes to an integer sink so that no FMA waits for it
                    const double t = (threadIdx.x & 16) ? a[i] : a[(i + 7) % NACC];
                    { const double rcv = __shfl_xor_sync(0xffffffffu, t, 16); n ^= __double2hiint(rcv) ^ __double2loint(rcv); }
                }
            }
            a[i] = fma(a[i], m, c);
        }
    }
    double r = sink + sink2 + f + n; for (int i = 0; i < NACC; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

static double base_ms = 0;
template <int K, int G>
void run(const char* name, int sms, double* out, const double* gin) {
    const int iters = 4096, W = 2, NACC = 60;
    const int smem = (227 * 1024) / W - 2048;
    cudaFuncSetAttribute(k_mix<K, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * W;
    k_mix<K, G><<<blocks, 128, smem>>>(out, gin, iters, 0);
    cudaDeviceSynchronize();
    cudaEventRecord(e0); k_mix<K, G><<<blocks, 128, smem>>>(out, gin, iters, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (K == NONE) base_ms = ms;
    const double tf = 2.0 * NACC * iters * (double)blocks * 128 / (ms * 1e-3) / 1e12;
    // extra time per inserted instruction, in units of one DFMA's time in the baseline
    const double per = (ms - base_ms) / base_ms * NACC / ((NACC + G - 1) / G);
    printf("%-14s 1 per %2d DFMA  %.3f ms  %.2f TFLOP/s  cost of one = %.2f DFMA slots\n", name, G, ms, tf, K == NONE ? 0.0 : per);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out, *gin; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 8 * 128); cudaMalloc(&gin, 4096);
    cudaMemset(gin, 0, 4096);
    const int s = p.multiProcessorCount;
    run<NONE, 6>("none", s, out, gin);
    run<LDS32, 6>("LDS.32", s, out, gin); run<LDS64, 6>("LDS.64", s, out, gin); run<LDS64, 3>("LDS.64", s, out, gin); run<LDS64, 12>("LDS.64", s, out, gin);
    run<LDS64_NODEP, 6>("LDS.64 nodep", s, out, gin);
    run<LDS128, 6>("LDS.128", s, out, gin); run<LDS128, 12>("LDS.128", s, out, gin);
    run<STS64, 6>("STS.64", s, out, gin); run<SHFL, 6>("SHFL.32", s, out, gin); run<LDS64_BCAST, 6>("LDS.64 bcast", s, out, gin);
    run<LDG64, 6>("LDG.64 (L1)", s, out, gin); run<LDC64, 6>("LDC.64", s, out, gin); run<IMAD, 6>("IMAD", s, out, gin); run<SHFL64, 6>("SHFL64 inv", s, out, gin); run<SHFL64ACC, 6>("SHFL64 acc", s, out, gin); run<SEL64, 6>("SEL64 acc", s, out, gin); run<SEL64C, 6>("SEL64 inv", s, out, gin); run<LOP2, 6>("LOP x2", s, out, gin); run<XCHG64, 12>("SEL+SHFL 64b", s, out, gin); run<XCHG64, 6>("SEL+SHFL 64b", s, out, gin); run<XCHG64, 3>("SEL+SHFL 64b", s, out, gin); run<SHFL, 3>("SHFL.32", s, out, gin); run<FFMA, 6>("FFMA", s, out, gin);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
