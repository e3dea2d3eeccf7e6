// Mini-K1: the step loop of the fused pass reduced to its resource pattern, to compare two ways of sharing the Jacobian
// rows between the two threads that split a residual block's 136-entry system:
//   MODE 0  two warps per tile, rows exchanged through shared memory (STS + named barrier + LDS) — the shipped K1
//   MODE 1  two half-warps per 16 blocks, rows exchanged with a conditional swap + 64-bit shuffles (no LSU traffic)
// Per step and thread: ~95 dependent FP64 ops ("projection" of one corner, 26 outputs), then NF FMAs per corner on NA
// accumulators for the own and the partner's corner.  Prints ns per step and the DFMA-equivalent pipe utilisation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_k1_exchange tools/ubench_k1_exchange.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NV = 26;

__device__ __forceinline__ void project(double x, double y, const double* __restrict__ c, double* __restrict__ J) {
    // a dependent chain of about 95 FP64 instructions with 26 outputs (shape of obs_rows: depth, distortion, chain rule)
    double p = fma(c[0], x, fma(c[1], y, c[2])), q = fma(c[3], x, fma(c[4], y, c[5])), z = fma(c[6], x, fma(c[7], y, c[8]));
    double iz = 1.0 / z;   // division sequence: ~10 dependent ops
    double mx = p * iz, my = q * iz, xx = mx * mx, yy = my * my, xy = mx * my, r2 = xx + yy, r4 = r2 * r2, r6 = r4 * r2;
    double rad = fma(c[9], r6, fma(c[10], r4, fma(c[11], r2, 1.0)));
    double a1 = 2 * xy, a2 = fma(2.0, xx, r2), a3 = fma(2.0, yy, r2);
    double xd = fma(mx, rad, fma(c[12], a1, c[13] * a2)), yd = fma(my, rad, fma(c[12], a3, c[13] * a1));
    double dr = fma(3 * c[9], r4, fma(2 * c[10], r2, c[11]));
    double dxx = fma(2 * xx, dr, rad) + fma(2 * c[12], my, 6 * c[13] * mx), dxy = fma(a1, dr, 2 * fma(c[12], mx, c[13] * my));
    double dyy = fma(2 * yy, dr, rad) + fma(6 * c[12], my, 2 * c[13] * mx);
    double ux = fma(c[14], dxx, c[15] * dxy), uy = fma(c[14], dxy, c[15] * dyy), vx = c[16] * dxy, vy = c[16] * dyy;
    double cu = -fma(ux, mx, uy * my), cv = -fma(vx, mx, vy * my);
    J[0] = fma(my, cu, -uy); J[1] = fma(-mx, cu, ux); J[2] = fma(mx, uy, -my * ux);
    J[3] = fma(my, cv, -vy); J[4] = fma(-mx, cv, vx); J[5] = fma(mx, vy, -my * vx);
    J[6] = ux * iz; J[7] = uy * iz; J[8] = cu * iz; J[9] = vx * iz; J[10] = vy * iz; J[11] = cv * iz;
    double fxs = fma(c[14], mx, c[15] * my), fyy = c[16] * my;
    J[12] = fxs * r2; J[13] = fxs * r4; J[14] = fxs * r6; J[15] = fma(c[14], a1, c[15] * a3); J[16] = fma(c[14], a2, c[15] * a1);
    J[17] = fyy * r2; J[18] = fyy * r4; J[19] = fyy * r6; J[20] = c[16] * a3; J[21] = c[16] * a1;
    J[22] = xd; J[23] = yd; J[24] = fma(c[14], xd, c[17]) - x; J[25] = fma(c[16], yd, c[18]) - y;
}

template <int NA, int NF>
__device__ __forceinline__ void accumulate(const double* __restrict__ X, double* __restrict__ acc) {
#pragma unroll
    for (int k = 0; k < NF; ++k) acc[k % NA] = fma(X[k % NV], X[(k * 7 + 3) % NV], acc[k % NA]);
}

template <int MODE>
__global__ void __launch_bounds__(64) k_mini(const double* __restrict__ in, double* __restrict__ out, int steps) {
    constexpr int NA = MODE == 0 ? 68 : 69, NF = MODE == 0 ? 105 : 111;
    __shared__ double2 xb[2][2][NV / 2][32];
    const int lane = threadIdx.x & 31, role = MODE == 0 ? threadIdx.x >> 5 : lane >> 4;
    double c[19];
#pragma unroll
    for (int i = 0; i < 19; ++i) c[i] = in[i];
    double acc[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) acc[i] = 0.0;
    double x = in[32 + threadIdx.x], y = in[96 + threadIdx.x];
    for (int s = 0; s < steps; ++s) {
        double J[NV];
        project(x + 1e-9 * s, y - 1e-9 * s, c, J);
        if (MODE == 0) {
            double* slot = reinterpret_cast<double*>(&xb[s & 1][role][0][0]);
#pragma unroll
            for (int i = 0; i < NV; ++i) slot[i * 32 + lane] = J[i];
            accumulate<NA, NF>(J, acc);
            asm volatile("bar.sync 1, 64;" ::: "memory");
            const double* other = reinterpret_cast<const double*>(&xb[s & 1][role ^ 1][0][0]);
            double P[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) P[i] = other[i * 32 + lane];
            accumulate<NA, NF>(P, acc);
        } else {
            // conditional swap of the 13 value pairs (role 1 holds its row in the permuted arrangement), then the partner's
            // row arrives by 64-bit shuffles: O[i] is used, O[i ^ 1] of the partner is what this lane's slot i needs
            double O[NV], P[NV];
#pragma unroll
            for (int i = 0; i < NV; i += 2) {
                O[i] = role ? J[i + 1] : J[i];
                O[i + 1] = role ? J[i] : J[i + 1];
            }
#pragma unroll
            for (int i = 0; i < NV; ++i) P[i] = __shfl_xor_sync(0xffffffffu, O[i ^ 1], 16);
            accumulate<NA, NF>(O, acc);
            accumulate<NA, NF>(P, acc);
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < NA; ++i) r += acc[i];
    out[blockIdx.x * 64 + threadIdx.x] = r;
}

template <int MODE>
void run(int sms, const double* in, double* out) {
    const int steps = 4400, blocks = sms * 4 * 8;   // 4 resident CTAs of 2 warps per SM, 8 waves
    cudaFuncSetAttribute(k_mini<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_mini<MODE>);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mini<MODE>, 64, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mini<MODE><<<blocks, 64>>>(in, out, steps); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_mini<MODE><<<blocks, 64>>>(in, out, steps); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // corners processed: MODE 0: a CTA step = 2 corners of 32 blocks = 64 observations; MODE 1: 2 warps x 2 corners x 16 blocks = 64
    const double obs = (double)blocks * steps * 64;
    printf("mode %d  regs %d  CTAs/SM %d  %.3f ms  %.2f ps per observation (C5 has 70.4 M: %.3f ms)\n", MODE, fa.numRegs, occ, ms, ms * 1e9 / obs,
           ms * 70.4e6 / obs);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *in, *out; cudaMalloc(&in, 4096); cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 32 * 64);
    double h[512]; for (int i = 0; i < 512; ++i) h[i] = 0.3 + 0.001 * i;
    h[2] = 0.1; h[5] = 0.2; h[8] = 2.0; h[6] = h[7] = 0.01;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    run<0>(p.multiProcessorCount, in, out); run<1>(p.multiProcessorCount, in, out);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
