// The reference's minimal-sample stream, generated on the device: std::sample(iota(N), K,
// std::mt19937_64(seed)) of libstdc++ (include/calib/estimation/common/ransac.h:135,144-145).  One
// warp per problem; the engine state lives in shared memory (312 words), the twist and the tempering
// run lane-parallel, and the selection-sampling walk takes 32 (index, index+1) pairs per step with one
// Lemire multiply per lane.  Integer arithmetic only: bit-exact.  Shared by the homography kernel
// (ransac.cu, K = 4) and the plane kernel (ransac_plane.cu, K = 3).
#pragma once
#include "dlt.cuh"
#include "ransac_iters.hpp"

namespace {

// ---- std::mt19937_64, state in shared memory, lane-parallel twist ----
__device__ __forceinline__ unsigned long long mt_temper(unsigned long long z) {
    z ^= (z >> 29) & 0x5555555555555555ULL;
    z ^= (z << 17) & 0x71D67FFFEDA60000ULL;
    z ^= (z << 37) & 0xFFF7EEE000000000ULL;
    z ^= z >> 43;
    return z;
}
__device__ void mt_seed(unsigned long long* mt, unsigned long long seed, int lane) {
    unsigned long long xv = seed;
    if (lane == 0) mt[0] = xv;
    for (int i = 1; i < 312; ++i) {
        xv = 6364136223846793005ULL * (xv ^ (xv >> 62)) + (unsigned long long)i;
        if ((i & 31) == lane) mt[i] = xv;
    }
    __syncwarp();
}
__device__ void mt_twist(unsigned long long* mt, int lane) {
    const unsigned long long UP = 0xFFFFFFFF80000000ULL, LO = 0x7FFFFFFFULL, A = 0xB5026F5AA96619E9ULL;
    for (int base = 0; base < 312; base += 32) {
        const int i = base + lane;
        unsigned long long nv = 0;
        if (i < 312) {
            const unsigned long long yv = (mt[i] & UP) | (mt[(i + 1) % 312] & LO);
            nv = mt[(i + 156) % 312] ^ (yv >> 1) ^ ((yv & 1ULL) ? A : 0ULL);
        }
        __syncwarp();
        if (i < 312) mt[i] = nv;
        __syncwarp();
    }
}
// one engine output (warp-uniform)
__device__ __forceinline__ unsigned long long mt_next(unsigned long long* mt, int& pos, int lane) {
    if (pos >= 312) { mt_twist(mt, lane); pos = 0; }
    return mt_temper(mt[pos++]);
}
// uniform_int_distribution<unsigned long>{0, range-1}: Lemire (uniform_int_dist.h:252-281), warp-uniform
__device__ unsigned long long lemire_uniform(unsigned long long* mt, int& pos, int lane, unsigned long long range) {
    unsigned long long r = mt_next(mt, pos, lane);
    unsigned long long low = r * range, hi = __umul64hi(r, range);
    if (low < range) {
        const unsigned long long threshold = (0ULL - range) % range;
        while (low < threshold) { r = mt_next(mt, pos, lane); low = r * range; hi = __umul64hi(r, range); }
    }
    return hi;
}
// std::sample(0..N-1, K) — libstdc++ selection sampling (stl_algo.h:5841-5907).  Results are warp-uniform.
template <int K>
__device__ void sample_k(unsigned long long* mt, int& pos, int lane, int N, int* idx) {
    unsigned long long uns = (unsigned long long)N;
    int need = N < K ? N : K, first = 0, o = 0;
    bool slow = false;
    while (need != 0 && uns >= 2 && !slow) {
        if (pos >= 312) { mt_twist(mt, lane); pos = 0; }
        const int avail = 312 - pos;
        const unsigned long long pairs_left = uns / 2;
        int B = 32; if (avail < B) B = avail; if (pairs_left < (unsigned long long)B) B = (int)pairs_left;
        unsigned p0 = 0xffffffffu, p1 = 0xffffffffu; bool rej = false;
        if (lane < B) {
            const unsigned long long uj = uns - 2ULL * lane, b1 = uj - 1ULL, range = uj * b1;
            const unsigned long long r = mt_temper(mt[pos + lane]);
            const unsigned long long low = r * range, hi = __umul64hi(r, range);
            rej = low < range && low < (0ULL - range) % range;   // would redraw: leave the fast path
            if (N < 65536) { const unsigned h32 = (unsigned)hi, b32 = (unsigned)b1; p0 = h32 / b32; p1 = h32 - p0 * b32; }
            else { p0 = (unsigned)(hi / b1); p1 = (unsigned)(hi % b1); }
        }
        if (__any_sync(kFull, rej)) { slow = true; break; }
        unsigned m = __ballot_sync(kFull, lane < B && (p0 < (unsigned)need || p1 < (unsigned)need));
        int consumed = B; bool done = false;
        while (m) {
            const int j = __ffs(m) - 1; m &= m - 1;
            const unsigned a = __shfl_sync(kFull, p0, j), b = __shfl_sync(kFull, p1, j);
            if (a < (unsigned)need) { idx[o++] = first + 2 * j; --need; }
            if (need == 0) { consumed = j + 1; done = true; break; }
            if (b < (unsigned)need) { idx[o++] = first + 2 * j + 1; --need; }
            if (need == 0) { consumed = j + 1; done = true; break; }
        }
        pos += consumed;
        if (!done) { uns -= 2ULL * B; first += 2 * B; }
    }
    if (slow) {  // exact sequential replay from the current state (a Lemire redraw occurred)
        while (need != 0 && uns >= 2) {
            const unsigned long long b1 = uns - 1ULL;
            const unsigned long long xx = lemire_uniform(mt, pos, lane, uns * b1);
            const unsigned long long q0 = xx / b1, q1 = xx % b1;
            --uns;
            if (q0 < (unsigned long long)need) { idx[o++] = first; --need; }
            ++first;
            if (need == 0) break;
            --uns;
            if (q1 < (unsigned long long)need) { idx[o++] = first; --need; }
            ++first;
        }
    }
    for (; need != 0; ++first) {  // one-at-a-time tail (stl_algo.h:5899-5905)
        --uns;
        if (lemire_uniform(mt, pos, lane, uns + 1ULL) < (unsigned long long)need) { idx[o++] = first; --need; }
    }
}

}  // namespace
