// TEST INFRASTRUCTURE — the PRODUCT's RANSAC kernels (calibration_b200/csrc/ransac_kernel.cuh: k_ransac;
// ransac_plane_kernel.cuh: k_ransac_plane; with dlt.cuh, ransac_sampler.cuh, plane_math.cuh) compiled by g++
// and run on the CPU under the lock-step SIMT shim: the same source the GPU runs, including the device-side
// std::sample replay, the ballot masks, the lane-private solves and the batch bookkeeping.  The launch
// parameters (4 warps per CTA, the shared-memory carve-up, the host-built iteration table) are taken from the
// product's own constants and ransac_iters.hpp.
#include "simt_shim.hpp"

namespace {
alignas(16) unsigned char smem_raw[232448];   // the 227 KB a CTA can have
}

#include "../../calibration_b200/csrc/ransac_kernel.cuh"
#include "../../calibration_b200/csrc/ransac_plane_kernel.cuh"

extern "C" int simt_ransac_homography(int64_t n_problems, int32_t n, const double* x, const double* y, const double* u, const double* v,
                                      const cal_ransac_options* o, int seed_per_problem, cal_ransac_result* res, uint8_t* mask) {
    const std::vector<int> table = build_niter_table(n, *o, 4);
    const int nw = (n + 31) / 32;
    const size_t per_warp = (size_t)4 * n * sizeof(double) + 312 * sizeof(unsigned long long) + (size_t)(1 + 2 * kBatch) * nw * sizeof(unsigned);
    if (kWarpsPerCta * ((per_warp + 15) / 16 * 16) > sizeof smem_raw) return 1;
    const unsigned grid = (unsigned)((n_problems + kWarpsPerCta - 1) / kWarpsPerCta);
    simt::launch(grid, 32 * kWarpsPerCta, [&] { k_ransac(n_problems, n, x, y, u, v, *o, seed_per_problem, table.data(), res, mask); });
    return 0;
}

extern "C" int simt_ransac_plane(int64_t n_problems, int32_t n, const double* x, const double* y, const double* z,
                                 const cal_ransac_options* o, int seed_per_problem, cal_plane_ransac_result* res, uint8_t* mask) {
    const std::vector<int> table = build_niter_table(n, *o, 3);
    if (kPlaneWarps * plane_per_warp_bytes(n) > sizeof smem_raw) return 1;
    const unsigned grid = (unsigned)((n_problems + kPlaneWarps - 1) / kPlaneWarps);
    simt::launch(grid, 32 * kPlaneWarps, [&] { k_ransac_plane(n_problems, n, x, y, z, *o, seed_per_problem, table.data(), res, mask); });
    return 0;
}
