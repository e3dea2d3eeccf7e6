// TEST INFRASTRUCTURE — the PRODUCT's AX = XB kernels (calibration_b200/csrc/axxb_kernels.cuh: k_axxb_transpose,
// k_axxb, k_pair_mask, k_axxb_otf, k_axxb_final) compiled by g++ and run on the CPU under the lock-step SIMT shim,
// launched in the order and with the grids of cal_axxb_create / cal_axxb_create_from_poses / axxb_pass (axxb.cu).
#define SIMT_SHARED_STORAGE static
#include "simt_shim.hpp"

#include "../../include/calib_b200.h"
#include "../../calibration_b200/csrc/axxb_kernels.cuh"

namespace {
void unpack(const double* o, int jac, double* cost, double* g6, double* H36) {
    *cost = o[27];
    if (!jac) return;
    int k = 0;
    for (int a = 0; a < 6; ++a) for (int b = a; b < 6; ++b) { H36[6 * a + b] = o[k]; H36[6 * b + a] = o[k]; ++k; }
    for (int a = 0; a < 6; ++a) g6[a] = o[21 + a];
}
}  // namespace

// materialised pairs (cal_axxb_create + one axxb_pass)
extern "C" int simt_axxb_eval(int64_t n, const double* rot_a, const double* rot_b, const double* tra_a, const double* tra_b, double huber,
                              const double* x7, int jac, double* cost, double* g6, double* H36) {
    std::vector<double> pairs(24 * (size_t)n);
    const double* srcs[4] = {rot_a, rot_b, tra_a, tra_b};
    const int widths[4] = {9, 9, 3, 3}, rows[4] = {0, 9, 18, 21};
    for (int k = 0; k < 4; ++k) {
        const int64_t tot = n * widths[k];
        simt::launch((unsigned)((tot + 255) / 256), 256, [&] { k_axxb_transpose(srcs[k], widths[k], n, pairs.data(), rows[k]); });
    }
    const int n_cta = (int)std::min<int64_t>(148 * 8, (n + 255) / 256);
    std::vector<double> partial((size_t)kAcc * n_cta), out(kAcc);
    if (jac) simt::launch(n_cta, 256, [&] { k_axxb<1>(pairs.data(), n, x7, huber, partial.data()); });
    else simt::launch(n_cta, 256, [&] { k_axxb<0>(pairs.data(), n, x7, huber, partial.data()); });
    simt::launch(1, 32 * kAcc, [&] { k_axxb_final(partial.data(), n_cta, out.data()); });
    unpack(out.data(), jac, cost, g6, H36);
    return 0;
}

// pairs formed on the fly from the poses (cal_axxb_create_from_poses + one axxb_pass); returns the pairs kept
extern "C" int64_t simt_axxb_eval_from_poses(int64_t n_poses, const double* bg12, const double* ct12, double min_angle_deg,
                                             int reject_axis_parallel, double axis_parallel_eps, double huber, const double* x7, int jac,
                                             double* cost, double* g6, double* H36) {
    const int T = (int)((n_poses + kTile - 1) / kTile);
    const int64_t n_tiles = (int64_t)T * (T + 1) / 2;
    std::vector<unsigned> mask(32 * (size_t)n_tiles);
    PairTiles tiles{};
    tiles.n_poses = n_poses; tiles.n_tiles_1d = T; tiles.G = bg12; tiles.C = ct12; tiles.mask = mask.data();
    unsigned long long kept = 0;
    const double kPi = 3.14159265358979323846;
    simt::launch((unsigned)n_tiles, 256, [&] { k_pair_mask(tiles, min_angle_deg * kPi / 180.0, reject_axis_parallel, axis_parallel_eps, &kept); });
    if (kept == 0) return 0;
    std::vector<double> partial((size_t)kAcc * n_tiles), out(kAcc);
    if (jac) simt::launch((unsigned)n_tiles, 256, [&] { k_axxb_otf<1>(tiles, x7, huber, partial.data()); });
    else simt::launch((unsigned)n_tiles, 256, [&] { k_axxb_otf<0>(tiles, x7, huber, partial.data()); });
    simt::launch(1, 32 * kAcc, [&] { k_axxb_final(partial.data(), (int)n_tiles, out.data()); });
    unpack(out.data(), jac, cost, g6, H36);
    return (int64_t)kept;
}
