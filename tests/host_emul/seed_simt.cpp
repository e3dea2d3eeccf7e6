// TEST INFRASTRUCTURE — the PRODUCT's seeding kernels (calibration_b200/csrc/seed_kernels.cuh: k_view_dlt,
// k_zhang_gram, k_pose_from_h, k_seed_from_ransac, k_gather_views, k_scatter_views) and its host-side Zhang
// closed form (seed_host.hpp) compiled by g++ and run on the CPU under the lock-step SIMT shim, launched in the
// order and with the grids of cal_seed_intrinsics / cal_seed_planar_poses (seed.cu).
#define SIMT_SHARED_STORAGE static
#include "simt_shim.hpp"

#include "../../calibration_b200/csrc/seed_host.hpp"
#include "../../calibration_b200/csrc/seed_kernels.cuh"

extern "C" int simt_seed_intrinsics(int64_t n_views, const int64_t* off, const int32_t* cam, const double* x, const double* y,
                                    const double* u, const double* v, int32_t n_cams, const cal_seed_options* opts, double* kmtx,
                                    int32_t* cam_success, int32_t* view_success, double* hmtx, double* sym_rms, double* poses) {
    std::vector<double> rows(12 * (size_t)n_views), gram(22 * (size_t)n_cams);
    SeedArgs a{n_views, off, cam, x, y, u, v, nullptr, hmtx, sym_rms, rows.data(), view_success, nullptr};
    const unsigned groups = (unsigned)((n_views + 31) / 32);
    simt::launch((groups + 3) / 4, 128, [&] { k_view_dlt<false>(a); });
    simt::launch((unsigned)n_cams, 256, [&] { k_zhang_gram(n_views, cam, view_success, rows.data(), gram.data()); });
    std::vector<int32_t> camok(n_cams, 0);
    for (int c = 0; c < n_cams; ++c) {   // the host step of cal_seed_intrinsics_ransac (seed.cu)
        double* k5 = kmtx + 5 * c;
        for (int i = 0; i < 5; ++i) k5[i] = 0.0;
        const double* g = &gram[22 * (size_t)c];
        if (g[21] < 4.0) continue;
        double G[6][6], b[6];
        { int o = 0; for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { G[i][j] = G[j][i] = g[o]; ++o; } }
        smallest_eigvec6(G, b);
        if (!kmtx_from_conic(b, k5)) { for (int i = 0; i < 5; ++i) k5[i] = 0.0; continue; }
        if (opts && opts->use_bounds) sanitize(k5, *opts);
        camok[c] = 1;
    }
    for (int c = 0; c < n_cams; ++c) cam_success[c] = camok[c];
    simt::launch((unsigned)((n_views + 127) / 128), 128, [&] { k_pose_from_h(n_views, cam, view_success, camok.data(), kmtx, hmtx, poses); });
    return 0;
}

extern "C" int simt_seed_planar_poses(int64_t n_views, const int64_t* off, const int32_t* cam, const double* x, const double* y,
                                      const double* u, const double* v, const double* kmtx, double* poses, int32_t* view_success) {
    SeedArgs a{n_views, off, cam, x, y, u, v, kmtx, nullptr, nullptr, nullptr, view_success, poses};
    const unsigned groups = (unsigned)((n_views + 31) / 32);
    simt::launch((groups + 3) / 4, 128, [&] { k_view_dlt<true>(a); });
    return 0;
}

// ragged views through the equal-size RANSAC kernel: the gather / scatter pair of cal_seed_intrinsics_ransac
extern "C" int simt_gather_scatter_roundtrip(int64_t cnt, int32_t n, const int64_t* ids, const int64_t* off, const double* x,
                                             const double* y, const double* u, const double* v, double* gx, double* gy, double* gu,
                                             double* gv, const cal_ransac_result* gres, const uint8_t* gmask, cal_ransac_result* res,
                                             uint8_t* mask) {
    const unsigned gb = (unsigned)std::min<int64_t>((cnt * n + 255) / 256, 8);
    simt::launch(gb, 256, [&] { k_gather_views(cnt, n, ids, off, x, y, u, v, gx, gy, gu, gv); });
    simt::launch(gb, 256, [&] { k_scatter_views(cnt, n, ids, off, gres, gmask, res, mask); });
    return 0;
}
