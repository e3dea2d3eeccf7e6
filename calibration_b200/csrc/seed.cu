// Linear seeding stage that feeds the refinement path, batched over views.
//
//   cal_seed_intrinsics     = estimate_intrinsics(views, opts) per camera, no-RANSAC branch
//                             (reference src/estimation/linear/intrinsicsdlt.cpp:101-145):
//                             per-view Hartley-normalised DLT homography (compute_planar_homographies
//                             :33-86 -> HomographyEstimator::fit, homographyestimator.cpp:16-78,123-143),
//                             symmetric_rms_px (:21-31), Zhang's closed form (zhang.cpp:9-208),
//                             sanitize_intrinsics (intrinsics_utils.h:12-62) and per-view
//                             pose_from_homography (posefromhomography.cpp:12-67)
//   cal_seed_planar_poses   = estimate_planar_pose(view, CameraMatrix) per view
//                             (src/estimation/linear/planarpose_linear.cpp:17-76)
//
// Device work decomposition.  k_view_dlt gives 32 consecutive views to one warp.  The O(points)
// passes of a view (centroids, mean distances, the 24 monomial sums that make up A^T A of the DLT
// design matrix, the symmetric transfer error) run warp-cooperatively with coalesced loads and
// shuffle reductions, one view after the other; lane j keeps the sums of view j.  The O(1) dense
// work per view (null vector of the 9x9 normal matrix, de-normalisation, Zhang's two design rows
// or the pose decomposition) then runs lane-private: 32 views at once, nothing computed redundantly.
// Zhang's 6x6 Gram matrix is summed per camera in a fixed order (no atomics); the 6x6 eigenproblem
// and the 3x3 Cholesky that recover K run on the host (O(1) per camera).
//
// Numerical differences to the reference are confined to rounding: null vectors come from the
// normal matrices (inverse iteration for the 9x9 DLT system, Jacobi eigenvalues of Zhang's 6x6 Gram
// matrix) instead of Eigen::JacobiSVD of the design matrices, the orthogonal polar factor (= U V^T of the
// SVD whenever det > 0, which r3 = r1 x r2 guarantees) from a scaled Newton iteration.
#include <cmath>
#include <cstring>
#include <string>
#include <map>
#include <vector>

#include "../../include/calib_b200.h"
#include "dlt.cuh"

extern "C" void cal_set_last_error_(const char* msg);
extern "C" int cal_device_count(void);

namespace {

cal_status sfail(cal_status s, const std::string& m) { cal_set_last_error_(m.c_str()); return s; }
#define SCUDA(expr)                                                                                   \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) return sfail(CAL_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// ---- 3x3 helpers (lane-private) ----
// Orthogonal polar factor of M by scaled Newton iteration X <- (g X + X^-T / g) / 2; equals U V^T of
// the SVD for det M > 0 (project_to_so3, se3_utils.h:10-19; the JacobiSVD step of
// pose_from_homography_normalized, planarpose_linear.cpp:35-41).  false if M is singular.
__device__ bool polar3(const double* M, double* R) {
    double X[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) X[i] = M[i];
    bool ok = true;
    for (int it = 0; it < 14; ++it) {
        const double c00 = X[4] * X[8] - X[5] * X[7], c01 = X[5] * X[6] - X[3] * X[8], c02 = X[3] * X[7] - X[4] * X[6];
        const double det = X[0] * c00 + X[1] * c01 + X[2] * c02;
        if (!(fabs(det) > 1e-300) || !isfinite(det)) { ok = false; break; }
        const double id = 1.0 / det;
        // X^-T = cofactor matrix / det
        double Y[9];
        Y[0] = c00 * id; Y[1] = c01 * id; Y[2] = c02 * id;
        Y[3] = (X[2] * X[7] - X[1] * X[8]) * id; Y[4] = (X[0] * X[8] - X[2] * X[6]) * id; Y[5] = (X[1] * X[6] - X[0] * X[7]) * id;
        Y[6] = (X[1] * X[5] - X[2] * X[4]) * id; Y[7] = (X[2] * X[3] - X[0] * X[5]) * id; Y[8] = (X[0] * X[4] - X[1] * X[3]) * id;
        double g = 1.0;
        if (it < 4) {
            double nx = 0, ny = 0;
#pragma unroll
            for (int i = 0; i < 9; ++i) { nx = fma(X[i], X[i], nx); ny = fma(Y[i], Y[i], ny); }
            g = sqrt(sqrt(ny / nx));
        }
        const double a = 0.5 * g, b = 0.5 / g;
#pragma unroll
        for (int i = 0; i < 9; ++i) X[i] = a * X[i] + b * Y[i];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = X[i];
    return ok;
}

__device__ void identity_pose(double* pose) {
#pragma unroll
    for (int i = 0; i < 12; ++i) pose[i] = (i < 9 && i % 4 == 0) ? 1.0 : 0.0;
}

// pose_from_homography_normalized (planarpose_linear.cpp:17-52)
__device__ void pose_from_h_normalized(const double* H, double* pose) {
    const double n1 = sqrt(H[0] * H[0] + H[3] * H[3] + H[6] * H[6]), n2 = sqrt(H[1] * H[1] + H[4] * H[4] + H[7] * H[7]);
    double s = sqrt(n1 * n2);
    if (s < 1e-12) s = 1.0;
    const double r1[3] = {H[0] / s, H[3] / s, H[6] / s}, r2[3] = {H[1] / s, H[4] / s, H[7] / s};
    const double r3[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
    const double M[9] = {r1[0], r2[0], r3[0], r1[1], r2[1], r3[1], r1[2], r2[2], r3[2]};
    double R[9];
    if (!polar3(M, R)) { identity_pose(pose); return; }
    double t[3] = {H[2] / s, H[5] / s, H[8] / s};
    if (R[8] < 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = -R[i];
        t[0] = -t[0]; t[1] = -t[1]; t[2] = -t[2];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) pose[i] = R[i];
    pose[9] = t[0]; pose[10] = t[1]; pose[11] = t[2];
}

// pose_from_homography (posefromhomography.cpp:12-67); k = fx, fy, cx, cy, skew
__device__ bool pose_from_h(const double* k, const double* H, double* pose) {
    identity_pose(pose);
    if (!isfinite(k[0]) || !isfinite(k[1]) || k[2] <= 0 || k[3] <= 0) return false;
    if (!isfinite(H[8])) return false;
    // K^-1 = [[1/fx, -s/(fx fy), (s cy - cx fy)/(fx fy)], [0, 1/fy, -cy/fy], [0, 0, 1]]
    const double ifx = 1.0 / k[0], ify = 1.0 / k[1];
    double Hn[9];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double r1 = (H[3 + j] - k[3] * H[6 + j]) * ify;
        Hn[3 + j] = r1;
        Hn[j] = (H[j] - k[4] * r1 - k[2] * H[6 + j]) * ifx;
        Hn[6 + j] = H[6 + j];
    }
    const double n1 = sqrt(Hn[0] * Hn[0] + Hn[3] * Hn[3] + Hn[6] * Hn[6]), n2 = sqrt(Hn[1] * Hn[1] + Hn[4] * Hn[4] + Hn[7] * Hn[7]);
    if (!(n1 > 1e-15) || !(n2 > 1e-15)) return false;
    const double s = 1.0 / ((n1 + n2) * 0.5);
    double M[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) { M[3 * i] = s * Hn[3 * i]; M[3 * i + 1] = s * Hn[3 * i + 1]; }
    M[2] = M[3] * M[7] - M[6] * M[4]; M[5] = M[6] * M[1] - M[0] * M[7]; M[8] = M[0] * M[4] - M[3] * M[1];
    double R[9];
    if (!polar3(M, R)) return false;
    double t[3] = {s * Hn[2], s * Hn[5], s * Hn[8]};
    if (t[2] <= 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = -R[i];
        t[0] = -t[0]; t[1] = -t[1]; t[2] = -t[2];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) pose[i] = R[i];
    pose[9] = t[0]; pose[10] = t[1]; pose[11] = t[2];
    return true;
}

// The two rows a homography contributes to Zhang's design matrix (zhang.cpp:92-181)
__device__ void zhang_rows(const double* Hin, double* rows) {
    double H[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) H[i] = Hin[i];
    bool fin = true;
#pragma unroll
    for (int i = 0; i < 9; ++i) fin = fin && isfinite(H[i]);
    if (fin) {  // normalize_hmtx (zhang.cpp:121-147)
        if (H[8] < 0.0) {
#pragma unroll
            for (int i = 0; i < 9; ++i) H[i] = -H[i];
        }
        const double h33 = H[8];
        if (fabs(h33) > 1e-12) {
#pragma unroll
            for (int i = 0; i < 9; ++i) H[i] = H[i] / h33;
        } else {
            double nf = 0;
#pragma unroll
            for (int i = 0; i < 9; ++i) nf = fma(H[i], H[i], nf);
            nf = sqrt(nf);
            if (nf > 1e-12) {
#pragma unroll
                for (int i = 0; i < 9; ++i) H[i] = H[i] / nf;
            }
        }
    }
    auto vij = [&](int i, int j, double* v) {
        const double h0i = H[i], h1i = H[3 + i], h2i = H[6 + i], h0j = H[j], h1j = H[3 + j], h2j = H[6 + j];
        v[0] = h0i * h0j; v[1] = h0i * h1j + h1i * h0j; v[2] = h1i * h1j;
        v[3] = h0i * h2j + h2i * h0j; v[4] = h1i * h2j + h2i * h1j; v[5] = h2i * h2j;
    };
    double v12[6], v11[6], v22[6];
    vij(0, 1, v12); vij(0, 0, v11); vij(1, 1, v22);
    double s1 = 0, s2 = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) { v11[i] -= v22[i]; s1 = fma(v12[i], v12[i], s1); s2 = fma(v11[i], v11[i], s2); }
    s1 = sqrt(s1); s2 = sqrt(s2);
#pragma unroll
    for (int i = 0; i < 6; ++i) { rows[i] = s1 > 0 ? v12[i] / s1 : v12[i]; rows[6 + i] = s2 > 0 ? v11[i] / s2 : v11[i]; }
}

struct SeedArgs {
    int64_t n_views;
    const int64_t* view_offset;
    const int32_t* view_cam;
    const double *x, *y, *u, *v;
    const double* kmtx;     // [n_cams][5] (planar-pose mode)
    double* hmtx;           // [n_views][9]
    double* sym_rms;        // [n_views]
    double* zrows;          // [n_views][12]
    int32_t* success;       // [n_views]
    double* poses;          // [n_views][12] (planar-pose mode)
};

// POSE = false: pixel homography + symmetric rms + Zhang rows.  POSE = true: estimate_planar_pose.
template <bool POSE>
__global__ void __launch_bounds__(128) k_view_dlt(SeedArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (base >= a.n_views) return;
    // sums of "my" view (view base + lane)
    double my_c[6] = {0, 0, 0, 0, 1, 1};  // scx, scy, dcx, dcy, ss, ds
    double my_m[4][6];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int e = 0; e < 6; ++e) my_m[k][e] = 0.0;
    int my_n = 0;
    const int nv = (int)min((int64_t)32, a.n_views - base);
    for (int j = 0; j < nv; ++j) {
        const int64_t o = a.view_offset[base + j];
        const int n = (int)(a.view_offset[base + j + 1] - o);
        double fx = 1, fy = 1, cx = 0, cy = 0, sk = 0;
        if (POSE) { const double* k = a.kmtx + 5 * a.view_cam[base + j]; fx = k[0]; fy = k[1]; cx = k[2]; cy = k[3]; sk = k[4]; }
        auto img = [&](int i, double& uu, double& vv) {
            uu = a.u[o + i]; vv = a.v[o + i];
            if (POSE) { vv = (vv - cy) / fy; uu = (uu - cx - sk * vv) / fx; }  // normalize (camera_matrix.h:34-40)
        };
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int i = lane; i < n; i += 32) { double uu, vv; img(i, uu, vv); s0 += a.x[o + i]; s1 += a.y[o + i]; s2 += uu; s3 += vv; }
        const double inv = 1.0 / (double)max(n, 1);
        const double scx = warp_sum(s0) * inv, scy = warp_sum(s1) * inv, dcx = warp_sum(s2) * inv, dcy = warp_sum(s3) * inv;
        s0 = s1 = 0;
        for (int i = lane; i < n; i += 32) {
            double uu, vv; img(i, uu, vv);
            const double dx = a.x[o + i] - scx, dy = a.y[o + i] - scy, du = uu - dcx, dv = vv - dcy;
            s0 += sqrt(dx * dx + dy * dy); s1 += sqrt(du * du + dv * dv);
        }
        const double sm = warp_sum(s0) * inv, dm = warp_sum(s1) * inv;
        const double ss = sm > 0 ? 1.4142135623730951 / sm : 1.0, ds = dm > 0 ? 1.4142135623730951 / dm : 1.0;
        double m[4][6];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int e = 0; e < 6; ++e) m[k][e] = 0.0;
        for (int i = lane; i < n; i += 32) {
            double uu, vv; img(i, uu, vv);
            const double px = ss * (a.x[o + i] - scx), py = ss * (a.y[o + i] - scy), qu = ds * (uu - dcx), qv = ds * (vv - dcy);
            const double pp[6] = {px * px, px * py, px, py * py, py, 1.0};
            const double wt[4] = {1.0, qu, qv, qu * qu + qv * qv};
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int e = 0; e < 6; ++e) m[k][e] = fma(wt[k], pp[e], m[k][e]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int e = 0; e < 6; ++e) { const double t = warp_sum(m[k][e]); if (lane == j) my_m[k][e] = t; }
        if (lane == j) { my_c[0] = scx; my_c[1] = scy; my_c[2] = dcx; my_c[3] = dcy; my_c[4] = ss; my_c[5] = ds; my_n = n; }
    }
    // ---- lane-private: null vector of A^T A, de-normalisation (homographyestimator.cpp:45-78) ----
    double H[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    bool ok = lane < nv && my_n >= 4;
    {
        double Gu[45]; dlt_normal_matrix(my_m, Gu);
        double z[9];
        const bool e = smallest_eigvec9(Gu, z);
        if (ok && e) {
            double hn[9];
            const double ih = 1.0 / z[8];
#pragma unroll
            for (int i = 0; i < 9; ++i) hn[i] = z[i] * ih;
            denormalise(hn, my_c[4], my_c[0], my_c[1], my_c[5], my_c[2], my_c[3], H);
            ok = isfinite(H[0]);
        } else ok = false;
        if (ok && fabs(H[8]) > 1e-15) {  // intrinsicsdlt.cpp:76-78 / planarpose_linear.cpp:72-74
            const double h = H[8];
#pragma unroll
            for (int i = 0; i < 9; ++i) H[i] /= h;
        }
        if (!ok) {
#pragma unroll
            for (int i = 0; i < 9; ++i) H[i] = (i % 4 == 0) ? 1.0 : 0.0;
        }
    }
    const int64_t view = base + lane;
    if (POSE) {
        if (lane < nv) {
            double pose[12];
            if (ok) pose_from_h_normalized(H, pose); else identity_pose(pose);
#pragma unroll
            for (int i = 0; i < 12; ++i) a.poses[view * 12 + i] = pose[i];
            if (a.success) a.success[view] = ok ? 1 : 0;
        }
        return;
    } else {
        // ---- symmetric_rms_px (intrinsicsdlt.cpp:21-31): warp-cooperative again, view after view ----
        double my_rms = 0.0;
        for (int j = 0; j < nv; ++j) {
            double Hj[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) Hj[i] = __shfl_sync(kFull, H[i], j);
            if (!__shfl_sync(kFull, ok ? 1 : 0, j)) continue;
            double Hi[9]; inv3(Hj, Hi);
            const int64_t o = a.view_offset[base + j];
            const int n = (int)(a.view_offset[base + j + 1] - o);
            double s = 0.0;
            for (int i = lane; i < n; i += 32) {
                const double x = a.x[o + i], y = a.y[o + i], u = a.u[o + i], v = a.v[o + i];
                const double qz = Hj[6] * x + Hj[7] * y + Hj[8];
                const double du = u - (Hj[0] * x + Hj[1] * y + Hj[2]) / qz, dv = v - (Hj[3] * x + Hj[4] * y + Hj[5]) / qz;
                const double pz = Hi[6] * u + Hi[7] * v + Hi[8];
                const double dx = x - (Hi[0] * u + Hi[1] * v + Hi[2]) / pz, dy = y - (Hi[3] * u + Hi[4] * v + Hi[5]) / pz;
                s += sqrt(0.5 * (du * du + dv * dv + dx * dx + dy * dy));
            }
            s = warp_sum(s);
            if (lane == j) my_rms = sqrt(s / (2.0 * (double)n));
        }
        if (lane < nv) {
            double rows[12];
            zhang_rows(H, rows);
#pragma unroll
            for (int i = 0; i < 9; ++i) a.hmtx[view * 9 + i] = H[i];
#pragma unroll
            for (int i = 0; i < 12; ++i) a.zrows[view * 12 + i] = ok ? rows[i] : 0.0;
            a.sym_rms[view] = ok ? my_rms : 0.0;
            a.success[view] = ok ? 1 : 0;
        }
    }
}

// compute_planar_homographies, RANSAC branch (intrinsicsdlt.cpp:50-64): model / h33, symmetric rms over the
// inliers (already evaluated by the RANSAC kernel; the residual does not depend on the scale of H)
__global__ void k_seed_from_ransac(int64_t n_views, const cal_ransac_result* __restrict__ res, double* __restrict__ hmtx,
                                   double* __restrict__ sym_rms, double* __restrict__ zrows, int32_t* __restrict__ success) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_views) return;
    const cal_ransac_result r = res[v];
    double H[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) H[i] = r.success ? r.hmtx[i] : ((i % 4 == 0) ? 1.0 : 0.0);
    if (r.success && fabs(H[8]) > 1e-15) {
        const double h = H[8];
#pragma unroll
        for (int i = 0; i < 9; ++i) H[i] /= h;
    }
    double rows[12];
    zhang_rows(H, rows);
#pragma unroll
    for (int i = 0; i < 9; ++i) hmtx[v * 9 + i] = H[i];
#pragma unroll
    for (int i = 0; i < 12; ++i) zrows[v * 12 + i] = r.success ? rows[i] : 0.0;
    sym_rms[v] = r.success ? r.symmetric_rms_px : 0.0;
    success[v] = r.success ? 1 : 0;
}

// ragged views through the equal-size RANSAC kernel: gather the views `ids` (all of n points) into a dense
// [group][n] batch, and scatter the batch's results / inlier masks back to their views
__global__ void k_gather_views(int64_t cnt, int n, const int64_t* __restrict__ ids, const int64_t* __restrict__ off,
                               const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ u,
                               const double* __restrict__ v, double* __restrict__ gx, double* __restrict__ gy, double* __restrict__ gu,
                               double* __restrict__ gv) {
    const int64_t total = cnt * n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = t / n; const int i = (int)(t - p * n);
        const int64_t s = off[ids[p]] + i;
        gx[t] = x[s]; gy[t] = y[s]; gu[t] = u[s]; gv[t] = v[s];
    }
}
__global__ void k_scatter_views(int64_t cnt, int n, const int64_t* __restrict__ ids, const int64_t* __restrict__ off,
                                const cal_ransac_result* __restrict__ gres, const uint8_t* __restrict__ gmask,
                                cal_ransac_result* __restrict__ res, uint8_t* __restrict__ mask) {
    const int64_t total = cnt * n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = t / n; const int i = (int)(t - p * n);
        if (i == 0) res[ids[p]] = gres[p];
        if (mask) mask[off[ids[p]] + i] = gmask[t];
    }
}

// Zhang's Gram matrix V^T V (upper triangle, 21) of one camera: one CTA per camera, strided
// per-thread sums over the views, fixed-order shared-memory tree.
__global__ void __launch_bounds__(256) k_zhang_gram(int64_t n_views, const int32_t* __restrict__ view_cam,
                                                    const int32_t* __restrict__ success, const double* __restrict__ zrows,
                                                    double* __restrict__ gram /*[n_cams][22]: 21 + count*/) {
    __shared__ double sm[256];
    const int cam = blockIdx.x;
    double g[21];
#pragma unroll
    for (int i = 0; i < 21; ++i) g[i] = 0.0;
    double cnt = 0.0;
    for (int64_t v = threadIdx.x; v < n_views; v += 256) {
        if (view_cam[v] != cam || !success[v]) continue;
        const double* r = zrows + v * 12;
        cnt += 1.0;
        int o = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) { g[o] = fma(r[i], r[j], fma(r[6 + i], r[6 + j], g[o])); ++o; }
    }
    for (int e = 0; e < 22; ++e) {
        sm[threadIdx.x] = e < 21 ? g[e] : cnt;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s]; __syncthreads(); }
        if (threadIdx.x == 0) gram[cam * 22 + e] = sm[0];
        __syncthreads();
    }
}

__global__ void k_pose_from_h(int64_t n_views, const int32_t* __restrict__ view_cam, const int32_t* __restrict__ success,
                              const int32_t* __restrict__ cam_ok, const double* __restrict__ kmtx, const double* __restrict__ hmtx,
                              double* __restrict__ poses) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_views) return;
    double pose[12]; identity_pose(pose);
    const int cam = view_cam[v];
    if (success[v] && cam_ok[cam]) {
        double H[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) H[i] = hmtx[v * 9 + i];
        pose_from_h(kmtx + 5 * cam, H, pose);
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) poses[v * 12 + i] = pose[i];
}

// ---- host side ----
// Eigenvector of the smallest eigenvalue of a symmetric positive semi-definite 6x6 by cyclic Jacobi
// rotations with the relative stopping rule |a_pq| <= eps sqrt(a_pp a_qq): for graded matrices such
// as Zhang's Gram matrix (columns scale like f^2, f, 1) Jacobi resolves the small eigenpairs to the
// accuracy the column-scaled matrix allows (Demmel & Veselic), which an unscaled QR iteration would not.
void smallest_eigvec6(double A[6][6], double* vec) {
    double V[6][6];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 5; ++p) for (int q = p + 1; q < 6; ++q) {
            const double apq = A[p][q];
            if (apq == 0.0 || std::fabs(apq) <= 1e-18 * std::sqrt(std::fabs(A[p][p] * A[q][q]))) continue;
            rotated = true;
            const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
            const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < 6; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
            for (int k = 0; k < 6; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
            for (int k = 0; k < 6; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
        }
        if (!rotated) break;
    }
    int m = 0; for (int j = 1; j < 6; ++j) if (A[j][j] < A[m][m]) m = j;
    for (int i = 0; i < 6; ++i) vec[i] = V[i][m];
}

// K from b = (B11, B12, B22, B13, B23, B33), B = K^-T K^-1 up to scale and sign (zhang.cpp:29-90)
bool kmtx_from_conic(const double* bin, double* k5) {
    for (int sign = 0; sign < 2; ++sign) {
        double b[6]; for (int i = 0; i < 6; ++i) b[i] = sign ? -bin[i] : bin[i];
        bool fin = true; for (int i = 0; i < 6; ++i) fin = fin && std::isfinite(b[i]);
        if (!fin) continue;
        // B = U^T U with U upper triangular (Cholesky); U = K^-1 up to scale
        const double B[3][3] = {{b[0], b[1], b[3]}, {b[1], b[2], b[4]}, {b[3], b[4], b[5]}};
        double L[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        bool pd = true;
        for (int j = 0; j < 3 && pd; ++j) {
            double s = B[j][j]; for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
            if (!(s > 0.0)) { pd = false; break; }
            L[j][j] = std::sqrt(s);
            for (int i = j + 1; i < 3; ++i) { double t = B[i][j]; for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k]; L[i][j] = t / L[j][j]; }
        }
        if (!pd) continue;
        // U = L^T; K = U^-1 (upper triangular inverse), then K /= K22
        const double u00 = L[0][0], u01 = L[1][0], u02 = L[2][0], u11 = L[1][1], u12 = L[2][1], u22 = L[2][2];
        double k00 = 1.0 / u00, k11 = 1.0 / u11, k22 = 1.0 / u22;
        double k01 = -u01 * k00 * k11, k12 = -u12 * k11 * k22, k02 = (u01 * u12 - u02 * u11) * k00 * k11 * k22;
        if (!std::isfinite(k00) || !std::isfinite(k11) || !std::isfinite(k22) || !std::isfinite(k01) || !std::isfinite(k02) || !std::isfinite(k12)) continue;
        if (std::fabs(k22) < 1e-15) continue;
        k00 /= k22; k01 /= k22; k02 /= k22; k11 /= k22; k12 /= k22;
        if (k00 <= 0.0 || k11 <= 0.0) { k00 = -k00; k01 = -k01; k02 = -k02; k11 = -k11; k12 = -k12; }
        k5[0] = k00; k5[1] = k11; k5[2] = k02; k5[3] = k12; k5[4] = k01;
        return true;
    }
    return false;
}

void sanitize(double* k5, const cal_seed_options& o) {  // intrinsics_utils.h:12-62
    auto min_focal = [](double v, double mn) { return (!std::isfinite(v) || v < mn) ? mn : v; };
    auto principal = [](double v, double mn, double mx) { return (!std::isfinite(v) || v < mn || v > mx) ? 0.5 * (mn + mx) : v; };
    k5[0] = min_focal(k5[0], o.fx_min); k5[1] = min_focal(k5[1], o.fy_min);
    k5[2] = principal(k5[2], o.cx_min, o.cx_max); k5[3] = principal(k5[3], o.cy_min, o.cy_max);
    const double smin = std::min(o.skew_min, o.skew_max), smax = std::max(o.skew_min, o.skew_max);
    if (!std::isfinite(k5[4]) || k5[4] < smin || k5[4] > smax) k5[4] = std::min(std::max(0.0, smin), smax);
}

// device view of a caller array that may live on the host or already on the device
template <class T>
struct DevIn {
    const T* p = nullptr; T* owned = nullptr;
    cudaError_t init(const T* src, size_t n, cudaStream_t st) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeDevice) { p = src; return cudaSuccess; }
        cudaGetLastError();
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&owned), std::max<size_t>(n, 1) * sizeof(T));
        if (e != cudaSuccess) return e;
        p = owned;
        return cudaMemcpyAsync(owned, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
    }
    ~DevIn() { if (owned) cudaFree(owned); }
};
template <class T>
struct DevBuf {
    T* p = nullptr;
    cudaError_t alloc(size_t n) { return cudaMalloc(reinterpret_cast<void**>(&p), std::max<size_t>(n, 1) * sizeof(T)); }
    ~DevBuf() { if (p) cudaFree(p); }
};

struct SeedInputs {
    DevIn<int64_t> off; DevIn<int32_t> cam; DevIn<double> x, y, u, v;
};

cal_status check_views(int64_t n_views, const int64_t* view_offset, const int32_t* view_cam, int32_t n_cams, int64_t* n_obs) {
    if (n_views <= 0 || !view_offset || !view_cam || n_cams <= 0) return sfail(CAL_ERR_INVALID_ARGUMENT, "bad argument");
    if (view_offset[0] != 0) return sfail(CAL_ERR_INVALID_ARGUMENT, "view_offset must start at 0");
    for (int64_t k = 0; k < n_views; ++k) {
        if (view_offset[k + 1] < view_offset[k]) return sfail(CAL_ERR_INVALID_ARGUMENT, "view_offset must be non-decreasing");
        if (view_cam[k] < 0 || view_cam[k] >= n_cams) return sfail(CAL_ERR_INVALID_ARGUMENT, "view_cam out of range");
    }
    *n_obs = view_offset[n_views];
    return CAL_OK;
}

}  // namespace

extern "C" cal_status cal_ransac_homography_batch_dev(int64_t n_problems, int32_t n, const double* x_dev, const double* y_dev,
                                                      const double* u_dev, const double* v_dev, const cal_ransac_options* opts,
                                                      int seed_per_problem, cal_ransac_result* results_dev,
                                                      uint8_t* inlier_mask_dev, float* ms);

extern "C" cal_status cal_seed_intrinsics_ransac(int64_t n_views, const int64_t* view_offset, const int32_t* view_cam, const double* x,
                                                 const double* y, const double* u, const double* v, int32_t n_cams,
                                                 const cal_seed_options* opts, const cal_ransac_options* ransac, int device, double* kmtx,
                                                 int32_t* cam_success, int32_t* view_success, double* hmtx, double* sym_rms,
                                                 double* poses, uint8_t* inlier_mask) {
    if (!x || !y || !u || !v || !kmtx || !cam_success) return sfail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    int64_t n_obs = 0;
    if (cal_status s = check_views(n_views, view_offset, view_cam, n_cams, &n_obs)) return s;
    if (cal_device_count() <= device) return sfail(CAL_ERR_CUDA, "no CUDA device: calib_b200 has no CPU fallback");
    SCUDA(cudaSetDevice(device));
    cudaStream_t st; SCUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
    SeedInputs in;
    SCUDA(in.off.init(view_offset, n_views + 1, st)); SCUDA(in.cam.init(view_cam, n_views, st));
    SCUDA(in.x.init(x, n_obs, st)); SCUDA(in.y.init(y, n_obs, st)); SCUDA(in.u.init(u, n_obs, st)); SCUDA(in.v.init(v, n_obs, st));
    DevBuf<double> dH, drms, drows, dgram, dk, dposes; DevBuf<int32_t> dsucc, dcamok;
    SCUDA(dH.alloc(9 * n_views)); SCUDA(drms.alloc(n_views)); SCUDA(drows.alloc(12 * n_views)); SCUDA(dsucc.alloc(n_views));
    SCUDA(dgram.alloc(22 * n_cams)); SCUDA(dk.alloc(5 * n_cams)); SCUDA(dcamok.alloc(n_cams)); SCUDA(dposes.alloc(12 * n_views));
    const unsigned groups = (unsigned)((n_views + 31) / 32);
    DevBuf<cal_ransac_result> dres; DevBuf<uint8_t> dmask;
    if (ransac) {
        // IntrinsicsEstimOptions::homography_ransac (intrinsicsdlt.cpp:50-64): every view is one problem of the batched
        // RANSAC kernel, all with the seed of the options (each reference call constructs its own engine from opts.seed)
        const int64_t n = view_offset[1] - view_offset[0];
        bool equal = true;
        for (int64_t k = 0; k < n_views; ++k) {
            const int64_t nk = view_offset[k + 1] - view_offset[k];
            if (nk > 0x7fffffff) return sfail(CAL_ERR_INVALID_ARGUMENT, "homography_ransac: view too large");
            equal = equal && nk == n;
        }
        SCUDA(dres.alloc(n_views)); if (inlier_mask) SCUDA(dmask.alloc((size_t)n_obs));
        if (equal) {
            if (n <= 0) return sfail(CAL_ERR_INVALID_ARGUMENT, "homography_ransac: empty views");
            SCUDA(cudaStreamSynchronize(st));  // the inputs are in place before the RANSAC launch on the default stream
            if (cal_status rs = cal_ransac_homography_batch_dev(n_views, (int32_t)n, in.x.p, in.y.p, in.u.p, in.v.p, ransac, 0, dres.p,
                                                                inlier_mask ? dmask.p : nullptr, nullptr)) return rs;
        } else {
            // ragged views: the batched kernel takes problems of one size, so the views are grouped by size; each
            // group is gathered into a dense [group][n] batch, solved by one launch and scattered back (results per
            // view, inlier mask in the observations' own CSR layout).  Views with fewer than four points stay
            // unsuccessful (intrinsicsdlt.cpp:41-45).
            std::map<int64_t, std::vector<int64_t>> by_size;
            for (int64_t k = 0; k < n_views; ++k) by_size[view_offset[k + 1] - view_offset[k]].push_back(k);
            SCUDA(cudaMemsetAsync(dres.p, 0, (size_t)n_views * sizeof(cal_ransac_result), st));
            if (inlier_mask) SCUDA(cudaMemsetAsync(dmask.p, 0, (size_t)n_obs, st));
            for (const auto& [ng, ids] : by_size) {
                if (ng < 4) continue;
                const int64_t cnt = (int64_t)ids.size();
                DevBuf<int64_t> dids; DevBuf<double> g; DevBuf<cal_ransac_result> gres; DevBuf<uint8_t> gmask;
                SCUDA(dids.alloc(cnt)); SCUDA(g.alloc((size_t)4 * cnt * ng)); SCUDA(gres.alloc(cnt));
                if (inlier_mask) SCUDA(gmask.alloc((size_t)cnt * ng));
                SCUDA(cudaMemcpyAsync(dids.p, ids.data(), (size_t)cnt * sizeof(int64_t), cudaMemcpyHostToDevice, st));
                double *gx = g.p, *gy = gx + cnt * ng, *gu = gy + cnt * ng, *gv = gu + cnt * ng;
                const unsigned gb = (unsigned)std::min<int64_t>((cnt * ng + 255) / 256, 148 * 8);
                k_gather_views<<<gb, 256, 0, st>>>(cnt, (int)ng, dids.p, in.off.p, in.x.p, in.y.p, in.u.p, in.v.p, gx, gy, gu, gv);
                SCUDA(cudaStreamSynchronize(st));  // ids (a host vector) consumed; inputs in place before the default-stream launch
                if (cal_status rs = cal_ransac_homography_batch_dev(cnt, (int32_t)ng, gx, gy, gu, gv, ransac, 0, gres.p,
                                                                    inlier_mask ? gmask.p : nullptr, nullptr)) return rs;
                k_scatter_views<<<gb, 256, 0, st>>>(cnt, (int)ng, dids.p, in.off.p, gres.p, inlier_mask ? gmask.p : nullptr, dres.p,
                                                    inlier_mask ? dmask.p : nullptr);
                SCUDA(cudaStreamSynchronize(st));  // the group's buffers are released at the end of this iteration
            }
        }
        k_seed_from_ransac<<<(unsigned)((n_views + 127) / 128), 128, 0, st>>>(n_views, dres.p, dH.p, drms.p, drows.p, dsucc.p);
        if (inlier_mask) SCUDA(cudaMemcpyAsync(inlier_mask, dmask.p, (size_t)n_obs, cudaMemcpyDeviceToHost, st));
    } else {
        SeedArgs a{n_views, in.off.p, in.cam.p, in.x.p, in.y.p, in.u.p, in.v.p, nullptr, dH.p, drms.p, drows.p, dsucc.p, nullptr};
        k_view_dlt<false><<<(groups + 3) / 4, 128, 0, st>>>(a);
    }
    k_zhang_gram<<<n_cams, 256, 0, st>>>(n_views, in.cam.p, dsucc.p, drows.p, dgram.p);
    std::vector<double> gram(22 * (size_t)n_cams);
    SCUDA(cudaMemcpyAsync(gram.data(), dgram.p, gram.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    SCUDA(cudaStreamSynchronize(st));
    SCUDA(cudaGetLastError());
    std::vector<int32_t> camok(n_cams, 0);
    for (int c = 0; c < n_cams; ++c) {
        double* k5 = kmtx + 5 * c;
        for (int i = 0; i < 5; ++i) k5[i] = 0.0;
        const double* g = &gram[22 * (size_t)c];
        if (g[21] < 4.0) continue;  // "Zhang method requires at least 4 views" (zhang.cpp:152-155)
        // b = right singular vector of the smallest singular value of V = eigenvector of the smallest
        // eigenvalue of V^T V (the SAME minimiser of |V b| over |b| = 1 as zhang.cpp:193-196)
        double G[6][6], b[6];
        { int o = 0; for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { G[i][j] = G[j][i] = g[o]; ++o; } }
        smallest_eigvec6(G, b);
        if (!kmtx_from_conic(b, k5)) { for (int i = 0; i < 5; ++i) k5[i] = 0.0; continue; }
        if (opts && opts->use_bounds) sanitize(k5, *opts);
        camok[c] = 1;
    }
    for (int c = 0; c < n_cams; ++c) cam_success[c] = camok[c];
    SCUDA(cudaMemcpyAsync(dk.p, kmtx, sizeof(double) * 5 * n_cams, cudaMemcpyHostToDevice, st));
    SCUDA(cudaMemcpyAsync(dcamok.p, camok.data(), sizeof(int32_t) * n_cams, cudaMemcpyHostToDevice, st));
    if (poses) {
        k_pose_from_h<<<(unsigned)((n_views + 127) / 128), 128, 0, st>>>(n_views, in.cam.p, dsucc.p, dcamok.p, dk.p, dH.p, dposes.p);
        SCUDA(cudaMemcpyAsync(poses, dposes.p, sizeof(double) * 12 * n_views, cudaMemcpyDeviceToHost, st));
    }
    if (hmtx) SCUDA(cudaMemcpyAsync(hmtx, dH.p, sizeof(double) * 9 * n_views, cudaMemcpyDeviceToHost, st));
    if (sym_rms) SCUDA(cudaMemcpyAsync(sym_rms, drms.p, sizeof(double) * n_views, cudaMemcpyDeviceToHost, st));
    if (view_success) SCUDA(cudaMemcpyAsync(view_success, dsucc.p, sizeof(int32_t) * n_views, cudaMemcpyDeviceToHost, st));
    SCUDA(cudaStreamSynchronize(st));
    SCUDA(cudaGetLastError());
    return CAL_OK;
}

extern "C" cal_status cal_seed_intrinsics(int64_t n_views, const int64_t* view_offset, const int32_t* view_cam, const double* x,
                                          const double* y, const double* u, const double* v, int32_t n_cams,
                                          const cal_seed_options* opts, int device, double* kmtx, int32_t* cam_success,
                                          int32_t* view_success, double* hmtx, double* sym_rms, double* poses) {
    return cal_seed_intrinsics_ransac(n_views, view_offset, view_cam, x, y, u, v, n_cams, opts, nullptr, device, kmtx, cam_success,
                                      view_success, hmtx, sym_rms, poses, nullptr);
}

extern "C" cal_status cal_seed_planar_poses(int64_t n_views, const int64_t* view_offset, const int32_t* view_cam, const double* x,
                                            const double* y, const double* u, const double* v, int32_t n_cams, const double* kmtx,
                                            int device, double* poses, int32_t* view_success) {
    if (!x || !y || !u || !v || !kmtx || !poses) return sfail(CAL_ERR_INVALID_ARGUMENT, "null argument");
    int64_t n_obs = 0;
    if (cal_status s = check_views(n_views, view_offset, view_cam, n_cams, &n_obs)) return s;
    if (cal_device_count() <= device) return sfail(CAL_ERR_CUDA, "no CUDA device: calib_b200 has no CPU fallback");
    SCUDA(cudaSetDevice(device));
    cudaStream_t st; SCUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
    SeedInputs in;
    SCUDA(in.off.init(view_offset, n_views + 1, st)); SCUDA(in.cam.init(view_cam, n_views, st));
    SCUDA(in.x.init(x, n_obs, st)); SCUDA(in.y.init(y, n_obs, st)); SCUDA(in.u.init(u, n_obs, st)); SCUDA(in.v.init(v, n_obs, st));
    DevBuf<double> dk, dposes; DevBuf<int32_t> dsucc;
    SCUDA(dk.alloc(5 * n_cams)); SCUDA(dposes.alloc(12 * n_views)); SCUDA(dsucc.alloc(n_views));
    SCUDA(cudaMemcpyAsync(dk.p, kmtx, sizeof(double) * 5 * n_cams, cudaMemcpyHostToDevice, st));
    SeedArgs a{n_views, in.off.p, in.cam.p, in.x.p, in.y.p, in.u.p, in.v.p, dk.p, nullptr, nullptr, nullptr, dsucc.p, dposes.p};
    const unsigned groups = (unsigned)((n_views + 31) / 32);
    k_view_dlt<true><<<(groups + 3) / 4, 128, 0, st>>>(a);
    SCUDA(cudaMemcpyAsync(poses, dposes.p, sizeof(double) * 12 * n_views, cudaMemcpyDeviceToHost, st));
    if (view_success) SCUDA(cudaMemcpyAsync(view_success, dsucc.p, sizeof(int32_t) * n_views, cudaMemcpyDeviceToHost, st));
    SCUDA(cudaStreamSynchronize(st));
    SCUDA(cudaGetLastError());
    return CAL_OK;
}
