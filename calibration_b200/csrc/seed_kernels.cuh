// Device code of the linear seeding stage (see seed.cu for the description).  Kept in a header so that
// tests/host_emul can run this very source on the CPU under a lock-step SIMT shim (test-only).
#pragma once
#include "../../include/calib_b200.h"
#include "dlt.cuh"

namespace {

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
// Three CTAs per SM (168 registers, ~50 bytes of spills): the kernel waits on global loads (long scoreboard 2.9 per issue) with few
// warps to interleave — 1.07 ms at two CTAs per SM (184 registers), 0.75 ms at three, 0.78 ms at four (128 registers, 470 bytes of
// spills), 160 000 views on a B200 (tools/seed_ab.sh).
#ifndef CALK_DLT_MINB
#define CALK_DLT_MINB 3
#endif
template <bool POSE>
__global__ void __launch_bounds__(128, CALK_DLT_MINB) k_view_dlt(SeedArgs a) {
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

}  // namespace
