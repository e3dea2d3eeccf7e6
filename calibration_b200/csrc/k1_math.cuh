// Per-observation arithmetic of the fused residual + Jacobian + J^T J pass (K1)
// and the per-block chain-rule transforms, written once as host/device inline
// functions.  The CUDA kernels in refine_kernels.cu are the only product
// callers; tests/host_emul compiles the same header with g++ to check the
// algebra against the CPU oracle without a GPU (test-only, never shipped).
//
// Formulation.  For one residual block (one view seen by one
// camera) the reference evaluates, per corner (X, Y):
//     P = R_ct (X, Y, 0)^T + t_ct ;  (u, v) = project(P ; intr) ;  r = (u, v) - (u_obs, v_obs)
// (src/estimation/residuals/{intrinsic,extrinsics,bundle}residual.h) and
// differentiates r w.r.t. up to two pose blocks and the intrinsics with Jets.
// Here every pose chain is reduced to ONE 6-dof left twist xi = (omega, nu) of
// the composite pose in the (sensor-rotated) camera frame, dP = omega x P + nu,
// so the per-observation Jacobian has 6 + PI columns whatever the chain; the
// chain rule to the actual parameter blocks is a constant 6x6 per block
// (block_transforms below) applied once per block, not per corner.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define CAL_HD __host__ __device__ __forceinline__
#else
#define CAL_HD inline
#endif

namespace calk {

enum { INTR_NONE = 0, INTR_NOSKEW = 1, INTR_SKEW = 2 };

// 1 / x for the projection's depth: on the device the hardware seed (MUFU.RCP64H, ~20 bits) and two
// Newton steps, straight-line code — the compiler's own division sequence carries a special-case branch
// (a reconvergence point in the middle of K1's hot loop).  ~1 ulp; depths are O(1) metres, never denormal.
CAL_HD double rcp_depth(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0); r = fma(r, e, r);
    e = fma(-x, r, 1.0); r = fma(r, e, r);
    e = fma(-x, r, 1.0); r = fma(r, e, r);
    return r;
#else
    return 1.0 / x;
#endif
}

// Per-camera constants derived from the intrinsic block
// [fx, fy, cx, cy, skew, k1, k2, k3, p1, p2 (, tau_x, tau_y)].
struct CamConst {
    double fx, fy, cx, cy, sk, k1, k2, k3, p1, p2;
    double mx0, my0;          // principal-ray intersection (scheimpflug.h:165-167); 0 for pinhole
    double dmx0[2], dmy0[2];  // d(mx0, my0) / d(tau_x, tau_y)
    double wy[3];             // axial vector of (dRs^T/dtau_y) Rs ; the tau_x one is (-1, 0, 0)
    double Rs[9];             // rot_sensor (scheimpflug.h:150-153), row-major; identity for pinhole
};

CAL_HD void cam_const_from_intr(const double* intr, int model, CamConst& c) {
    c.fx = intr[0]; c.fy = intr[1]; c.cx = intr[2]; c.cy = intr[3]; c.sk = intr[4];
    c.k1 = intr[5]; c.k2 = intr[6]; c.k3 = intr[7]; c.p1 = intr[8]; c.p2 = intr[9];
    if (model == 1) {
        const double tx = intr[10], ty = intr[11];
        const double ctx = cos(tx), stx = sin(tx), cty = cos(ty), sty = sin(ty);
        c.Rs[0] = cty;  c.Rs[1] = stx * sty; c.Rs[2] = ctx * sty;
        c.Rs[3] = 0.0;  c.Rs[4] = ctx;       c.Rs[5] = -stx;
        c.Rs[6] = -sty; c.Rs[7] = stx * cty; c.Rs[8] = ctx * cty;
        c.mx0 = -sty / (ctx * cty); c.my0 = stx / ctx;
        c.dmx0[0] = -sty * stx / (ctx * ctx * cty); c.dmx0[1] = -1.0 / (ctx * cty * cty);
        c.dmy0[0] = 1.0 / (ctx * ctx); c.dmy0[1] = 0.0;
        c.wy[0] = 0.0; c.wy[1] = -ctx; c.wy[2] = stx;
    } else {
        for (int i = 0; i < 9; ++i) c.Rs[i] = (i % 4 == 0) ? 1.0 : 0.0;
        c.mx0 = c.my0 = 0.0; c.dmx0[0] = c.dmx0[1] = c.dmy0[0] = c.dmy0[1] = 0.0;
        c.wy[0] = c.wy[1] = c.wy[2] = 0.0;
    }
}

// Column layout of the local (per-block) system: [omega(3) nu(3) | intr tangent | r]
template <int MODEL, int IMODE>
struct Local {
    static constexpr int P = MODEL == 1 ? 12 : 10;
    static constexpr int PI = IMODE == INTR_NONE ? 0 : (IMODE == INTR_NOSKEW ? P - 1 : P);
    static constexpr int NC = 6 + PI;
    static constexpr int NL = NC + 1;
    static constexpr int NE = NL * (NL + 1) / 2;
    // intr tangent column ids (relative to column 6); -1 when absent
    static constexpr int c_fx = IMODE == INTR_NONE ? -1 : 0;
    static constexpr int c_fy = IMODE == INTR_NONE ? -1 : 1;
    static constexpr int c_cx = IMODE == INTR_NONE ? -1 : 2;
    static constexpr int c_cy = IMODE == INTR_NONE ? -1 : 3;
    static constexpr int c_sk = IMODE == INTR_SKEW ? 4 : -1;
    static constexpr int c_d0 = IMODE == INTR_NONE ? -1 : (IMODE == INTR_SKEW ? 5 : 4);  // k1,k2,k3,p1,p2
    static constexpr int c_tau = (MODEL == 1 && IMODE != INTR_NONE) ? c_d0 + 5 : -1;
    // structural sparsity of the u / v Jacobian rows; column NC is the residual itself
    static constexpr bool has_u(int c) {
        return c < 6 || c == NC || (c - 6 != c_fy && c - 6 != c_cy);
    }
    static constexpr bool has_v(int c) {
        return c < 6 || c == NC || (c - 6 != c_fx && c - 6 != c_cx && c - 6 != c_sk);
    }
    static constexpr int idx(int a, int b) { return a * NL - a * (a - 1) / 2 + (b - a); }  // a <= b
};

// Residual + local Jacobian rows of one observation.  Ju/Jv have NL entries;
// entry NC holds the residual.  A = frame (sensor-rotated [r1 r2 t], row-major 3x3).
template <int MODEL, int IMODE>
CAL_HD void obs_rows(const CamConst& c, const double* A, double X, double Y, double uo, double vo, double* Ju,
                     double* Jv) {
    using L = Local<MODEL, IMODE>;
    const double Px = fma(A[0], X, fma(A[1], Y, A[2]));
    const double Py = fma(A[3], X, fma(A[4], Y, A[5]));
    const double Pz = fma(A[6], X, fma(A[7], Y, A[8]));
    const double iz = rcp_depth(Pz);
    const double mx = Px * iz, my = Py * iz;
    const double x = MODEL == 1 ? mx - c.mx0 : mx;
    const double y = MODEL == 1 ? my - c.my0 : my;
    const double xx = x * x, yy = y * y, xy = x * y;
    const double r2 = xx + yy, r4 = r2 * r2, r6 = r4 * r2;
    const double rad = fma(c.k3, r6, fma(c.k2, r4, fma(c.k1, r2, 1.0)));
    const double a1 = 2.0 * xy, a2 = fma(2.0, xx, r2), a3 = fma(2.0, yy, r2);
    const double xd = fma(x, rad, fma(c.p1, a1, c.p2 * a2));
    const double yd = fma(y, rad, fma(c.p1, a3, c.p2 * a1));
    const double xs = MODEL == 1 ? xd + c.mx0 : xd;
    const double ys = MODEL == 1 ? yd + c.my0 : yd;
    const double u = fma(c.fx, xs, fma(c.sk, ys, c.cx));
    const double v = fma(c.fy, ys, c.cy);
    Ju[L::NC] = u - uo;
    Jv[L::NC] = v - vo;
    // distortion Jacobian d(xd, yd)/d(x, y) (symmetric off-diagonal)
    const double drad = fma(3.0 * c.k3, r4, fma(2.0 * c.k2, r2, c.k1));
    const double dxx = fma(2.0 * xx, drad, rad) + fma(2.0 * c.p1, y, 6.0 * c.p2 * x);
    const double dxy = fma(a1, drad, 2.0 * fma(c.p1, x, c.p2 * y));
    const double dyy = fma(2.0 * yy, drad, rad) + fma(6.0 * c.p1, y, 2.0 * c.p2 * x);
    const double ux = fma(c.fx, dxx, c.sk * dxy), uy = fma(c.fx, dxy, c.sk * dyy);
    const double vx = c.fy * dxy, vy = c.fy * dyy;
    // d(u,v)/dP' = (ux, uy, -(ux mx + uy my)) / Pz ; twist columns: omega = P' x g, nu = g
    const double cu = -fma(ux, mx, uy * my), cv = -fma(vx, mx, vy * my);
    Ju[0] = fma(my, cu, -uy); Ju[1] = fma(-mx, cu, ux); Ju[2] = fma(mx, uy, -my * ux);
    Jv[0] = fma(my, cv, -vy); Jv[1] = fma(-mx, cv, vx); Jv[2] = fma(mx, vy, -my * vx);
    Ju[3] = ux * iz; Ju[4] = uy * iz; Ju[5] = cu * iz;
    Jv[3] = vx * iz; Jv[4] = vy * iz; Jv[5] = cv * iz;
    if (IMODE != INTR_NONE) {
        double* ju = Ju + 6; double* jv = Jv + 6;
        ju[L::c_fx] = xs; jv[L::c_fx] = 0.0;
        ju[L::c_fy] = 0.0; jv[L::c_fy] = ys;
        ju[L::c_cx] = 1.0; jv[L::c_cx] = 0.0;
        ju[L::c_cy] = 0.0; jv[L::c_cy] = 1.0;
        if (IMODE == INTR_SKEW) { ju[4] = ys; jv[4] = 0.0; }
        // d(xd, yd)/d(k1,k2,k3,p1,p2) = [x r2, x r4, x r6, 2xy, r2+2x^2 ; y r2, y r4, y r6, r2+2y^2, 2xy]
        const double fxs = fma(c.fx, x, c.sk * y), fyy = c.fy * y;
        const int d0 = L::c_d0;
        ju[d0 + 0] = fxs * r2; ju[d0 + 1] = fxs * r4; ju[d0 + 2] = fxs * r6;
        ju[d0 + 3] = fma(c.fx, a1, c.sk * a3); ju[d0 + 4] = fma(c.fx, a2, c.sk * a1);
        jv[d0 + 0] = fyy * r2; jv[d0 + 1] = fyy * r4; jv[d0 + 2] = fyy * r6;
        jv[d0 + 3] = c.fy * a3; jv[d0 + 4] = c.fy * a1;
        if (MODEL == 1) {
            // tau_k: rotation of the sensor frame (twist w_k, w_x = (-1,0,0)) plus the moving
            // principal-ray intersection (scheimpflug.h:160-180)
            const int t0 = L::c_tau;
            const double eu0 = c.fx - ux, eu1 = c.sk - uy, ev0 = -vx, ev1 = c.fy - vy;
            ju[t0 + 0] = -Ju[0] + fma(eu0, c.dmx0[0], eu1 * c.dmy0[0]);
            jv[t0 + 0] = -Jv[0] + fma(ev0, c.dmx0[0], ev1 * c.dmy0[0]);
            ju[t0 + 1] = fma(Ju[1], c.wy[1], Ju[2] * c.wy[2]) + fma(eu0, c.dmx0[1], eu1 * c.dmy0[1]);
            jv[t0 + 1] = fma(Jv[1], c.wy[1], Jv[2] * c.wy[2]) + fma(ev0, c.dmx0[1], ev1 * c.dmy0[1]);
        }
    }
}

// Residual only (cost passes).
template <int MODEL>
CAL_HD double obs_ssr(const CamConst& c, const double* A, double X, double Y, double uo, double vo) {
    const double Px = fma(A[0], X, fma(A[1], Y, A[2]));
    const double Py = fma(A[3], X, fma(A[4], Y, A[5]));
    const double Pz = fma(A[6], X, fma(A[7], Y, A[8]));
    const double iz = 1.0 / Pz;
    const double mx = Px * iz, my = Py * iz;
    const double x = MODEL == 1 ? mx - c.mx0 : mx;
    const double y = MODEL == 1 ? my - c.my0 : my;
    const double xx = x * x, yy = y * y, xy = x * y;
    const double r2 = xx + yy, r4 = r2 * r2, r6 = r4 * r2;
    const double rad = fma(c.k3, r6, fma(c.k2, r4, fma(c.k1, r2, 1.0)));
    const double a1 = 2.0 * xy, a2 = fma(2.0, xx, r2), a3 = fma(2.0, yy, r2);
    const double xd = fma(x, rad, fma(c.p1, a1, c.p2 * a2));
    const double yd = fma(y, rad, fma(c.p1, a3, c.p2 * a1));
    const double xs = MODEL == 1 ? xd + c.mx0 : xd;
    const double ys = MODEL == 1 ? yd + c.my0 : yd;
    const double ru = fma(c.fx, xs, fma(c.sk, ys, c.cx)) - uo;
    const double rv = fma(c.fy, ys, c.cy) - vo;
    return fma(ru, ru, rv * rv);
}

// ---------------------------------------------------------------------------
// small fixed-size helpers
// ---------------------------------------------------------------------------
// Eigen toRotationMatrix of a non-normalised quaternion (w,x,y,z) — observationutils.h:20-24
CAL_HD void quat_to_R(const double* q, double* R) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}
CAL_HD void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
CAL_HD void mat3_tmul(const double* A, const double* B, double* C) {  // A^T B
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
CAL_HD void mat3_vec(const double* A, const double* v, double* o) {
    for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
CAL_HD void mat3_tvec(const double* A, const double* v, double* o) {  // A^T v
    for (int i = 0; i < 3; ++i) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}
// [t]x M
CAL_HD void skew_mul(const double* t, const double* M, double* O) {
    for (int j = 0; j < 3; ++j) {
        O[j] = t[1] * M[6 + j] - t[2] * M[3 + j];
        O[3 + j] = t[2] * M[j] - t[0] * M[6 + j];
        O[6 + j] = t[0] * M[3 + j] - t[1] * M[j];
    }
}

// Composite pose of a residual block and the chain-rule maps from the Ceres
// tangent increments (delta_q (3), delta_t (3)) of its pose blocks to the
// sensor-frame twist.  With the QuaternionManifold update R <- R(dq(delta)) R
// (rotation by 2|delta|) and t <- t + delta_t:
//   view-type block  T_v = Rs6^T [[2M, 0], [2 [t_ct]x M, M]]
//       intrinsics: M = I            (pose = c_se3_t itself,  intrinsicresidual.h:22-29)
//       extrinsics: M = R_cr         (c_se3_t = c_se3_r r_se3_t, extrinsicsresidual.h:14-20)
//       bundle    : M = R_gc^T R_bg^T (c_se3_t = g_se3_c^-1 b_se3_g^-1 b_se3_t, bundleresidual.h:15-27)
//   camera-type block (constant per camera)
//       extrinsics: T_c = Rs6^T [[2I, 0], [2 [t_cr]x, I]]
//       bundle    : T_c = Rs6^T [[-2 R_gc^T, 0], [0, -R_gc^T]]
// T matrices are row-major 6x6 mapping (delta_q, delta_t) -> (omega', nu').
struct BlockPose {
    double R[9], t[3];  // composite c_se3_t
    double M[9];        // rotation in front of the view-type block
};

CAL_HD void compose_intrinsics(const double* q, const double* t, BlockPose& bp) {
    quat_to_R(q, bp.R); bp.t[0] = t[0]; bp.t[1] = t[1]; bp.t[2] = t[2];
    for (int i = 0; i < 9; ++i) bp.M[i] = (i % 4 == 0) ? 1.0 : 0.0;
}
CAL_HD void compose_extrinsics(const double* qc, const double* tc, const double* qv, const double* tv, BlockPose& bp) {
    double Rv[9]; quat_to_R(qc, bp.M); quat_to_R(qv, Rv);
    mat3_mul(bp.M, Rv, bp.R);
    double r[3]; mat3_vec(bp.M, tv, r);
    bp.t[0] = r[0] + tc[0]; bp.t[1] = r[1] + tc[1]; bp.t[2] = r[2] + tc[2];
}
CAL_HD void compose_bundle(const double* qb, const double* tb, const double* qg, const double* tg, const double* bTg,
                           BlockPose& bp) {
    double Rbt[9], Rgc[9]; quat_to_R(qb, Rbt); quat_to_R(qg, Rgc);
    // M = R_gc^T R_bg^T
    double Rbg_t[9]; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rbg_t[3 * i + j] = bTg[3 * j + i];
    mat3_tmul(Rgc, Rbg_t, bp.M);
    mat3_mul(bp.M, Rbt, bp.R);
    // t_ct = R_gc^T ( R_bg^T (t_bt - t_bg) - t_gc )
    const double d[3] = {tb[0] - bTg[9], tb[1] - bTg[10], tb[2] - bTg[11]};
    double e[3]; mat3_vec(Rbg_t, d, e);
    e[0] -= tg[0]; e[1] -= tg[1]; e[2] -= tg[2];
    mat3_tvec(Rgc, e, bp.t);
}
// frame A' = Rs^T [r1 r2 t] (row-major 3x3) consumed by obs_rows
CAL_HD void block_frame(const BlockPose& bp, const double* Rs, double* A) {
    const double B[9] = {bp.R[0], bp.R[1], bp.t[0], bp.R[3], bp.R[4], bp.t[1], bp.R[6], bp.R[7], bp.t[2]};
    mat3_tmul(Rs, B, A);
}
CAL_HD void fill_T(const double* TL, const double* BL, const double* BR, double* T) {
    // T = [[TL, 0], [BL, BR]]
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        T[6 * i + j] = TL[3 * i + j]; T[6 * i + 3 + j] = 0.0;
        T[6 * (3 + i) + j] = BL[3 * i + j]; T[6 * (3 + i) + 3 + j] = BR[3 * i + j];
    }
}
CAL_HD void view_transform(const BlockPose& bp, const double* Rs, double* T) {
    double RsM[9]; mat3_tmul(Rs, bp.M, RsM);               // Rs^T M
    double tM[9]; skew_mul(bp.t, bp.M, tM);                // [t_ct]x M
    double RstM[9]; mat3_tmul(Rs, tM, RstM);               // Rs^T [t_ct]x M
    double TL[9], BL[9];
    for (int i = 0; i < 9; ++i) { TL[i] = 2.0 * RsM[i]; BL[i] = 2.0 * RstM[i]; }
    fill_T(TL, BL, RsM, T);
}
CAL_HD void cam_transform_extrinsics(const double* tc, const double* Rs, double* T) {
    double Rst[9]; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rst[3 * i + j] = Rs[3 * j + i];
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double tI[9]; skew_mul(tc, I, tI);
    double BL[9]; mat3_tmul(Rs, tI, BL);
    double TL[9]; for (int i = 0; i < 9; ++i) { TL[i] = 2.0 * Rst[i]; BL[i] *= 2.0; }
    fill_T(TL, BL, Rst, T);
}
CAL_HD void cam_transform_bundle(const double* qg, const double* Rs, double* T) {
    double Rgc[9]; quat_to_R(qg, Rgc);
    double RgcT[9]; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) RgcT[3 * i + j] = Rgc[3 * j + i];
    double B[9]; mat3_tmul(Rs, RgcT, B);  // Rs^T R_gc^T
    double TL[9], BR[9], Z[9];
    for (int i = 0; i < 9; ++i) { TL[i] = -2.0 * B[i]; BR[i] = -B[i]; Z[i] = 0.0; }
    fill_T(TL, Z, BR, T);
}

// HuberLoss::Evaluate with the Corrector's rho'' <= 0 branch:
// returns rho(s) and the block weight rho'(s).
CAL_HD void huber_weight(double delta, double s, double& rho, double& w) {
    if (delta > 0.0 && s > delta * delta) {
        const double r = sqrt(s);
        rho = 2.0 * delta * r - delta * delta;
        w = delta / r; if (w < 2.2250738585072014e-308) w = 2.2250738585072014e-308;
    } else { rho = s; w = 1.0; }
}

// QuaternionManifold::Plus
CAL_HD void quat_plus(const double* q, const double* dl, double* out) {
    const double nd = sqrt(dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
    if (nd == 0.0) { out[0] = q[0]; out[1] = q[1]; out[2] = q[2]; out[3] = q[3]; return; }
    const double sd = sin(nd) / nd;
    const double d0 = cos(nd), d1 = sd * dl[0], d2 = sd * dl[1], d3 = sd * dl[2];
    out[0] = d0 * q[0] - d1 * q[1] - d2 * q[2] - d3 * q[3];
    out[1] = d0 * q[1] + d1 * q[0] + d2 * q[3] - d3 * q[2];
    out[2] = d0 * q[2] - d1 * q[3] + d2 * q[0] + d3 * q[1];
    out[3] = d0 * q[3] + d1 * q[2] - d2 * q[1] + d3 * q[0];
}

// 6x6 Cholesky (lower, in place, row-major) and solve; false if not PD.
CAL_HD bool chol6(double* A) {
    for (int j = 0; j < 6; ++j) {
        double s = A[7 * j];
        for (int k = 0; k < j; ++k) s -= A[6 * j + k] * A[6 * j + k];
        if (!(s > 0.0)) return false;
        const double l = sqrt(s), il = 1.0 / l;
        A[7 * j] = l;
        for (int i = j + 1; i < 6; ++i) {
            double t = A[6 * i + j];
            for (int k = 0; k < j; ++k) t -= A[6 * i + k] * A[6 * j + k];
            A[6 * i + j] = t * il;
        }
    }
    return true;
}
CAL_HD void chol6_solve(const double* L, double* b) {
    for (int i = 0; i < 6; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[6 * i + k] * b[k]; b[i] = s / L[7 * i]; }
    for (int i = 5; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < 6; ++k) s -= L[6 * k + i] * b[k]; b[i] = s / L[7 * i]; }
}

}  // namespace calk
