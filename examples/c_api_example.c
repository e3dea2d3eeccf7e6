/* Minimal C99 client of the C ABI (include/calib_b200.h): optimize_bundle-shaped problem through
 * cal_refine_create / cal_refine_solve, exactly what a cgo / JNI / ctypes binding would call.
 * Build:  gcc -std=c99 -Iinclude examples/c_api_example.c -Lcalibration_b200/_build -lcalib_b200 -lm -o c_api_example
 * Run on a machine with a B200 (there is no CPU fallback: without a device it prints the error and exits 2). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "calib_b200.h"

int main(void) {
    /* one camera looking at a 6 x 5 board from 24 robot poses (identity hand-eye, target 1 m in front) */
    enum { NV = 24, NX = 6, NY = 5, NP = NX * NY };
    static double x[NV * NP], y[NV * NP], u[NV * NP], v[NV * NP], bTg[NV * 12];
    static int64_t off[NV + 1];
    static int32_t cam[NV];
    const double fx = 900, fy = 910, cx = 640, cy = 360;
    for (int k = 0; k < NV; ++k) {
        const double a = 0.25 * sin(0.7 * k), b = 0.2 * cos(1.3 * k);          /* small rotations about x and y */
        const double ca = cos(a), sa = sin(a), cb = cos(b), sb = sin(b);
        /* b_T_g = Ry(b) Rx(a), translation on a small circle; camera = gripper, target at z = 1 in the base frame */
        const double R[9] = {cb, sb * sa, sb * ca, 0, ca, -sa, -sb, cb * sa, cb * ca};
        const double t[3] = {0.05 * cos(0.5 * k), 0.05 * sin(0.5 * k), 0.02 * sin(0.9 * k)};
        memcpy(bTg + 12 * k, R, sizeof R); memcpy(bTg + 12 * k + 9, t, sizeof t);
        off[k] = (int64_t)k * NP; cam[k] = 0;
        for (int i = 0; i < NP; ++i) {
            const double X = 0.03 * (i % NX - 0.5 * (NX - 1)), Y = 0.03 * (i / NX - 0.5 * (NY - 1));
            /* c_T_t = b_T_g^-1 b_T_t with b_T_t = translation (0, 0, 1) */
            const double d[3] = {X - t[0], Y - t[1], 1.0 - t[2]};
            const double P[3] = {R[0] * d[0] + R[3] * d[1] + R[6] * d[2], R[1] * d[0] + R[4] * d[1] + R[7] * d[2], R[2] * d[0] + R[5] * d[1] + R[8] * d[2]};
            x[k * NP + i] = X; y[k * NP + i] = Y;
            u[k * NP + i] = fx * P[0] / P[2] + cx; v[k * NP + i] = fy * P[1] / P[2] + cy;
        }
    }
    off[NV] = (int64_t)NV * NP;
    cal_problem_desc d;
    memset(&d, 0, sizeof d);
    d.kind = CAL_KIND_BUNDLE; d.model = CAL_MODEL_PINHOLE_BC5; d.n_cams = 1; d.n_views = 0;
    d.n_blocks = NV; d.n_obs = NV * NP;
    d.obj_x = x; d.obj_y = y; d.img_u = u; d.img_v = v; d.block_offset = off; d.block_cam = cam; d.block_b_se3_g = bTg;
    d.optimize_intrinsics = 1; d.optimize_hand_eye = 1; d.optimize_target_pose = 1; d.huber_delta = 1.0;
    cal_refine_handle* h = NULL;
    if (cal_refine_create(&d, 0, &h) != CAL_OK) { fprintf(stderr, "cal_refine_create: %s\n", cal_last_error()); return 2; }
    /* parameters in the reference's block order: intrinsics(10) | g_q_c(4) | g_t_c(3) | b_q_t(4) | b_t_t(3), perturbed start */
    double p[24] = {880, 930, 650, 350, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0.005, -0.004, 0.003, 1, 0, 0, 0, 0.01, -0.01, 1.02};
    cal_optim_options o;
    memset(&o, 0, sizeof o);
    o.max_iterations = 100; o.epsilon = 1e-9; o.compute_covariance = 1;
    cal_optim_result r;
    static double cov[24 * 24];
    const cal_status s = cal_refine_solve(h, &o, p, &r, cov);
    cal_refine_destroy(h);
    if (s != CAL_OK) { fprintf(stderr, "cal_refine_solve: %s\n", cal_last_error()); return 2; }
    printf("%s\nfx %.6f fy %.6f cx %.6f cy %.6f (truth %g %g %g %g), sigma(fx) %.3g\n", r.report, p[0], p[1], p[2], p[3], fx, fy, cx, cy,
           r.covariance_ok ? sqrt(cov[0]) : -1.0);
    return r.success && fabs(p[0] - fx) < 1e-5 && fabs(p[3] - cy) < 1e-5 ? 0 : 1;
}
