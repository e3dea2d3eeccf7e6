/* ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp header).
 *
 * C interface of the CPU restatement, loaded with ctypes by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs.  The struct layouts deliberately equal the product's
 * include/calib_b200.h so one ctypes definition drives both sides; the oracle
 * does NOT include the product header and the product never links this file.
 */
#ifndef ORACLE_API_H
#define ORACLE_API_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_KIND_INTRINSICS = 0, ORC_KIND_EXTRINSICS = 1, ORC_KIND_BUNDLE = 2 };
enum { ORC_MODEL_PINHOLE_BC5 = 0, ORC_MODEL_SCHEIMPFLUG_BC5 = 1 };

typedef struct orc_problem_desc {
    int32_t kind;
    int32_t model;
    int32_t n_cams;
    int32_t n_views;      /* per-view pose blocks (intrinsics, extrinsics); 0 for bundle */
    int64_t n_blocks;     /* residual blocks = (view, camera) pairs with observations */
    int64_t n_obs;
    const double* obj_x;  /* SoA observations, length n_obs, grouped by residual block */
    const double* obj_y;
    const double* img_u;
    const double* img_v;
    const int64_t* block_offset; /* CSR, n_blocks + 1 */
    const int32_t* block_cam;    /* camera of each residual block */
    const int32_t* block_view;   /* per-view pose block of each residual block (unused for bundle) */
    const double* block_b_se3_g; /* bundle only: [n_blocks][12] = R row-major (9) then t (3) */
    int32_t optimize_intrinsics;
    int32_t optimize_skew;
    int32_t optimize_extrinsics;
    int32_t optimize_target_pose;
    int32_t optimize_hand_eye;
    int32_t reserved;
    double huber_delta;
} orc_problem_desc;

typedef struct orc_optim_options {
    int32_t optimizer;       /* accepted and ignored (all four solve the same damped system) */
    int32_t max_iterations;
    double epsilon;          /* function = gradient = parameter tolerance (ceresutils.h:32-34) */
    int32_t compute_covariance;
    int32_t verbose;
    int32_t num_threads;     /* 0 = all */
    int32_t reserved;
} orc_optim_options;

typedef struct orc_optim_result {
    int32_t success;         /* termination == CONVERGENCE (ceresutils.h:42) */
    int32_t iterations;      /* LM iterations run (successful + unsuccessful) */
    int32_t num_jac_evals;
    int32_t num_cost_evals;
    int32_t termination;     /* 0 convergence, 1 no_convergence, 2 failure */
    int32_t covariance_ok;
    double initial_cost;
    double final_cost;
    char report[256];
} orc_optim_result;

int64_t orc_param_count(const orc_problem_desc* d);
int64_t orc_tangent_count(const orc_problem_desc* d);
/* cost = 1/2 sum rho(s_b); g = J^T r and H = J^T J in tangent space with the
 * loss applied (dense, canonical tangent order, row-major n_tan x n_tan).
 * g / H may be NULL.  mode 0 = forward-mode duals (what Ceres autodiff does). */
int orc_refine_eval(const orc_problem_desc* d, const double* x, double* cost, double* g, double* H,
                    int num_threads);
/* Restated Ceres 2.2 trust-region LM. cov may be NULL;
 * otherwise n_amb x n_amb row-major. force_dense != 0 solves the full dense
 * normal equations instead of eliminating the per-view pose blocks. */
int orc_refine_solve(const orc_problem_desc* d, const orc_optim_options* o, double* x_inout,
                     orc_optim_result* res, double* cov, int force_dense);
/* Per-block sum of squared residuals (unweighted), length n_blocks. */
int orc_block_ssr(const orc_problem_desc* d, const double* x, double* ssr, int num_threads);

/* model-level known answers (scheimpflug_test.cpp:11-51) */
void orc_project(int model, const double* intr, const double* P, double* uv);

/* AX = XB (optimize_handeye) */
typedef struct orc_axxb_desc {
    int64_t n_pairs;
    const double* rot_a; /* [n_pairs][9] row-major */
    const double* rot_b;
    const double* tra_a; /* [n_pairs][3] */
    const double* tra_b;
    double huber_delta;
} orc_axxb_desc;
int orc_axxb_eval(const orc_axxb_desc* d, const double* x7, double* cost, double* g6, double* H36,
                  int num_threads);
int orc_axxb_solve(const orc_axxb_desc* d, const orc_optim_options* o, double* x7_inout,
                   orc_optim_result* res, double* cov49);
/* build_all_pairs (linear/handeyedlt.cpp:51-81). poses are [n][12] (R row-major, t).
 * Returns the number of kept pairs; out arrays sized for n(n-1)/2 pairs (may be NULL to count). */
int64_t orc_build_all_pairs(int64_t n, const double* base_se3_gripper, const double* cam_se3_target,
                            double min_angle_deg, double* rot_a, double* rot_b, double* tra_a,
                            double* tra_b);

/* RANSAC homography (common/ransac.h:121-194 + linear/homographyestimator.cpp) */
typedef struct orc_ransac_options {
    int32_t max_iters;
    int32_t min_inliers;
    double thresh;
    double confidence;
    uint64_t seed;
    int32_t refit_on_inliers;
    int32_t reserved;
} orc_ransac_options;
typedef struct orc_ransac_result {
    int32_t success;
    int32_t iters;       /* best.iters */
    int32_t n_inliers;
    int32_t iters_run;   /* loop trip count */
    double hmtx[9];
    double inlier_rms;
    double symmetric_rms_px; /* optim/homography.cpp:18-28 (sums roots) */
    double min_margin;   /* min |r - thresh| over every scored (hypothesis, point) */
} orc_ransac_result;
/* libstdc++-13 std::sample(0..n-1, k=4) stream for std::mt19937_64(seed): writes iters*4 indices */
void orc_sample_stream(uint64_t seed, int32_t n, int32_t iters, int32_t* out);
/* the same stream drawn with the real std::sample of this toolchain (known-answer check) */
void orc_sample_stream_libstdcxx(uint64_t seed, int32_t n, int32_t iters, int32_t* out);
/* one problem; sample_idx may be NULL (drawn internally from opts.seed).
 * inlier_mask: n bytes. */
int orc_ransac_homography(int32_t n, const double* x, const double* y, const double* u,
                          const double* v, const orc_ransac_options* o, const int32_t* sample_idx,
                          orc_ransac_result* res, uint8_t* inlier_mask);
/* batch over problems of n correspondences each, laid out [problem][n]; OpenMP over problems */
int orc_ransac_homography_batch(int64_t n_problems, int32_t n, const double* x, const double* y,
                                const double* u, const double* v, const orc_ransac_options* o,
                                int seed_per_problem, orc_ransac_result* res, uint8_t* inlier_mask,
                                int num_threads);
/* plain DLT on all points (estimate_homography without RANSAC, optim/homography.cpp:30-43) */
int orc_homography_dlt(int32_t n, const double* x, const double* y, const double* u,
                       const double* v, double* hmtx);


/* RANSAC plane fit (fit_plane_ransac, src/estimation/linear/planefit.cpp:9-104): the same loop, k = 3 */
typedef struct orc_plane_result {
    int32_t success;
    int32_t iters;
    int32_t n_inliers;
    int32_t iters_run;
    double plane[4];     /* (n, d), |n| = 1; Zero when !success (planefit.h:16) */
    double inlier_rms;
    double min_margin;
} orc_plane_result;
/* project_to_so3 / log_so3 (common/se3_utils.h:10-42), used by the pair filter of build_all_pairs */
void orc_project_to_so3(const double* R9, double* out9);
void orc_log_so3(const double* R9, double* w3);
void orc_ref_plane_data(double* plane, double* xyz /* [140][3] */);
void orc_sample_stream_k(uint64_t seed, int32_t n, int32_t k, int32_t iters, int32_t* out);
void orc_sample_stream_k_libstdcxx(uint64_t seed, int32_t n, int32_t k, int32_t iters, int32_t* out);
/* fit_plane_svd (planefit.cpp:66-84); returns 1 for n < 3 (the reference throws) */
int orc_fit_plane_svd(int32_t n, const double* x, const double* y, const double* z, double* plane);
int orc_ransac_plane(int32_t n, const double* x, const double* y, const double* z, const orc_ransac_options* o,
                     orc_plane_result* res, uint8_t* inlier_mask);
int orc_ransac_plane_batch(int64_t n_problems, int32_t n, const double* x, const double* y, const double* z,
                           const orc_ransac_options* o, int seed_per_problem, orc_plane_result* res,
                           uint8_t* inlier_mask, int num_threads);

#ifdef __cplusplus
}
#endif
#endif
