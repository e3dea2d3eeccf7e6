/* calib_b200 — C ABI of the B200-native refinement hot path.
 *
 * Drop-in boundary for the data-parallel refinement path of
 * VitalyVorobyev/calibration.  The reference has no FFI seam of its own: the
 * path sits behind C++ functions of the static library calib_estimation_optim
 * (SURVEY §8b).  Each entry point below names the reference interface it
 * replaces; INTEGRATION.md shows the C++ adapter that re-exposes the exact
 * calib::optimize_* / calib::estimate_homography signatures on top of it.
 *
 * Conventions
 *   - plain pointers and sizes; all floating point data is FP64;
 *   - host pointers unless a name ends in _dev;
 *   - the caller owns every buffer passed in; handles own their device copies;
 *   - every call returns a cal_status; cal_last_error() gives the message of
 *     the last failure on the calling thread;
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point returns CAL_ERR_CUDA.
 *
 * Parameter vector `x` (ambient coordinates) uses the reference's own block
 * order, which is also its covariance row/column order:
 *   intrinsics  [intr(P)] [quat(4) x n_views] [tran(3) x n_views]
 *               (IntrinsicBlocks::get_param_blocks, src/estimation/optim/intrinsics.cpp:34-50)
 *   extrinsics  [intr(P) x n_cams] [c_q_r x n_cams] [c_t_r x n_cams] [r_q_t x n_views] [r_t_t x n_views]
 *               (ExtrinsicBlocks::get_param_blocks, src/estimation/optim/extrinsics.cpp:50-69)
 *   bundle      [intr(P) x n_cams] [g_q_c x n_cams] [g_t_c x n_cams] [b_q_t] [b_t_t]
 *               (BundleBlocks::get_param_blocks, src/estimation/optim/bundle.cpp:48-68)
 *   hand-eye    [quat(4)] [tran(3)]   (HandeyeBlocks, src/estimation/optim/handeye.cpp:17-43)
 * quaternions are (w, x, y, z) (observationutils.h:43-48); intr is
 * [fx, fy, cx, cy, skew, k1, k2, k3, p1, p2 (, tau_x, tau_y)] (models/pinhole.h:125-146,
 * models/scheimpflug.h:241-259), P = 10 or 12.
 */
#ifndef CALIB_B200_H
#define CALIB_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum cal_status {
    CAL_OK = 0,
    CAL_ERR_INVALID_ARGUMENT = 1, /* adapter rethrows std::invalid_argument */
    CAL_ERR_RUNTIME = 2,          /* adapter rethrows std::runtime_error    */
    CAL_ERR_CUDA = 3,             /* no device / CUDA failure: there is no CPU path */
    CAL_ERR_COMM = 4
} cal_status;

enum { CAL_KIND_INTRINSICS = 0, CAL_KIND_EXTRINSICS = 1, CAL_KIND_BUNDLE = 2 };
enum { CAL_MODEL_PINHOLE_BC5 = 0, CAL_MODEL_SCHEIMPFLUG_BC5 = 1 };
enum { CAL_TERM_CONVERGENCE = 0, CAL_TERM_NO_CONVERGENCE = 1, CAL_TERM_FAILURE = 2 };

/* One reprojection refinement problem in SoA / CSR form.  Replaces the
 * std::vector<PlanarView> / MulticamPlanarView / BundleObservation inputs of
 * optimize_intrinsics (optim/intrinsics.h:35-39), optimize_extrinsics
 * (optim/extrinsics.h:29-34) and optimize_bundle (optim/bundle.h:58-63); the
 * flags are IntrinsicsOptimOptions / ExtrinsicOptions / BundleOptions. */
typedef struct cal_problem_desc {
    int32_t kind;
    int32_t model;
    int32_t n_cams;
    int32_t n_views;     /* per-view pose blocks (intrinsics, extrinsics); 0 for bundle */
    int64_t n_blocks;    /* residual blocks: non-empty (view, camera) pairs */
    int64_t n_obs;
    const double* obj_x; /* PlanarObservation::object_xy / image_uv (linear/planarpose.h:22-26), SoA */
    const double* obj_y;
    const double* img_u;
    const double* img_v;
    const int64_t* block_offset; /* CSR into the observation arrays, n_blocks + 1 */
    const int32_t* block_cam;
    const int32_t* block_view;   /* index of the per-view pose block (ignored for bundle) */
    const double* block_b_se3_g; /* bundle: BundleObservation::b_se3_g, [n_blocks][12] = R row-major, t */
    int32_t optimize_intrinsics; /* extrinsics / bundle; always on for the intrinsics kind */
    int32_t optimize_skew;
    int32_t optimize_extrinsics;  /* extrinsics kind */
    int32_t optimize_target_pose; /* bundle kind */
    int32_t optimize_hand_eye;    /* bundle kind */
    int32_t view_base;            /* multi-GPU shards of the per-view kinds: global index of this shard's view 0
                                     (the gauge fixes GLOBAL view 0, extrinsics.cpp:133-139); 0 otherwise */
    double huber_delta;           /* OptimOptions::huber_delta; <= 0 disables the loss */
    /* Optional shared-board form (board_n > 0): every view of a calibration observes the same planar target, so
     * object_xy repeats from view to view.  Every residual block then has exactly board_n observations, in board
     * order; observation k of a block has object_xy = (board_x[k], board_y[k]); obj_x / obj_y are ignored and may
     * be NULL.  Halves the bytes that cross PCIe at create (16 instead of 32 per observation); the device layout
     * and every result are identical to the per-observation form.  board_n = 0: per-observation obj_x / obj_y. */
    const double* board_x;
    const double* board_y;
    int32_t board_n;
    int32_t n_views_total;        /* multi-GPU shards of the intrinsics kind: views of the WHOLE problem (the reference's
                                     "at least 4 views" check, intrinsics.cpp:92-96, applies to it); 0: n_views */
} cal_problem_desc;
/* Limits of the per-view kinds (intrinsics, extrinsics): the shared block — intrinsics of all cameras and, for the extrinsics
 * kind, the 6-dof poses of cameras 1 .. n_cams - 1 — may have at most 175 tangent columns (the Schur-complement kernel holds
 * an (n_s + 1)^2 tile set per CTA): 12 pinhole or 10 Scheimpflug cameras with every block free; cal_refine_create returns
 * CAL_ERR_INVALID_ARGUMENT beyond that.  A (view, camera) pair may appear in at most one residual block. */

/* calib::OptimOptions (optim/optimize.h:24-33) */
typedef struct cal_optim_options {
    int32_t optimizer;        /* OptimizerType; accepted and ignored: all four solve the same damped system */
    int32_t max_iterations;
    double epsilon;
    int32_t compute_covariance;
    int32_t verbose;
    int32_t num_threads;      /* unused on the device path */
    int32_t reserved;
} cal_optim_options;

/* calib::OptimResult (optim/optimize.h:35-40) plus iteration accounting */
typedef struct cal_optim_result {
    int32_t success;          /* termination == CONVERGENCE (detail/ceresutils.h:42) */
    int32_t iterations;
    int32_t num_jac_evals;    /* fused residual + Jacobian + J^T J passes */
    int32_t num_cost_evals;   /* residual-only passes */
    int32_t termination;
    int32_t covariance_ok;
    double initial_cost;
    double final_cost;
    char report[256];         /* same text as ceres::Solver::Summary::BriefReport() */
} cal_optim_result;

typedef struct cal_refine_handle cal_refine_handle;

const char* cal_last_error(void);
/* number of CUDA devices visible (0 when there is none) */
int cal_device_count(void);

/* Replaces the problem builders build_problem / set_residual_blocks
 * (optim/intrinsics.cpp:63-90, optim/extrinsics.cpp:87-150, optim/bundle.cpp:84-133):
 * validates like the reference (its std::invalid_argument cases), uploads the
 * observations once and lays them out for the kernels.  device = CUDA ordinal. */
cal_status cal_refine_create(const cal_problem_desc* desc, int device, cal_refine_handle** out);
void cal_refine_destroy(cal_refine_handle* h);
int64_t cal_refine_param_count(const cal_refine_handle* h);   /* ambient */
int64_t cal_refine_tangent_count(const cal_refine_handle* h); /* free tangent dims */

/* One fused residual + Jacobian + J^T J pass (what Ceres does per iteration
 * through ResidualBlock::Evaluate on the functors of src/estimation/residuals/
 * + normal-equation formation).  cost = 1/2 sum rho(s_b).  g (n_tan) and
 * H (n_tan x n_tan, row-major) are the loss-corrected tangent-space J^T r and
 * J^T J in the block order of x with constant blocks removed; either may be
 * NULL.  Dense H is only produced while n_tan <= 8192. */
cal_status cal_refine_eval(cal_refine_handle* h, const double* x, double* cost, double* g, double* H);
/* Residual-only pass: cost and, if ssr != NULL, the per-residual-block sum of squares. */
cal_status cal_refine_cost(cal_refine_handle* h, const double* x, double* cost, double* block_ssr);
/* Timed device-resident passes for benchmarking: runs `reps` fused passes on
 * parameters already on the device.  ms_total is the CUDA-event time of the
 * whole region on the handle's stream, ms_k1 the summed event time of the
 * dominant kernel (K1, or the residual-only kernel when jacobian == 0). */
cal_status cal_refine_bench_pass(cal_refine_handle* h, const double* x, int reps, int jacobian, float* ms_total,
                                 float* ms_k1, double* cost);
/* Of the last cal_refine_bench_pass: device time (ms) of the first set-up kernel, and of [reduction (+ all-reduce) of one pass +
 * set-up of the next] averaged over the repetitions: what a pass costs besides K1. */
cal_status cal_refine_bench_breakdown(const cal_refine_handle* h, float* ms_first_setup, float* ms_reduce_plus_setup);
/* kernels launched through this handle so far */
int64_t cal_refine_launch_count(const cal_refine_handle* h);
/* layout facts for the roofline arithmetic: segments, 32-segment tiles, bytes of
 * the tile-transposed observation store, entries of the per-segment local
 * system and the number of K1 passes */
cal_status cal_refine_layout_info(const cal_refine_handle* h, int64_t* n_segments, int64_t* n_tiles, int64_t* obs_bytes,
                                  int32_t* local_entries, int32_t* k1_passes);
/* sustained FP64 FMA rate of the device (FMA-chain microbenchmark), TFLOP/s */
cal_status cal_fp64_peak_tflops(int device, double* tflops);

/* Per-view reprojection diagnostics at x: block_rms[b] = sqrt(sum r^2 / (2 points)) of residual block b
 * (view_errors, src/estimation/optim/intrinsicssemidlt.cpp:137-151) and the measurement-weighted global
 * RMS (src/pipeline/reports/intrinsics.cpp:12-31).  One residual-only pass; either output may be NULL. */
cal_status cal_refine_view_errors(cal_refine_handle* h, const double* x, double* block_rms, double* global_rms);

/* Replaces solve_problem (detail/ceresutils.h:27-43) + compute_covariance
 * (detail/ceresutils.h:69-126): Levenberg–Marquardt with Ceres 2.2 semantics
 * on the host, every O(observations) pass on the device.  x is updated in
 * place; cov may be NULL, otherwise n_amb x n_amb row-major. */
cal_status cal_refine_solve(cal_refine_handle* h, const cal_optim_options* opts, double* x_inout,
                            cal_optim_result* result, double* cov);

/* Page-locked host memory for staging large inputs (copies from it run at full PCIe speed and overlap with kernels).
 * Blocks come from a process-wide pool that is never shrunk: a long-running caller pays the page-locking once. */
cal_status cal_host_borrow(size_t bytes, void** out);
void cal_host_return(void* ptr);

/* Optional multi-GPU sharding (one process per GPU).  Each rank creates its
 * handle over its own shard of the residual blocks (all ranks pass identical
 * n_cams and shared parameters) and attaches a communicator; the per-camera
 * normal-equation blocks are then summed with ncclAllReduce (FP64) after the
 * reductions of every pass, and the host LM runs replicated on all ranks.
 * unique_id is the 128-byte ncclUniqueId created by rank 0 (cal_comm_unique_id)
 * and distributed by the caller; the communicator is a process-lifetime object
 * that can serve any number of handles. */
typedef struct cal_comm cal_comm;
cal_status cal_comm_unique_id(uint8_t out128[128]);
cal_status cal_comm_create(const uint8_t unique_id[128], int rank, int world_size, int device, cal_comm** out);
void cal_comm_destroy(cal_comm* c);
cal_status cal_refine_attach_comm(cal_refine_handle* h, cal_comm* c);
/* Optional NVLink peer-memory path for that all-reduce (blocks of <= 4096 doubles): every rank exports its
 * receive region as a 64-byte CUDA IPC handle (cal_comm_peer_export), the caller all-gathers the handles in
 * rank order and passes them to cal_comm_peer_enable.  The all-reduce then is ONE kernel — remote stores of
 * the local block into every peer's slot over NVLink, a system-scope flag, a rank-ordered (bitwise
 * reproducible) local sum — instead of an NCCL call; NCCL remains the path for larger payloads.
 * cal_comm_allreduce_test all-reduces a host vector through either path (self-test). */
cal_status cal_comm_peer_export(cal_comm* c, uint8_t handle_out[64]);
cal_status cal_comm_peer_enable(cal_comm* c, const uint8_t* handles /* [world_size][64] */);
void cal_comm_peer_disable(cal_comm* c); /* back to NCCL */
cal_status cal_comm_allreduce_test(cal_comm* c, double* host_inout, int32_t n, int use_peer);

/* ---- AX = XB hand-eye refinement: optimize_handeye (optim/handeye.h:40-43,
 * src/estimation/optim/handeye.cpp:45-78) over MotionPairs (linear/handeye.h:29-32). */
typedef struct cal_axxb_desc {
    int64_t n_pairs;
    const double* rot_a; /* [n_pairs][9] row-major */
    const double* rot_b;
    const double* tra_a; /* [n_pairs][3] */
    const double* tra_b;
    double huber_delta;
} cal_axxb_desc;
typedef struct cal_axxb_handle cal_axxb_handle;
cal_status cal_axxb_create(const cal_axxb_desc* desc, int device, cal_axxb_handle** out);
/* The same problem straight from the poses, as optimize_handeye(base_se3_gripper, camera_se3_target, ..)
 * receives them ([n_poses][12] = R row-major then t, host or device memory): build_all_pairs
 * (src/estimation/linear/handeyedlt.cpp:51-81; optimize_handeye uses min_angle_deg = 0.5,
 * reject_axis_parallel = true, axis_parallel_eps = 1e-3) runs on the device and the n (n - 1) / 2
 * motion pairs are formed on the fly in every pass instead of being materialised (192 B each).
 * CAL_ERR_RUNTIME mirrors the std::runtime_error of inconsistent sizes / no valid pair. */
cal_status cal_axxb_create_from_poses(int64_t n_poses, const double* base_se3_gripper, const double* cam_se3_target,
                                      double min_angle_deg, int reject_axis_parallel, double axis_parallel_eps,
                                      double huber_delta, int device, cal_axxb_handle** out, int64_t* n_pairs_kept);
void cal_axxb_destroy(cal_axxb_handle* h);
cal_status cal_axxb_eval(cal_axxb_handle* h, const double* x7, double* cost, double* g6, double* H36);
cal_status cal_axxb_solve(cal_axxb_handle* h, const cal_optim_options* opts, double* x7_inout,
                          cal_optim_result* result, double* cov49);
/* Shard a handle created from poses over the ranks of a communicator (cal_comm_create, below): each rank evaluates a
 * contiguous slice of the pair tiles, the 6x6 + 6 + 1 sums are all-reduced after every pass.  NULL detaches. */
cal_status cal_axxb_attach_comm(cal_axxb_handle* h, cal_comm* comm);
/* Benchmark hooks (no reference counterpart): `reps` residual + Jacobian passes back to back, timed with CUDA events on
 * the handle's stream; kernels launched by the handle so far. */
cal_status cal_axxb_bench_pass(cal_axxb_handle* h, const double* x7, int reps, float* ms_total);
int64_t cal_axxb_launch_count(const cal_axxb_handle* h);

/* ---- batched RANSAC homography: estimate_homography(view, RansacOptions)
 * (linear/homography.h:22-24) = ransac<HomographyEstimator> (common/ransac.h:121-194). */
typedef struct cal_ransac_options { /* calib::RansacOptions (common/ransac.h:22-29) */
    int32_t max_iters;
    int32_t min_inliers;
    double thresh;
    double confidence;
    uint64_t seed;
    int32_t refit_on_inliers;
    int32_t reserved;
} cal_ransac_options;
typedef struct cal_ransac_result { /* HomographyResult (linear/homography.h:15-20) + RansacResult fields */
    int32_t success;
    int32_t iters;
    int32_t n_inliers;
    int32_t iters_run;
    double hmtx[9];
    double inlier_rms;
    double symmetric_rms_px;
    double min_margin; /* unused by the device path (kept for layout parity with the test oracle) */
} cal_ransac_result;
/* n_problems independent problems of n correspondences each, arrays laid out
 * [problem][n].  seed of problem p = opts->seed + p when seed_per_problem != 0
 * (else opts->seed for all).  The minimal-sample stream is the libstdc++
 * std::sample / std::mt19937_64 stream of the reference (ransac.h:135,144-145),
 * generated on the device.  inlier_mask ([problem][n] bytes) may be NULL. */
cal_status cal_ransac_homography_batch(int64_t n_problems, int32_t n, const double* x, const double* y,
                                       const double* u, const double* v, const cal_ransac_options* opts,
                                       int seed_per_problem, int device, cal_ransac_result* results,
                                       uint8_t* inlier_mask);
/* The same batch split by problem over several devices of the box, one host thread per device, no communication
 * (independent-problem batches); devices[d] takes a contiguous slice and its seeds continue the global
 * problem index, so every result is what one device returns for the whole batch.  Device ids may repeat. */
cal_status cal_ransac_homography_batch_multi(int64_t n_problems, int32_t n, const double* x, const double* y,
                                             const double* u, const double* v, const cal_ransac_options* opts,
                                             int seed_per_problem, int32_t n_devices, const int32_t* devices,
                                             cal_ransac_result* results, uint8_t* inlier_mask);
/* Device-resident variant used by the benchmark: pointers are device memory,
 * results stay on the device; returns the CUDA-event time of the kernel. */
cal_status cal_ransac_homography_batch_dev(int64_t n_problems, int32_t n, const double* x_dev, const double* y_dev,
                                           const double* u_dev, const double* v_dev, const cal_ransac_options* opts,
                                           int seed_per_problem, cal_ransac_result* results_dev,
                                           uint8_t* inlier_mask_dev, float* ms);

/* ---- batched RANSAC plane fit: fit_plane_ransac(pts, RansacOptions) (linear/planefit.h:23-24,
 * src/estimation/linear/planefit.cpp:86-104) = ransac<PlaneRansacEstimator> (planefit.cpp:9-62) with
 * fit_plane_svd (:66-84) as the refit — the batched RANSAC kernel generalised to the reference's second
 * estimator. */
typedef struct cal_plane_ransac_result { /* PlaneRansacResult (linear/planefit.h:14-19) + RansacResult fields */
    int32_t success;
    int32_t iters;
    int32_t n_inliers;
    int32_t iters_run;
    double plane[4];   /* (n, d): n . p + d = 0, |n| = 1; all zero when !success (planefit.h:16) */
    double inlier_rms;
    double min_margin; /* unused by the device path (layout parity with the test oracle) */
} cal_plane_ransac_result;
/* n_problems independent point sets of n points each, arrays laid out [problem][n]; seeds, sample stream
 * and inlier_mask as cal_ransac_homography_batch (k_min_samples = 3).  A plane that comes from the refit
 * (fit_plane_svd) has the sign that makes its largest normal component positive — the reference leaves
 * that sign to Eigen::JacobiSVD and its tests align it (tests/unit/planefit_test.cpp:18-20). */
cal_status cal_ransac_plane_batch(int64_t n_problems, int32_t n, const double* x, const double* y, const double* z,
                                  const cal_ransac_options* opts, int seed_per_problem, int device,
                                  cal_plane_ransac_result* results, uint8_t* inlier_mask);
/* Device-resident variant: pointers are device memory, results stay on the device; returns the
 * CUDA-event time of the kernel. */
cal_status cal_ransac_plane_batch_dev(int64_t n_problems, int32_t n, const double* x_dev, const double* y_dev,
                                      const double* z_dev, const cal_ransac_options* opts, int seed_per_problem,
                                      cal_plane_ransac_result* results_dev, uint8_t* inlier_mask_dev, float* ms);

/* ---- linear seeding stage that feeds the refinement, batched over views.
 * Views are given as SoA observations + CSR offsets, exactly like the residual blocks of
 * cal_problem_desc (view k = observations [view_offset[k], view_offset[k+1]) seen by camera
 * view_cam[k]).  The observation arrays may be host or device pointers. */
typedef struct cal_seed_options { /* IntrinsicsEstimOptions.bounds (linear/intrinsics.h:26-30), CalibrationBounds (models/camera_matrix.h:50-72) */
    int32_t use_bounds; /* 0: bounds = std::nullopt */
    int32_t reserved;
    double fx_min, fx_max, fy_min, fy_max, cx_min, cx_max, cy_min, cy_max, skew_min, skew_max;
} cal_seed_options;
/* estimate_intrinsics(views, opts) for every camera, homography_ransac = nullopt
 * (src/estimation/linear/intrinsicsdlt.cpp:101-145): per-view DLT homography (h33 = 1) and
 * symmetric_rms_px, Zhang's closed form for K (zhang.cpp:183-208), sanitize_intrinsics, and
 * pose_from_homography (posefromhomography.cpp:12-67) per view.
 * kmtx: [n_cams][5] = fx, fy, cx, cy, skew; cam_success[n_cams]; per view (any may be NULL):
 * view_success, hmtx [9] row-major, sym_rms, poses [12] = R row-major then t (identity when failed). */
cal_status cal_seed_intrinsics(int64_t n_views, const int64_t* view_offset, const int32_t* view_cam, const double* x,
                               const double* y, const double* u, const double* v, int32_t n_cams,
                               const cal_seed_options* opts, int device, double* kmtx, int32_t* cam_success,
                               int32_t* view_success, double* hmtx, double* sym_rms, double* poses);
/* The same with IntrinsicsEstimOptions::homography_ransac set (intrinsicsdlt.cpp:50-64): every view's homography
 * comes from ransac<HomographyEstimator> with the options' seed (the batched RANSAC kernel: one launch when the
 * views have equal size, else one launch per distinct view size), hmtx = model / h33, sym_rms over the inliers;
 * inlier_mask (one byte per observation, in the observations' own order) may be NULL.
 * ransac == NULL is cal_seed_intrinsics.  (cal_ransac_options is declared above.) */
cal_status cal_seed_intrinsics_ransac(int64_t n_views, const int64_t* view_offset, const int32_t* view_cam, const double* x,
                                      const double* y, const double* u, const double* v, int32_t n_cams,
                                      const cal_seed_options* opts, const cal_ransac_options* ransac, int device, double* kmtx,
                                      int32_t* cam_success, int32_t* view_success, double* hmtx, double* sym_rms, double* poses,
                                      uint8_t* inlier_mask);
/* estimate_planar_pose(view, CameraMatrix) for every view (src/estimation/linear/planarpose_linear.cpp:54-76
 * with pose_from_homography_normalized :17-52); kmtx as above; identity for views with < 4 points
 * or a failed DLT (view_success, optional, tells which). */
cal_status cal_seed_planar_poses(int64_t n_views, const int64_t* view_offset, const int32_t* view_cam, const double* x,
                                 const double* y, const double* u, const double* v, int32_t n_cams, const double* kmtx,
                                 int device, double* poses, int32_t* view_success);

/* ---- columnar observation store: the SoA + CSR arrays of cal_problem_desc / cal_seed_*
 * in one file, mmap-ed read-only (zero parsing, zero copies before the H2D transfer), and a streaming
 * converter from the reference's PlanarDetections JSON (schemas/calib_dataset.schema.json,
 * include/calib/pipeline/dataset.h:15-39) that replaces the nlohmann DOM + collect_planar_views /
 * make_planar_view packing (src/pipeline/facades/intrinsics.cpp:38-59, detail/planar_utils.cpp:45-52). */
typedef struct cal_dataset {
    int64_t n_views;
    int64_t n_obs;
    int32_t n_cams;
    int32_t pinned;             /* the mapping is page-locked (cudaHostRegister) */
    const int64_t* view_offset; /* [n_views + 1] */
    const int32_t* view_cam;    /* [n_views] */
    const double* obj_x;        /* local_x */
    const double* obj_y;        /* local_y */
    const double* img_u;        /* x (pixels) */
    const double* img_v;        /* y (pixels) */
    void* impl;
} cal_dataset;
cal_status cal_dataset_write(const char* path, int64_t n_views, int32_t n_cams, const int64_t* view_offset,
                             const int32_t* view_cam, const double* x, const double* y, const double* u, const double* v);
cal_status cal_dataset_open(const char* path, int pin, cal_dataset* out);
void cal_dataset_close(cal_dataset* d);
/* One PlanarDetections JSON document per camera -> one columnar file.  Images with fewer than
 * min_corners_per_view points are dropped (collect_planar_views, facades/intrinsics.cpp:45-47).
 * CAL_ERR_INVALID_ARGUMENT with the byte offset for malformed documents. */
cal_status cal_dataset_from_planar_json(const char* const* json_paths, int32_t n_cams, int32_t min_corners_per_view,
                                        const char* out_path, int64_t* n_views_out, int64_t* n_obs_out);

#ifdef __cplusplus
}
#endif
#endif /* CALIB_B200_H */
