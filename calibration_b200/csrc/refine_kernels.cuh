// Device-side data layout and kernel launchers of the refinement hot path.
// See DESIGN.md §3-§4 for the layout in HBM and the roofline of each kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <vector>

#include "comm_peer.cuh"
#include "k1_math.cuh"

namespace calk {

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the (function, device) pair: a handle on a second device
// of the same process must set it again.  One bit per device ordinal; setting it twice from two threads is harmless.
struct PerDeviceOnce {
    std::atomic<uint64_t> mask{0};
    bool first() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return true;
        const uint64_t bit = 1ull << (dev & 63);
        if (mask.load(std::memory_order_acquire) & bit) return false;
        mask.fetch_or(bit, std::memory_order_acq_rel);
        return true;
    }
};

// Observations are stored tile-transposed: residual blocks are cut into
// segments of <= seg_len corners, 32 segments form a tile (one warp), and the
// tile stores corner k of its 32 segments contiguously:
//     obs[(tile_off[t] + k) * 128 + comp * 32 + lane],  comp in {x, y, u, v}
// so the lane-per-segment kernels read 256-byte fully coalesced rows.
struct DevLayout {
    int64_t n_seg = 0;        // padded to a multiple of 32
    int64_t n_tiles = 0;
    int64_t n_blk = 0;        // device blocks: sorted by camera, camera groups padded to 32
    int64_t n_slices = 0;     // total k-slices (each 128 doubles)
    int32_t one_seg_per_blk = 0;  // every real block has exactly one segment
    int32_t fused = 0;            // segment s == device block s (padding blocks included): K1 runs its fused epilogue
    double* obs = nullptr;
    int64_t* tile_off = nullptr;  // [n_tiles]
    int32_t* tile_depth = nullptr;
    int32_t* seg_len = nullptr;   // [n_seg]
    int32_t* seg_blk = nullptr;   // [n_seg] device block id
    int32_t* seg_cam = nullptr;   // [n_seg]
    int32_t* blk_cam = nullptr;   // [n_blk]
    int32_t* blk_view = nullptr;  // [n_blk] (-1: none / padding)
    int64_t* blk_orig = nullptr;  // [n_blk] original residual block id, -1 for padding
    int32_t* blk_seg_off = nullptr;  // [n_blk + 1]
    double* blk_bTg = nullptr;    // bundle: [12][n_blk]
    int32_t* blk_vfree = nullptr; // [n_blk] 1 if the view-type pose of this block is free
};

struct EvalBuffers {
    double* x = nullptr;          // ambient parameters [n_amb]
    CamConst* camc = nullptr;     // [n_cams]
    double* camT = nullptr;       // [n_cams][36]
    double* seg_frame = nullptr;  // [9][n_seg]
    double* blk_Tv = nullptr;     // [36][n_blk]
    double* segN = nullptr;       // [NE][n_seg]
    double* seg_ssr = nullptr;    // [n_seg] (cost passes)
    double* blk_ssr = nullptr;    // [n_blk]
    double* blk_w = nullptr;      // [n_blk] Huber weight rho'(s_b)
    double* seg_w = nullptr;      // [n_seg] weight of the segment's block
    double* blk_rows = nullptr;   // block-indexed rows [1 + 63 + 6 PI][n_blk]: cost | Hvv gv Q Evi (bundle)
    // per-block view-type outputs for the per-view (Schur) kinds: [entry][n_blk]
    double* blk_Hvv = nullptr;    // 21
    double* blk_gv = nullptr;     // 6
    double* blk_Evc = nullptr;    // 36  (T_v^T N_xixi T_c)
    double* blk_Evi = nullptr;    // 6 * PI
    double* partial = nullptr;    // [n_seg_chunks][NE]
    double* partial_blk = nullptr;// [n_blk_chunks][NV - NE]
    double* cam_sums = nullptr;   // [n_cams][NV]
    // fused K1: one row of NVT values per tile, summed per camera by k_tile_reduce
    double* tile_vals = nullptr;  // [n_tiles][nvt]
    double* partial_tile = nullptr;  // [n_tile_chunks][nvt]
    int32_t* tile_vmap = nullptr; // [nvt] -> index into a camera's cam_sums row
};

struct ProblemShape {
    int kind, model, imode;  // imode: INTR_NONE / NOSKEW / SKEW
    int n_cams, n_views;
    int P, PI, NC, NL, NE;
    int NV;                  // reduced values per camera
    int view_free_global;    // bundle: b pose free
    int cam_pose_kind;       // 0 none, 1 extrinsics, 2 bundle
    // offsets of parameter blocks inside x
    int off_intr, off_camq, off_camt, off_viewq, off_viewt;
    double huber_delta;
};

// Composite pose c_se3_t of device block b (the pose chains of src/estimation/residuals/{intrinsic,extrinsics,bundle}residual.h)
// from the parameters x (k_block_setup).  Deriving it inside K1 per block instead of through the set-up kernel was measured
// on B200 and is slower (the dependent chain is not hidden with two warps per sub-partition): +6.5 % on K1 against the 46 us launch.
// real == false (padding block of a camera group): the identity, so nothing downstream sees garbage.
CAL_HD void block_pose(const ProblemShape& S, const DevLayout& L, const double* __restrict__ x, int64_t b, int cam, bool real, BlockPose& bp) {
    if (!real) {
        for (int i = 0; i < 9; ++i) bp.R[i] = bp.M[i] = (i % 4 == 0) ? 1.0 : 0.0;
        bp.t[0] = bp.t[1] = bp.t[2] = 0.0;
    } else if (S.kind == 0) {
        const int v = L.blk_view[b];
        compose_intrinsics(x + S.off_viewq + 4 * v, x + S.off_viewt + 3 * v, bp);
    } else if (S.kind == 1) {
        const int v = L.blk_view[b];
        compose_extrinsics(x + S.off_camq + 4 * cam, x + S.off_camt + 3 * cam, x + S.off_viewq + 4 * v, x + S.off_viewt + 3 * v, bp);
    } else {
        double bTg[12];
        for (int i = 0; i < 12; ++i) bTg[i] = L.blk_bTg[(int64_t)i * L.n_blk + b];
        compose_bundle(x + S.off_viewq, x + S.off_viewt, x + S.off_camq + 4 * cam, x + S.off_camt + 3 * cam, bTg, bp);
    }
}
CAL_HD const double* cam_intrinsics(const ProblemShape& S, const double* __restrict__ x, int cam) { return x + S.off_intr + (S.kind == 0 ? 0 : cam * S.P); }

// contiguous column ranges (within one camera group) reduced by one CTA of k_colsum
struct ColChunk { int32_t cam; int32_t pad; int64_t begin, end; };
struct ReduceDesc {
    ColChunk* seg_chunks = nullptr; int n_seg_chunks = 0; int32_t* seg_cam_chunk_off = nullptr;  // [n_cams + 1]
    ColChunk* blk_chunks = nullptr; int n_blk_chunks = 0; int32_t* blk_cam_chunk_off = nullptr;
    ColChunk* tile_chunks = nullptr; int n_tile_chunks = 0; int32_t* tile_cam_chunk_off = nullptr; int nvt = 0;  // fused K1
    unsigned* tile_tickets = nullptr;   // [n_cams + 1] arrival counters of k_tile_reduce (zero between launches)
    int n_active_cams = 0;              // cameras with at least one tile on this GPU
};

void launch_repack(const DevLayout& L, const double* sx, const double* sy, const double* su, const double* sv,
                   const int64_t* seg_src, int board_n, cudaStream_t st);
void launch_btg_permute(const DevLayout& L, const double* src, cudaStream_t st);
void launch_setup(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, cudaStream_t st);
void launch_k1(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, cudaStream_t st);
void launch_cost(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, cudaStream_t st);
// jac != 0: reduce the K1 output; else reduce seg_ssr to per-camera cost.  Returns the number of launches.
// peer (world > 1): the fused layout's reduction all-reduces cam_sums over NVLink itself; *peer_done says whether it did.
int launch_assemble(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, const ReduceDesc& R, int jac,
                    const calcomm::PeerArgs* peer, bool* peer_done, cudaStream_t st);
int k1_num_passes(const ProblemShape& S);  // roles (warps per tile) of K1
// fused K1: number of roles, values per tile and their index in a camera's cam_sums row
void k1_tile_value_map(const ProblemShape& S, int* n_roles, int* nvt, std::vector<int32_t>* map);
int launch_tile_reduce(const ProblemShape& S, const EvalBuffers& B, const ReduceDesc& R, int nvt, const calcomm::PeerArgs& peer, cudaStream_t st);
float dfma_peak_ms(double* scratch, int blocks, int threads, int iters, cudaStream_t st);

// ---- per-view (Schur) machinery -------------------------------------------------
constexpr int kSyrkTile = 8;
constexpr int kSyrkThreads = 480;  // at most fifteen warps (syrk_shape)
constexpr int kReduceViewsCtas = 64;  // CTAs of k_reduce_views at most; ViewBuffers::red_part holds [kReduceViewsCtas][4]
constexpr int kReducedMaxN = 160;   // widest shared block k_reduced_solve takes (reduced_solve_smem_doubles(160) = 230 KB of the 227 KiB a CTA may have)
constexpr int kSyrkMaxN = 176;  // ns + 1 must not exceed this (22 x 8 tiles, 253 <= 256 threads)

struct ViewBuffers {
    int32_t* view_blk_off = nullptr;  // [n_views + 1] CSR of device block ids per view
    int32_t* view_blk_idx = nullptr;
    int32_t* view_free = nullptr;     // [n_views]
    int32_t* cam_col_q = nullptr;     // [n_cams] shared column of the camera pose quat tangent (-1 const)
    int32_t* cam_col_t = nullptr;
    int32_t* cam_col_i = nullptr;
    double* Hpp = nullptr;      // [n_views][36]
    double* gp = nullptr;       // [n_views][6]
    double* sp = nullptr;       // jacobi scale [n_views][6]
    double* dp = nullptr;       // clamped LM diagonal [n_views][6]
    double* Lp = nullptr;       // cholesky factors [n_views][36]
    double* Linv = nullptr;     // reciprocals of their diagonals [n_views][6]
    double* view_f = nullptr;   // L^-1 (sp o gp) [n_views][6]
    double* Fd = nullptr;       // dense rows [F_v | f_v] = L_v^-1 [E_v | g_v], [n_views][6][ncp]; zeros where a camera does not see the view
    int ncp = 0, ns = 0;        // row pitch of Fd (ns + 1 rounded up to the SYRK tile width) and the shared tangent width
    double* delta_p = nullptr;  // tangent step [n_views][6]
    double* s_shared = nullptr; // [ns] jacobi scale of the shared columns
    double* y_shared = nullptr; // [ns] reduced solution
    double* C = nullptr;        // [ns*ns] sum E^T A^-1 E
    double* c = nullptr;        // [ns]    sum E^T A^-1 g
    double* partialC = nullptr; // [n_cta][(ns+1)^2]
    double* red = nullptr;      // [n_views][4] per-view scalars to reduce
    double* red_out = nullptr;  // [4]
    double* red_part = nullptr; // [kReduceViewsCtas][4] per-CTA partials of k_reduce_views
    unsigned* red_ticket = nullptr;  // its arrival counter (zero between launches)
    double* x_cand = nullptr;   // candidate parameters [n_amb]
    int32_t* fail = nullptr;    // cholesky failure flag
};
constexpr int kSchurMaxCtas = 444;   // 3 x 148
int schur_num_ctas(int n_views);
void launch_view_gather(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, const ViewBuffers& V,
                        cudaStream_t st);
void launch_view_scale(const ProblemShape& S, const ViewBuffers& V, int compute_scale, cudaStream_t st);
void launch_schur(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B, const ViewBuffers& V, int ns,
                  double radius, cudaStream_t st);
bool launch_reduced_solve(const double* Sm, const double* gss, const ViewBuffers& V, int ns, int32_t* info, cudaStream_t st);
void launch_backsub(const ProblemShape& S, const DevLayout& L, const ViewBuffers& V, int ns, cudaStream_t st);
void launch_view_plus(const ProblemShape& S, const EvalBuffers& B, const ViewBuffers& V, double t, cudaStream_t st);
void launch_view_norms(const ProblemShape& S, const EvalBuffers& B, const ViewBuffers& V, cudaStream_t st);
void launch_reduce_views(const ViewBuffers& V, int n_views, cudaStream_t st);
// block-structured covariance of the per-view kinds (after launch_schur with an infinite radius): W = inverse of the
// reduced (scaled) shared block [ns][ns]; Z, G: [n_views][6][ns]; Ainv: [n_views][36]; cov: [na][na] zero-initialised
void launch_cov_views(const ProblemShape& S, const DevLayout& L, const ViewBuffers& V, const double* x, int ns, const double* W,
                      double* Z, double* G, double* Ainv, double* cov, int64_t na, cudaStream_t st);

}  // namespace calk
