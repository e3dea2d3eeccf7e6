// TEST INFRASTRUCTURE — the PRODUCT's fused pass for the bundle kind, device side, on the CPU: k_repack,
// k_btg_permute and k_block_setup (calibration_b200/csrc/refine_setup_kernels.cuh) and K1 itself (k1_kernel.cuh:
// role warps, the shared-memory exchange of Jacobian rows, the predicate-free and the predicated step loops, the
// fused epilogue with its padded transposes) compiled by g++ and run under the lock-step SIMT shim, followed by the
// product's host assembly (refine_model.hpp).  The integer layout tables (device blocks sorted by camera, camera
// groups padded to whole tiles of 32, tile depths and offsets) are built here the way cal_refine_create builds them
// for the fused layout; the per-camera sums of the tile rows go through k_tile_colsum / k_tile_final.
#define CALIB_SIMT_SHIM 1
#include "simt_shim.hpp"

namespace calk {
__attribute__((aligned(128))) unsigned char k1_smem[232448];
}

#include "../../calibration_b200/csrc/k1_kernel.cuh"
#include "../../calibration_b200/csrc/refine_model.hpp"
#include "../../calibration_b200/csrc/refine_setup_kernels.cuh"

using namespace calk;

namespace {

// k_tile_reduce (launch_tile_reduce, k1_fused.cu).  The product groups 32 tiles per chunk; one tile per chunk here, so
// that the last-arriving CTA's sum over a camera's chunks sees several chunks per camera at test sizes.
void tile_reduce(const ProblemShape& S, const DevLayout& L, const std::vector<double>& tile_vals, int nvt, const std::vector<int32_t>& vmap,
                 std::vector<double>& cam_sums) {
    std::vector<ColChunk> tc; std::vector<int32_t> to(S.n_cams + 1, 0);
    for (int c = 0; c < S.n_cams; ++c) {
        to[c] = (int32_t)tc.size();
        for (int64_t t = 0; t < L.n_tiles; ++t) if (L.seg_cam[t * 32] == c) tc.push_back(ColChunk{c, 0, t, t + 1});
    }
    to[S.n_cams] = (int32_t)tc.size();
    std::vector<double> partial(tc.size() * (size_t)nvt);
    std::vector<unsigned> tickets(S.n_cams + 1, 0u);
    int active = 0;
    for (int c = 0; c < S.n_cams; ++c) active += to[c + 1] > to[c] ? 1 : 0;
    const TileReduceArgs A{tile_vals.data(), nvt, tc.data(), partial.data(), to.data(), S.n_cams, vmap.data(), cam_sums.data(), S.NV, tickets.data(), active};
    simt::launch((unsigned)tc.size(), 256, [&] { k_tile_reduce(A, calcomm::PeerArgs{}); });
    for (unsigned t : tickets) if (t != 0u) std::abort();   // the last CTA leaves the tickets ready for the next launch
}

// setup + K1 + per-camera sums of the tile rows for one (model, intrinsics mode) instantiation
template <int MODEL, int IMODE>
void pass_t(const ProblemShape& S, const DevLayout& L, const double* x, int n_amb, std::vector<double>& cam_sums, int* n_roles) {
    using RT = K1Roles<MODEL, IMODE>;
    const int64_t nb = L.n_blk;
    const bool reduce_rows = S.view_free_global != 0;
    const int nvt = RT::nvt(reduce_rows);
    std::vector<int32_t> vmap; RT::value_map(reduce_rows, vmap);
    std::vector<CamConst> camc(S.n_cams);
    std::vector<double> xs(x, x + n_amb), camT((size_t)36 * S.n_cams), seg_frame((size_t)9 * nb), blk_Tv((size_t)36 * nb, 0.0),
        blk_ssr(nb), tile_vals((size_t)L.n_tiles * nvt, 0.0);
    EvalBuffers B;
    B.x = xs.data(); B.camc = camc.data(); B.camT = camT.data(); B.seg_frame = seg_frame.data(); B.blk_Tv = blk_Tv.data();
    B.blk_ssr = blk_ssr.data(); B.tile_vals = tile_vals.data();
    const int64_t n = std::max<int64_t>(nb, S.n_cams);
    simt::launch((unsigned)((n + 127) / 128), 128, [&] { k_block_setup(S, L, B); });
    K1Args P{L, B, S.huber_delta, nvt};
    if (K1Smem<MODEL, IMODE>::kBytes > (int)sizeof k1_smem) std::abort();
    if (reduce_rows) simt::launch((unsigned)L.n_tiles, RT::NROLE * 32, [&] { k1_kernel<MODEL, IMODE, VIEW_REDUCE>(P); });
    else simt::launch((unsigned)L.n_tiles, RT::NROLE * 32, [&] { k1_kernel<MODEL, IMODE, VIEW_NONE>(P); });
    tile_reduce(S, L, tile_vals, nvt, vmap, cam_sums);
    *n_roles = RT::NROLE;
}

}  // namespace

extern "C" int simt_bundle_eval(const cal_problem_desc* dp, const double* x, double* cost, double* g, double* H, int32_t* n_roles_out,
                                int32_t* uniform_tiles_out) {
    const cal_problem_desc& d = *dp;
    if (d.kind != CAL_KIND_BUNDLE) return 1;
    HostModel M; M.init_model(d);
    const ProblemShape& S = M.S;
    // ---- integer layout: device blocks sorted by camera, every camera group padded to whole tiles ----
    std::vector<int64_t> blk_orig; std::vector<int32_t> blk_cam;
    for (int c = 0; c < S.n_cams; ++c) {
        for (int64_t b = 0; b < d.n_blocks; ++b) if (d.block_cam[b] == c) { blk_orig.push_back(b); blk_cam.push_back(c); }
        while (blk_orig.size() % 32) { blk_orig.push_back(-1); blk_cam.push_back(c); }
    }
    const int64_t nb = (int64_t)blk_orig.size(), nt = nb / 32;
    std::vector<int32_t> seg_len(nb), seg_blk(nb), blk_view(nb, -1), blk_seg_off(nb + 1), blk_vfree(nb, 1), tile_depth(nt);
    std::vector<int64_t> seg_src(nb, 0), tile_off(nt);
    for (int64_t b = 0; b < nb; ++b) {
        const int64_t o = blk_orig[b];
        seg_len[b] = o >= 0 ? (int32_t)(d.block_offset[o + 1] - d.block_offset[o]) : 0;
        seg_src[b] = o >= 0 ? d.block_offset[o] : 0;
        seg_blk[b] = (int32_t)b; blk_seg_off[b] = (int32_t)b;
    }
    blk_seg_off[nb] = (int32_t)nb;
    int64_t slices = 0; int uniform = 0;
    for (int64_t t = 0; t < nt; ++t) {
        int dep = 0; bool uni = true;
        for (int l = 0; l < 32; ++l) dep = std::max(dep, seg_len[t * 32 + l]);
        for (int l = 0; l < 32; ++l) uni = uni && seg_len[t * 32 + l] == dep;
        uniform += uni ? 1 : 0;
        tile_depth[t] = dep; tile_off[t] = slices; slices += dep;
    }
    std::vector<double> obs((size_t)slices * 128), bTg((size_t)12 * nb);
    DevLayout L;
    L.n_seg = nb; L.n_tiles = nt; L.n_blk = nb; L.n_slices = slices; L.one_seg_per_blk = 1; L.fused = 1;
    L.obs = obs.data(); L.tile_off = tile_off.data(); L.tile_depth = tile_depth.data(); L.seg_len = seg_len.data(); L.seg_blk = seg_blk.data();
    L.seg_cam = blk_cam.data(); L.blk_cam = blk_cam.data(); L.blk_view = blk_view.data(); L.blk_orig = blk_orig.data();
    L.blk_seg_off = blk_seg_off.data(); L.blk_bTg = bTg.data(); L.blk_vfree = blk_vfree.data();
    // shared-board form of the descriptor (board_n > 0): the object points come from the one board
    const double* ox = d.board_n > 0 ? d.board_x : d.obj_x;
    const double* oy = d.board_n > 0 ? d.board_y : d.obj_y;
    simt::launch((unsigned)nt, 128, [&] { k_repack(L, ox, oy, d.img_u, d.img_v, seg_src.data(), d.board_n); });
    simt::launch((unsigned)((nb * 12 + 255) / 256), 256, [&] { k_btg_permute(L, d.block_b_se3_g); });
    // ---- setup, K1, per-camera sums, host assembly ----
    std::vector<double> cam_sums((size_t)S.n_cams * S.NV, 0.0);
    int n_roles = 0;
    if (S.model == 0 && S.imode == 0) pass_t<0, 0>(S, L, x, M.n_amb, cam_sums, &n_roles);
    else if (S.model == 0 && S.imode == 1) pass_t<0, 1>(S, L, x, M.n_amb, cam_sums, &n_roles);
    else if (S.model == 0 && S.imode == 2) pass_t<0, 2>(S, L, x, M.n_amb, cam_sums, &n_roles);
    else if (S.model == 1 && S.imode == 0) pass_t<1, 0>(S, L, x, M.n_amb, cam_sums, &n_roles);
    else if (S.model == 1 && S.imode == 1) pass_t<1, 1>(S, L, x, M.n_amb, cam_sums, &n_roles);
    else pass_t<1, 2>(S, L, x, M.n_amb, cam_sums, &n_roles);
    double c = 0; for (int k = 0; k < S.n_cams; ++k) c += cam_sums[(size_t)k * S.NV + S.NE];
    *cost = c;
    std::vector<double> Hss, gs;
    M.assemble_shared(cam_sums.data(), x, Hss, gs);
    std::memcpy(g, gs.data(), gs.size() * sizeof(double));
    std::memcpy(H, Hss.data(), Hss.size() * sizeof(double));
    if (n_roles_out) *n_roles_out = n_roles;
    if (uniform_tiles_out) *uniform_tiles_out = uniform;
    return 0;
}

extern "C" int64_t simt_tangent_count(const cal_problem_desc* dp) { HostModel M; M.init_model(*dp); return M.n_tan; }

// ---- kinds with per-view pose blocks (intrinsics, extrinsics): K1 with the VIEW_STORE epilogue (per-block H_vv, g_v,
// E_vc = Q T_c, E_vi written for the Schur kernels) + k_view_gather, then the product's dense host assembly ----
#include "../../calibration_b200/csrc/refine_schur_kernels.cuh"

namespace {
template <int MODEL, int IMODE>
void views_pass_t(const ProblemShape& S, const DevLayout& L, const double* x, int n_amb, std::vector<double>& cam_sums, EvalBuffers& B, int* n_roles) {
    using RT = K1Roles<MODEL, IMODE>;
    const int64_t nb = L.n_blk;
    const int nvt = RT::nvt(false);
    std::vector<int32_t> vmap; RT::value_map(false, vmap);
    std::vector<CamConst> camc(S.n_cams);
    std::vector<double> xs(x, x + n_amb), camT((size_t)36 * S.n_cams), seg_frame((size_t)9 * nb), blk_Tv((size_t)36 * nb, 0.0), blk_ssr(nb),
        tile_vals((size_t)L.n_tiles * nvt, 0.0);
    B.x = xs.data(); B.camc = camc.data(); B.camT = camT.data(); B.seg_frame = seg_frame.data(); B.blk_Tv = blk_Tv.data();
    B.blk_ssr = blk_ssr.data(); B.tile_vals = tile_vals.data();
    const int64_t n = std::max<int64_t>(nb, S.n_cams);
    simt::launch((unsigned)((n + 127) / 128), 128, [&] { k_block_setup(S, L, B); });
    K1Args P{L, B, S.huber_delta, nvt};
    if (K1Smem<MODEL, IMODE>::kBytes > (int)sizeof k1_smem) std::abort();
    simt::launch((unsigned)L.n_tiles, RT::NROLE * 32, [&] { k1_kernel<MODEL, IMODE, VIEW_STORE>(P); });
    tile_reduce(S, L, tile_vals, nvt, vmap, cam_sums);
    B.x = nullptr; B.camc = nullptr; B.camT = nullptr; B.seg_frame = nullptr; B.blk_Tv = nullptr; B.blk_ssr = nullptr; B.tile_vals = nullptr;
    *n_roles = RT::NROLE;
}
}  // namespace

extern "C" int simt_views_eval(const cal_problem_desc* dp, const double* x, double* cost, double* g, double* H, int32_t* n_roles_out) {
    const cal_problem_desc& d = *dp;
    if (d.kind == CAL_KIND_BUNDLE) return 1;
    HostModel M; M.init_model(d);
    const ProblemShape& S = M.S;
    const int nv = S.n_views;
    auto view_of = [&](int64_t b) { return S.kind == CAL_KIND_INTRINSICS ? (int32_t)b : d.block_view[b]; };
    // fused layout as cal_refine_create builds it: device blocks sorted by camera, camera groups padded to whole tiles
    std::vector<int64_t> blk_orig; std::vector<int32_t> blk_cam, blk_view;
    for (int c = 0; c < S.n_cams; ++c) {
        for (int64_t b = 0; b < d.n_blocks; ++b) if (d.block_cam[b] == c) { blk_orig.push_back(b); blk_cam.push_back(c); blk_view.push_back(view_of(b)); }
        while (blk_orig.size() % 32) { blk_orig.push_back(-1); blk_cam.push_back(c); blk_view.push_back(-1); }
    }
    const int64_t nb = (int64_t)blk_orig.size(), nt = nb / 32;
    std::vector<int32_t> seg_len(nb), seg_blk(nb), blk_seg_off(nb + 1), blk_vfree(nb, 0), tile_depth(nt), vfree(nv);
    std::vector<int64_t> seg_src(nb, 0), tile_off(nt);
    for (int v = 0; v < nv; ++v) vfree[v] = !M.pbs[M.pb_viewq(v)].constant;
    for (int64_t b = 0; b < nb; ++b) {
        const int64_t o = blk_orig[b];
        seg_len[b] = o >= 0 ? (int32_t)(d.block_offset[o + 1] - d.block_offset[o]) : 0;
        seg_src[b] = o >= 0 ? d.block_offset[o] : 0;
        seg_blk[b] = (int32_t)b; blk_seg_off[b] = (int32_t)b;
        blk_vfree[b] = o >= 0 && vfree[blk_view[b]];
    }
    blk_seg_off[nb] = (int32_t)nb;
    int64_t slices = 0;
    for (int64_t t = 0; t < nt; ++t) {
        int dep = 0;
        for (int l = 0; l < 32; ++l) dep = std::max(dep, seg_len[t * 32 + l]);
        tile_depth[t] = dep; tile_off[t] = slices; slices += dep;
    }
    std::vector<double> obs((size_t)slices * 128);
    DevLayout L;
    L.n_seg = nb; L.n_tiles = nt; L.n_blk = nb; L.n_slices = slices; L.one_seg_per_blk = 1; L.fused = 1;
    L.obs = obs.data(); L.tile_off = tile_off.data(); L.tile_depth = tile_depth.data(); L.seg_len = seg_len.data(); L.seg_blk = seg_blk.data();
    L.seg_cam = blk_cam.data(); L.blk_cam = blk_cam.data(); L.blk_view = blk_view.data(); L.blk_orig = blk_orig.data();
    L.blk_seg_off = blk_seg_off.data(); L.blk_vfree = blk_vfree.data();
    const double* ox = d.board_n > 0 ? d.board_x : d.obj_x;
    const double* oy = d.board_n > 0 ? d.board_y : d.obj_y;
    simt::launch((unsigned)nt, 128, [&] { k_repack(L, ox, oy, d.img_u, d.img_v, seg_src.data(), d.board_n); });
    // K1 with per-block outputs (zero-initialised, as cal_refine_create leaves them: held views and padding stay zero)
    const int PIe = std::max(S.PI, 1);
    std::vector<double> cam_sums((size_t)S.n_cams * S.NV, 0.0), bHvv((size_t)21 * nb, 0.0), bgv((size_t)6 * nb, 0.0), bEvc((size_t)36 * nb, 0.0),
        bEvi((size_t)6 * PIe * nb, 0.0);
    EvalBuffers B;
    B.blk_Hvv = bHvv.data(); B.blk_gv = bgv.data(); B.blk_Evc = bEvc.data(); B.blk_Evi = bEvi.data();
    int n_roles = 0;
    if (S.model == 0 && S.imode == 0) views_pass_t<0, 0>(S, L, x, M.n_amb, cam_sums, B, &n_roles);
    else if (S.model == 0 && S.imode == 1) views_pass_t<0, 1>(S, L, x, M.n_amb, cam_sums, B, &n_roles);
    else if (S.model == 0 && S.imode == 2) views_pass_t<0, 2>(S, L, x, M.n_amb, cam_sums, B, &n_roles);
    else if (S.model == 1 && S.imode == 0) views_pass_t<1, 0>(S, L, x, M.n_amb, cam_sums, B, &n_roles);
    else if (S.model == 1 && S.imode == 1) views_pass_t<1, 1>(S, L, x, M.n_amb, cam_sums, B, &n_roles);
    else views_pass_t<1, 2>(S, L, x, M.n_amb, cam_sums, B, &n_roles);
    double c = 0; for (int k = 0; k < S.n_cams; ++k) c += cam_sums[(size_t)k * S.NV + S.NE];
    *cost = c;
    // k_view_gather over the view -> device-block CSR
    std::vector<int32_t> off(nv + 1, 0), idx;
    for (int64_t b = 0; b < nb; ++b) if (blk_orig[b] >= 0) off[blk_view[b] + 1]++;
    for (int v = 0; v < nv; ++v) off[v + 1] += off[v];
    idx.resize(off[nv]);
    { std::vector<int32_t> cur(off.begin(), off.end() - 1); for (int64_t b = 0; b < nb; ++b) if (blk_orig[b] >= 0) idx[cur[blk_view[b]]++] = (int32_t)b; }
    std::vector<double> Hpp((size_t)nv * 36, 0.0), gp((size_t)nv * 6, 0.0);
    ViewBuffers V;
    V.view_blk_off = off.data(); V.view_blk_idx = idx.data(); V.view_free = vfree.data(); V.Hpp = Hpp.data(); V.gp = gp.data();
    simt::launch((unsigned)((nv + 127) / 128), 128, [&] { k_view_gather(S, L, B, V); });
    std::vector<double> Hss, gs, Hd, gd;
    M.assemble_shared(cam_sums.data(), x, Hss, gs);
    std::vector<char> vf(vfree.begin(), vfree.end());
    M.assemble_dense(Hss, gs, Hpp.data(), gp.data(), bEvc.data(), bEvi.data(), nb, vf.data(), off.data(), idx.data(), blk_cam.data(), Hd, gd);
    std::memcpy(g, gd.data(), gd.size() * sizeof(double));
    std::memcpy(H, Hd.data(), Hd.size() * sizeof(double));
    if (n_roles_out) *n_roles_out = n_roles;
    return 0;
}
