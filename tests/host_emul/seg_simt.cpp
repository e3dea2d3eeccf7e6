// TEST INFRASTRUCTURE — the PRODUCT's segment layout (the default for problems below 16 384 residual blocks: BASELINE
// configs[0] and configs[2]) on the CPU: k_repack, k_block_setup, K1 with the NOT_FUSED epilogue (local system per
// segment), k_block_weight, k_view_part, k_colsum / k_final_reduce (calibration_b200/csrc/k1_kernel.cuh,
// refine_setup_kernels.cuh, refine_assemble_kernels.cuh) and, for the per-view kinds, k_view_gather
// (refine_schur_kernels.cuh), compiled by g++ and run under the lock-step SIMT shim in the launch order of
// device_pass / launch_assemble, followed by the product's host assembly (refine_model.hpp).  The integer layout
// tables are built here the way cal_refine_create builds them for the segment layout.
#define CALIB_SIMT_SHIM 1
#include "simt_shim.hpp"

namespace calk {
__attribute__((aligned(128))) unsigned char k1_smem[232448];
}

#include "../../calibration_b200/csrc/k1_kernel.cuh"
#include "../../calibration_b200/csrc/refine_model.hpp"
#include "../../calibration_b200/csrc/refine_setup_kernels.cuh"
#include "../../calibration_b200/csrc/refine_schur_kernels.cuh"
#include "../../calibration_b200/csrc/refine_cost_kernel.cuh"   // k_cost with the per-warp TMA ring of tile_stage.cuh
// k_colsum's 8-double scratch is the only static shared array used here: one CTA runs at a time, so a plain static serves
#undef __shared__
#define __shared__ static
#include "../../calibration_b200/csrc/refine_assemble_kernels.cuh"

using namespace calk;

namespace {

void make_chunks(int n_cams, const std::vector<int32_t>& cam_of, int64_t n, std::vector<ColChunk>& ch, std::vector<int32_t>& off, int64_t kChunk) {
    off.assign(n_cams + 1, 0);
    std::vector<std::vector<ColChunk>> per(n_cams);
    for (int64_t i = 0; i < n;) {
        const int cam = cam_of[i]; int64_t j = i;
        while (j < n && cam_of[j] == cam) ++j;
        for (int64_t k = i; k < j; k += kChunk) per[cam].push_back(ColChunk{cam, 0, k, std::min(j, k + kChunk)});
        i = j;
    }
    for (int c = 0; c < n_cams; ++c) { off[c] = (int32_t)ch.size(); ch.insert(ch.end(), per[c].begin(), per[c].end()); }
    off[n_cams] = (int32_t)ch.size();
}

template <int MODEL, int IMODE>
void k1_not_fused(const K1Args& P) {
    using RT = K1Roles<MODEL, IMODE>;
    if (K1Smem<MODEL, IMODE>::kBytes > (int)sizeof k1_smem) std::abort();
    simt::launch((unsigned)P.L.n_tiles, RT::NROLE * 32, [&] { k1_kernel<MODEL, IMODE, NOT_FUSED>(P); });
}
template <int MODEL, int IMODE>
void view_part(const ProblemShape& S, const DevLayout& L, const EvalBuffers& B) {
    const unsigned g = (unsigned)((L.n_blk + 127) / 128);
    if (L.one_seg_per_blk) simt::launch(g, 128, [&] { k_view_part<MODEL, IMODE, true>(S, L, B); });
    else simt::launch(g, 128, [&] { k_view_part<MODEL, IMODE, false>(S, L, B); });
}
#define DISPATCH(FN, ...)                                                  \
    do {                                                                   \
        if (S.model == 0 && S.imode == 0) FN<0, 0>(__VA_ARGS__);           \
        else if (S.model == 0 && S.imode == 1) FN<0, 1>(__VA_ARGS__);      \
        else if (S.model == 0 && S.imode == 2) FN<0, 2>(__VA_ARGS__);      \
        else if (S.model == 1 && S.imode == 0) FN<1, 0>(__VA_ARGS__);      \
        else if (S.model == 1 && S.imode == 1) FN<1, 1>(__VA_ARGS__);      \
        else FN<1, 2>(__VA_ARGS__);                                        \
    } while (0)

}  // namespace

// target_len: corners per segment (cal_refine_create uses min(max(8, n_obs / 65536), longest block)); chunk: columns per
// k_colsum CTA (8192 in the product; small values exercise several chunks per camera)
// cost_only != 0: the residual-only pass (k_cost); g then receives the per-block sums of squared residuals (n_blocks values)
extern "C" int simt_segment_eval(const cal_problem_desc* dp, const double* x, int target_len, int chunk, int cost_only, double* cost, double* g,
                                 double* H, int32_t* n_segments_out) {
    const cal_problem_desc& d = *dp;
    HostModel M; M.init_model(d);
    const ProblemShape& S = M.S;
    const int nv = S.n_views;
    const bool bundle = d.kind == CAL_KIND_BUNDLE;
    auto view_of = [&](int64_t b) { return d.kind == CAL_KIND_INTRINSICS ? (int32_t)b : (bundle ? -1 : d.block_view[b]); };
    // device blocks sorted by camera, camera groups padded to 32 blocks
    std::vector<int64_t> blk_orig; std::vector<int32_t> blk_cam, blk_view;
    for (int c = 0; c < S.n_cams; ++c) {
        for (int64_t b = 0; b < d.n_blocks; ++b) if (d.block_cam[b] == c) { blk_orig.push_back(b); blk_cam.push_back(c); blk_view.push_back(view_of(b)); }
        while (blk_orig.size() % 32) { blk_orig.push_back(-1); blk_cam.push_back(c); blk_view.push_back(-1); }
    }
    const int64_t nb = (int64_t)blk_orig.size();
    std::vector<int32_t> vfree(std::max(nv, 0));
    for (int v = 0; v < nv; ++v) vfree[v] = !M.pbs[M.pb_viewq(v)].constant;
    // segments of at most target_len corners, equalised within a block (refine_host.cu)
    std::vector<int32_t> seg_len, seg_blk, seg_cam, blk_seg_off(nb + 1, 0), blk_vfree(nb, 0);
    std::vector<int64_t> seg_src;
    int one_seg = 1;
    for (int64_t b = 0; b < nb; ++b) {
        blk_seg_off[b] = (int32_t)seg_len.size();
        if (blk_orig[b] < 0) continue;
        const int64_t o0 = d.block_offset[blk_orig[b]], len = d.block_offset[blk_orig[b] + 1] - o0;
        const int64_t nsegb = (len + target_len - 1) / target_len, sl = (len + nsegb - 1) / nsegb;
        if (nsegb != 1) one_seg = 0;
        for (int64_t k = 0; k < len; k += sl) { seg_len.push_back((int32_t)std::min<int64_t>(sl, len - k)); seg_blk.push_back((int32_t)b); seg_cam.push_back(blk_cam[b]); seg_src.push_back(o0 + k); }
        blk_vfree[b] = bundle ? (d.optimize_target_pose != 0) : (blk_view[b] >= 0 && vfree[blk_view[b]]);
    }
    blk_seg_off[nb] = (int32_t)seg_len.size();
    const int64_t nseg_used = (int64_t)seg_len.size();
    while (seg_len.size() % 32) { seg_len.push_back(0); seg_blk.push_back(0); seg_cam.push_back(0); seg_src.push_back(0); }
    const int64_t nseg = (int64_t)seg_len.size(), nt = nseg / 32;
    std::vector<int64_t> tile_off(nt); std::vector<int32_t> tile_depth(nt);
    int64_t slices = 0;
    for (int64_t t = 0; t < nt; ++t) { int32_t dep = 0; for (int l = 0; l < 32; ++l) dep = std::max(dep, seg_len[t * 32 + l]); tile_off[t] = slices; tile_depth[t] = dep; slices += dep; }
    std::vector<double> obs((size_t)slices * 128), bTg((size_t)12 * nb);
    DevLayout L;
    L.n_seg = nseg; L.n_tiles = nt; L.n_blk = nb; L.n_slices = slices; L.one_seg_per_blk = one_seg; L.fused = 0;
    L.obs = obs.data(); L.tile_off = tile_off.data(); L.tile_depth = tile_depth.data(); L.seg_len = seg_len.data(); L.seg_blk = seg_blk.data();
    L.seg_cam = seg_cam.data(); L.blk_cam = blk_cam.data(); L.blk_view = blk_view.data(); L.blk_orig = blk_orig.data();
    L.blk_seg_off = blk_seg_off.data(); L.blk_bTg = bTg.data(); L.blk_vfree = blk_vfree.data();
    const double* ox = d.board_n > 0 ? d.board_x : d.obj_x;
    const double* oy = d.board_n > 0 ? d.board_y : d.obj_y;
    simt::launch((unsigned)nt, 128, [&] { k_repack(L, ox, oy, d.img_u, d.img_v, seg_src.data(), d.board_n); });
    if (bundle) simt::launch((unsigned)((nb * 12 + 255) / 256), 256, [&] { k_btg_permute(L, d.block_b_se3_g); });
    // reduction chunk tables
    std::vector<ColChunk> sc, bc; std::vector<int32_t> so, bo;
    std::vector<int32_t> seg_cam_trim(seg_cam.begin(), seg_cam.begin() + nseg_used);
    make_chunks(S.n_cams, seg_cam_trim, nseg_used, sc, so, chunk);
    make_chunks(S.n_cams, blk_cam, nb, bc, bo, chunk);
    // evaluation buffers (cleared once, as at create)
    const int n_brows = S.NV - S.NE, PIe = std::max(S.PI, 1);
    std::vector<CamConst> camc(S.n_cams);
    std::vector<double> xs(x, x + M.n_amb), camT((size_t)36 * S.n_cams), seg_frame((size_t)9 * nseg, 0.0), blk_Tv((size_t)36 * nb, 0.0),
        segN((size_t)S.NE * nseg, 0.0), seg_ssr(nseg, 0.0), blk_ssr(nb, 0.0), partial((size_t)std::max<size_t>(sc.size(), 1) * S.NE),
        partial_blk((size_t)std::max<size_t>(bc.size(), 1) * n_brows), cam_sums((size_t)S.n_cams * S.NV, 0.0), blk_w(nb, 0.0), seg_w(nseg, 0.0),
        blk_rows((size_t)n_brows * nb, 0.0), bHvv((size_t)21 * nb, 0.0), bgv((size_t)6 * nb, 0.0), bEvc((size_t)36 * nb, 0.0), bEvi((size_t)6 * PIe * nb, 0.0);
    EvalBuffers B;
    B.x = xs.data(); B.camc = camc.data(); B.camT = camT.data(); B.seg_frame = seg_frame.data(); B.blk_Tv = blk_Tv.data(); B.segN = segN.data();
    B.seg_ssr = seg_ssr.data(); B.blk_ssr = blk_ssr.data(); B.partial = partial.data(); B.partial_blk = partial_blk.data(); B.cam_sums = cam_sums.data();
    B.blk_w = blk_w.data(); B.seg_w = seg_w.data(); B.blk_rows = blk_rows.data(); B.blk_Hvv = bHvv.data(); B.blk_gv = bgv.data(); B.blk_Evc = bEvc.data();
    B.blk_Evi = bEvi.data();
    const int64_t n = std::max<int64_t>(nb, S.n_cams);
    simt::launch((unsigned)((n + 127) / 128), 128, [&] { k_block_setup(S, L, B); });
    const unsigned gb = (unsigned)((nb + 127) / 128);
    const int rr_row = S.NL * S.NC - S.NC * (S.NC - 1) / 2;
    const ColChunk* scp = sc.data(); const ColChunk* bcp = bc.data();
    if (cost_only) {
        // device_pass(residual only): k_cost, then launch_assemble(jac = 0): block weight from seg_ssr, one block-indexed row
        if (4 * kWarpStageBytes > (int)sizeof k1_smem) std::abort();
        if (S.model == 0) simt::launch((unsigned)((nt + 3) / 4), 128, [&] { k_cost<0>(L, B); });
        else simt::launch((unsigned)((nt + 3) / 4), 128, [&] { k_cost<1>(L, B); });
        simt::launch(gb, 128, [&] { k_block_weight<0>(S, L, B, rr_row); });
        std::vector<double> pb((size_t)std::max<size_t>(bc.size(), 1)), cs((size_t)S.n_cams, 0.0);
        simt::launch((unsigned)bc.size(), 256, [&] { k_colsum<false>(B.blk_rows, L.n_blk, nullptr, bcp, pb.data(), 1); }, 1u);
        simt::launch((unsigned)((S.n_cams + 127) / 128), 128, [&] { k_final_reduce(pb.data(), bo.data(), S.n_cams, 1, cs.data(), 1, 0); });
        double c = 0; for (int k = 0; k < S.n_cams; ++k) c += cs[k];
        *cost = c;
        for (int64_t b = 0; b < nb; ++b) if (blk_orig[b] >= 0) g[blk_orig[b]] = blk_ssr[b];   // per-block sum of squares, in the caller's block order
        if (n_segments_out) *n_segments_out = (int32_t)nseg_used;
        return 0;
    }
    // device_pass(jacobian): setup, K1, launch_assemble
    K1Args P{L, B, S.huber_delta, 0};
    DISPATCH(k1_not_fused, P);
    simt::launch(gb, 128, [&] { k_block_weight<1>(S, L, B, rr_row); });
    const bool has_view_part = bundle ? S.view_free_global != 0 : S.n_views > 0;
    if (has_view_part) DISPATCH(view_part, S, L, B);
    simt::launch((unsigned)sc.size(), 256, [&] { k_colsum<true>(B.segN, L.n_seg, B.seg_w, scp, B.partial, S.NE); }, (unsigned)S.NE);
    simt::launch((unsigned)((S.n_cams * S.NE + 127) / 128), 128, [&] { k_final_reduce(B.partial, so.data(), S.n_cams, S.NE, B.cam_sums, S.NV, 0); });
    simt::launch((unsigned)bc.size(), 256, [&] { k_colsum<false>(B.blk_rows, L.n_blk, nullptr, bcp, B.partial_blk, n_brows); }, (unsigned)n_brows);
    simt::launch((unsigned)((S.n_cams * n_brows + 127) / 128), 128, [&] { k_final_reduce(B.partial_blk, bo.data(), S.n_cams, n_brows, B.cam_sums, S.NV, S.NE); });
    double c = 0; for (int k = 0; k < S.n_cams; ++k) c += cam_sums[(size_t)k * S.NV + S.NE];
    *cost = c;
    std::vector<double> Hss, gs;
    M.assemble_shared(cam_sums.data(), x, Hss, gs);
    if (bundle) {
        std::memcpy(g, gs.data(), gs.size() * sizeof(double));
        std::memcpy(H, Hss.data(), Hss.size() * sizeof(double));
    } else {
        std::vector<int32_t> off(nv + 1, 0), idx;
        for (int64_t b = 0; b < nb; ++b) if (blk_orig[b] >= 0) off[blk_view[b] + 1]++;
        for (int v = 0; v < nv; ++v) off[v + 1] += off[v];
        idx.resize(off[nv]);
        { std::vector<int32_t> cur(off.begin(), off.end() - 1); for (int64_t b = 0; b < nb; ++b) if (blk_orig[b] >= 0) idx[cur[blk_view[b]]++] = (int32_t)b; }
        std::vector<double> Hpp((size_t)nv * 36, 0.0), gp((size_t)nv * 6, 0.0), Hd, gd;
        ViewBuffers V;
        V.view_blk_off = off.data(); V.view_blk_idx = idx.data(); V.view_free = vfree.data(); V.Hpp = Hpp.data(); V.gp = gp.data();
        simt::launch((unsigned)((nv + 127) / 128), 128, [&] { k_view_gather(S, L, B, V); });
        std::vector<char> vf(vfree.begin(), vfree.end());
        M.assemble_dense(Hss, gs, Hpp.data(), gp.data(), bEvc.data(), bEvi.data(), nb, vf.data(), off.data(), idx.data(), blk_cam.data(), Hd, gd);
        std::memcpy(g, gd.data(), gd.size() * sizeof(double));
        std::memcpy(H, Hd.data(), Hd.size() * sizeof(double));
    }
    if (n_segments_out) *n_segments_out = (int32_t)nseg_used;
    return 0;
}

extern "C" int64_t simt_seg_tangent_count(const cal_problem_desc* dp) { HostModel M; M.init_model(*dp); return M.n_tan; }
