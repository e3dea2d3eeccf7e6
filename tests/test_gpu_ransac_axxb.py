"""Parity of the batched RANSAC homography kernel (K3) and the AX = XB kernel with the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as O
import ref_scenarios as RS
from calibration_b200 import abi, capi, synth
from calibration_b200 import geometry as G

pytestmark = pytest.mark.gpu


def compare_ransac(x, y, u, v, opts, seed_per_problem=True, min_margin=1e-9):
    ro, mo = O.ransac_batch(x, y, u, v, opts, seed_per_problem)
    rg, mg = capi.ransac_homography_batch(x, y, u, v, opts, seed_per_problem)
    n_checked = 0
    for p in range(x.shape[0]):
        if ro[p].min_margin <= min_margin:   # a residual within rounding distance of the threshold: not comparable bit for bit
            continue
        n_checked += 1
        assert rg[p].success == ro[p].success, p
        assert np.array_equal(mg[p], mo[p]), p                   # bit-exact inlier mask
        assert rg[p].n_inliers == ro[p].n_inliers and rg[p].iters == ro[p].iters and rg[p].iters_run == ro[p].iters_run, p
        if ro[p].success:
            Ho, Hg = np.array(ro[p].hmtx), np.array(rg[p].hmtx)
            assert np.abs(Hg - Ho).max() <= 1e-7 * np.abs(Ho).max(), p
            # absolute floors: an exact fit leaves residuals at rounding level (1e-14), and symmetric_rms_px takes
            # the square root of a SUM OF ROOTS, which turns 1e-12 into 1e-6
            assert abs(rg[p].inlier_rms - ro[p].inlier_rms) <= 1e-8 * ro[p].inlier_rms + 1e-9
            assert abs(rg[p].symmetric_rms_px - ro[p].symmetric_rms_px) <= 1e-8 * ro[p].symmetric_rms_px + 1e-5
    return n_checked, ro, rg


def test_batched_ransac_bit_exact_masks():
    # BASELINE configs[1] shape: 500 correspondences, 30 % outliers, seed 1234567 + problem id
    x, y, u, v, _ = synth.synth_ransac(seed=17, n_problems=512, n=500)
    n_checked, ro, rg = compare_ransac(x, y, u, v, abi.RansacOptions.default())
    assert n_checked >= 0.99 * 512
    assert sum(r.success for r in rg) == 512
    assert 300 < np.mean([r.n_inliers for r in rg]) < 380


def test_batched_ransac_ten_thousand_problems_through_the_chunked_pipeline(monkeypatch):
    """A tenth of BASELINE configs[1] (10 000 x 500): the host entry cuts the batch into chunks that overlap upload, kernel
    and download; every chunk's seeds continue the global problem index, so the masks stay bit-exact against the oracle."""
    x, y, u, v, _ = synth.synth_ransac(seed=23, n_problems=10000, n=500)
    monkeypatch.setenv("CALIB_B200_RANSAC_CHUNK", "3000")   # three chunks here (the default chunk is 25 000 problems)
    n_checked, ro, rg = compare_ransac(x, y, u, v, abi.RansacOptions.default())
    assert n_checked >= 0.99 * 10000
    assert sum(r.success for r in rg) == sum(r.success for r in ro) >= 9990


def test_batch_split_over_devices_returns_what_one_device_returns():
    """cal_ransac_homography_batch_multi (independent problems, no communication): slices on several devices — here the
    same device three times, and every device of the box when there are more — give the single-device results bit for bit."""
    x, y, u, v, _ = synth.synth_ransac(seed=29, n_problems=1500, n=300)
    opts = abi.RansacOptions.default()
    r1, m1 = capi.ransac_homography_batch(x, y, u, v, opts)
    for devices in ([0, 0, 0], list(range(capi.device_count()))):
        rm, mm = capi.ransac_homography_batch(x, y, u, v, opts, devices=devices)
        assert np.array_equal(mm, m1)
        assert bytes(rm) == bytes(r1)
    with pytest.raises(capi.CalibCudaError):
        capi.ransac_homography_batch(x, y, u, v, opts, devices=[0, 99])


@pytest.mark.parametrize("n", [4, 5, 31, 33, 130, 257])
def test_ragged_sizes(n):
    x, y, u, v, _ = synth.synth_ransac(seed=n, n_problems=40, n=n, outlier_fraction=0.2)
    opts = abi.RansacOptions.default(min_inliers=min(12, n), max_iters=200)
    n_checked, _, _ = compare_ransac(x, y, u, v, opts)
    assert n_checked >= 36


def test_options_variants():
    x, y, u, v, _ = synth.synth_ransac(seed=5, n_problems=64, n=300, outlier_fraction=0.5)
    for opts in (abi.RansacOptions.default(refit_on_inliers=0), abi.RansacOptions.default(min_inliers=250, max_iters=150),
                 abi.RansacOptions.default(thresh=0.8, confidence=0.999), abi.RansacOptions.default(confidence=0.0, max_iters=40)):
        n_checked, _, _ = compare_ransac(x, y, u, v, opts)
        assert n_checked >= 60
    n_checked, _, _ = compare_ransac(x, y, u, v, abi.RansacOptions.default(seed=99), seed_per_problem=False)
    assert n_checked >= 60


def test_reference_homography_tests_on_gpu():
    # homography_test.cpp:104-134
    H, xyuv = O.homography_testdata(100, 0.0, 30, 7)
    opts = abi.RansacOptions.default(thresh=1.0, min_inliers=90, seed=123)
    cols = [np.ascontiguousarray(xyuv[None, :, k]) for k in range(4)]
    res, mask = capi.ransac_homography_batch(*cols, opts, seed_per_problem=False)
    ro, mo = O.ransac(xyuv[:, 0], xyuv[:, 1], xyuv[:, 2], xyuv[:, 3], opts)
    assert res[0].success and res[0].n_inliers >= 95 and res[0].symmetric_rms_px < 1e-3
    assert np.array_equal(mask[0], mo) and res[0].iters == ro.iters
    Hr = np.array(res[0].hmtx).reshape(3, 3)
    assert np.allclose(Hr / Hr[2, 2], H, rtol=1e-2, atol=1e-2)
    # homography_test.cpp:137-160
    H, xyuv = O.homography_testdata(4, 0.0, 50, 3)
    opts = abi.RansacOptions.default(thresh=0.5, min_inliers=10, seed=42)
    cols = [np.ascontiguousarray(xyuv[None, :, k]) for k in range(4)]
    res, mask = capi.ransac_homography_batch(*cols, opts, seed_per_problem=False)
    assert not res[0].success and mask.sum() == 0


def test_fewer_than_four_points():
    z = np.array([[0.0, 1.0, 0.0]])
    res, mask = capi.ransac_homography_batch(z, z[:, ::-1].copy(), z + 10, z, abi.RansacOptions.default())
    assert not res[0].success and res[0].iters == 0 and mask.sum() == 0


# ---------------------------------------------------------------------------
# AX = XB
# ---------------------------------------------------------------------------
def handeye_reference_scenario():
    """CeresAXXBRefine.ImprovesOverInitializer (handeye_test.cpp:101-152)."""
    b_se3_g, pre, post = O.handeye_sequence(2024, 18, n_pre=2, n_post=1)
    X_gt = G.make_pose([0.02, -0.01, 0.09], pre[0], RS.deg2rad(10.0))
    b_t = G.make_pose([0.25, 0.05, 0.55], pre[1], RS.deg2rad(18.0))
    c_se3_t = [G.inv_pose(X_gt) @ G.inv_pose(T) @ b_t for T in b_se3_g]
    X0 = X_gt.copy()
    X0[:3, :3] = G.angle_axis_to_R(post[0] / np.linalg.norm(post[0]), RS.deg2rad(2.0)) @ X0[:3, :3]
    X0[:3, 3] += [0.01, -0.005, 0.004]
    return b_se3_g, c_se3_t, X_gt, X0


def test_axxb_reference_scenario():
    bg, ct, X_gt, X0 = handeye_reference_scenario()
    ra, rb, ta, tb = O.build_all_pairs(bg, ct, 0.5)   # optimize_handeye uses min_angle 0.5 deg (handeye.cpp:64)
    assert len(ta) > 100
    x0 = G.pack_handeye(X0)
    opts = abi.OptimOptions.default(max_iterations=60)
    d = O.axxb_desc(ra, rb, ta, tb, 1.0)
    h = capi.AxxbHandle(ra, rb, ta, tb, 1.0)
    c_o, g_o, H_o = O.axxb_eval(d, x0)
    c_g, g_g, H_g = h.eval(x0)
    assert abs(c_g - c_o) <= 1e-12 * c_o
    assert np.abs(g_g - g_o).max() <= 1e-10 * np.abs(g_o).max() and np.abs(H_g - H_o).max() <= 1e-10 * np.abs(H_o).max()
    x_o, r_o, cov_o = O.axxb_solve(d, x0, opts)
    x_g, r_g, cov_g = h.solve(x0, opts)
    h.close()
    assert r_g.success == r_o.success and np.abs(x_g - x_o).max() <= 1e-8
    assert r_g.covariance_ok and np.abs(cov_g - cov_o).max() <= 1e-6 * np.abs(cov_o).max()
    Xr = G.qt_to_pose(x_g[:4], x_g[4:])
    err0 = np.degrees(G.rotation_angle(X0[:3, :3].T @ X_gt[:3, :3]))
    err1 = np.degrees(G.rotation_angle(Xr[:3, :3].T @ X_gt[:3, :3]))
    assert err1 < err0 and err1 < 0.05 and np.linalg.norm(Xr[:3, 3] - X_gt[:3, 3]) < 0.002


def test_axxb_noisy_pairs_at_scale():
    """~0.5 M motion pairs with perturbed camera poses: Huber active, analytic Jacobian vs the oracle's duals."""
    bg, ct, X_gt = synth.make_handeye_poses(seed=3, n=1000)
    rng = np.random.default_rng(0)
    ct = [synth.perturb_pose(rng, T, 0.3, 0.002) for T in ct]
    ra, rb, ta, tb = O.build_all_pairs(bg, ct, 0.5)
    assert len(ta) > 400000
    x0 = G.pack_handeye(synth.perturb_pose(rng, X_gt, 2.0, 0.01))
    d = O.axxb_desc(ra, rb, ta, tb, 0.05)
    h = capi.AxxbHandle(ra, rb, ta, tb, 0.05)
    c_o, g_o, H_o = O.axxb_eval(d, x0)
    c_g, g_g, H_g = h.eval(x0)
    assert abs(c_g - c_o) <= 1e-11 * c_o
    assert np.abs(g_g - g_o).max() <= 1e-9 * np.abs(g_o).max() and np.abs(H_g - H_o).max() <= 1e-9 * np.abs(H_o).max()
    x_g, r_g, _ = h.solve(x0)
    x_o, r_o, _ = O.axxb_solve(d, x0)
    h.close()
    assert r_g.success and np.abs(x_g - x_o).max() <= 1e-8


def test_axxb_pairs_on_the_fly_match_materialised_pairs():
    """cal_axxb_create_from_poses: build_all_pairs (handeyedlt.cpp:51-81) on the device, pairs formed in every pass.
    Same kept-pair count as the oracle's build_all_pairs, same normal equations and the same solution."""
    bg, ct, X_gt, X0 = handeye_reference_scenario()
    # a stationary robot step and a pure translation: is_good_pair must drop their pairs like the reference
    bg = bg + [bg[-1].copy(), bg[3] @ G.make_pose([0.05, 0.0, 0.01])]
    ct = ct + [ct[-1].copy(), G.inv_pose(X_gt) @ G.inv_pose(bg[-1]) @ (bg[3] @ X_gt @ ct[3])]
    ra, rb, ta, tb = O.build_all_pairs(bg, ct, 0.5)
    n_all = len(bg) * (len(bg) - 1) // 2
    assert 100 < len(ta) < n_all
    x0 = G.pack_handeye(X0)
    h = capi.AxxbHandle.from_poses(bg, ct, 1.0)
    assert h.n_pairs == len(ta)
    d = O.axxb_desc(ra, rb, ta, tb, 1.0)
    c_o, g_o, H_o = O.axxb_eval(d, x0)
    c_g, g_g, H_g = h.eval(x0)
    assert abs(c_g - c_o) <= 1e-12 * c_o
    assert np.abs(g_g - g_o).max() <= 1e-10 * np.abs(g_o).max() and np.abs(H_g - H_o).max() <= 1e-10 * np.abs(H_o).max()
    opts = abi.OptimOptions.default(max_iterations=60)
    x_o, r_o, cov_o = O.axxb_solve(d, x0, opts)
    x_g, r_g, cov_g = h.solve(x0, opts)
    h.close()
    assert r_g.success == r_o.success and np.abs(x_g - x_o).max() <= 1e-8
    assert np.abs(cov_g - cov_o).max() <= 1e-6 * np.abs(cov_o).max()


def test_axxb_pairs_on_the_fly_at_scale():
    """1000 noisy poses (499 500 candidate pairs, several tiles incl. ragged edge tiles) against the materialised path."""
    bg, ct, X_gt = synth.make_handeye_poses(seed=3, n=1000)
    rng = np.random.default_rng(0)
    ct = [synth.perturb_pose(rng, T, 0.3, 0.002) for T in ct]
    ra, rb, ta, tb = O.build_all_pairs(bg, ct, 0.5)
    x0 = G.pack_handeye(synth.perturb_pose(rng, X_gt, 2.0, 0.01))
    hm = capi.AxxbHandle(ra, rb, ta, tb, 0.05)
    hf = capi.AxxbHandle.from_poses(bg, ct, 0.05)
    assert hf.n_pairs == len(ta)
    c_m, g_m, H_m = hm.eval(x0)
    c_f, g_f, H_f = hf.eval(x0)
    assert abs(c_f - c_m) <= 1e-11 * c_m and np.abs(g_f - g_m).max() <= 1e-9 * np.abs(g_m).max() and np.abs(H_f - H_m).max() <= 1e-9 * np.abs(H_m).max()
    c2, g2, H2 = hf.eval(x0)
    assert c2 == c_f and np.array_equal(g2, g_f) and np.array_equal(H2, H_f)   # deterministic
    x_m, r_m, _ = hm.solve(x0); x_f, r_f, _ = hf.solve(x0)
    hm.close(); hf.close()
    assert r_f.success and np.abs(x_f - x_m).max() <= 1e-9


def test_axxb_at_the_named_size_five_thousand_poses():
    """BASELINE configs[3]: optimize_handeye from 5 000 robot poses.  The device forms all 12.5 M motion pairs on the fly; the
    oracle materialises them (build_all_pairs) — same kept-pair count, same normal equations, same solution and covariance."""
    bg, ct, X_gt = synth.make_handeye_poses(seed=3, n=5000)
    rng = np.random.default_rng(0)
    ct = [synth.perturb_pose(rng, T, 0.3, 0.002) for T in ct]
    x0 = G.pack_handeye(synth.perturb_pose(rng, X_gt, 2.0, 0.01))
    ra, rb, ta, tb = O.build_all_pairs(bg, ct, 0.5)
    h = capi.AxxbHandle.from_poses(bg, ct, 0.05)
    try:
        assert h.n_pairs == len(ta) > 12_000_000
        d = O.axxb_desc(ra, rb, ta, tb, 0.05)
        c_o, g_o, H_o = O.axxb_eval(d, x0)
        c_g, g_g, H_g = h.eval(x0)
        assert abs(c_g - c_o) <= 1e-11 * c_o
        assert np.abs(g_g - g_o).max() <= 1e-9 * np.abs(g_o).max() and np.abs(H_g - H_o).max() <= 1e-9 * np.abs(H_o).max()
        x_o, r_o, cov_o = O.axxb_solve(d, x0)
        x_g, r_g, cov_g = h.solve(x0)
    finally:
        h.close()
    assert r_g.success == r_o.success and r_g.iterations == r_o.iterations and np.abs(x_g - x_o).max() <= 1e-8
    assert np.abs(cov_g - cov_o).max() <= 1e-6 * np.abs(cov_o).max()


def test_axxb_from_poses_errors_mirror_reference():
    I = [np.eye(4)] * 5
    with pytest.raises(RuntimeError, match="No valid motion pairs"):     # handeyedlt.cpp:76-79
        capi.AxxbHandle.from_poses(I, I)
    with pytest.raises(RuntimeError, match="Inconsistent hand-eye input sizes"):   # handeyedlt.cpp:56-58
        capi.AxxbHandle.from_poses(I[:1], I[:1])


def test_axxb_no_pairs_is_runtime_error():
    with pytest.raises(RuntimeError, match="No valid motion pairs"):
        capi.AxxbHandle(np.zeros((0, 9)), np.zeros((0, 9)), np.zeros((0, 3)), np.zeros((0, 3)))
