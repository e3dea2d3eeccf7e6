"""Parity of the batched seeding kernels (cal_seed_intrinsics, cal_seed_planar_poses) with the CPU oracle
restatement of estimate_intrinsics / estimate_planar_pose, on the reference's own test scenario and on
noisy multi-camera data with ragged and too-short views.  Tolerances: the GPU takes null vectors from
normal matrices instead of a Jacobi SVD of the design matrices, so results agree to ~1e-10 relative on
these well-conditioned problems; 1e-7 is asserted (these are seeds for the refinement, which the
reference's tests accept at 1e-6)."""
import numpy as np
import pytest

import oracle_lib as O
from calibration_b200 import abi, capi, synth
from test_oracle_seed import intrinsics_estimate_scenario

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))


def test_reference_scenario_matches_oracle_and_ground_truth():
    intr, c_se3_t, (xs, ys, us, vs, off) = intrinsics_estimate_scenario()
    g = capi.seed_intrinsics(xs, ys, us, vs, off)
    o = O.estimate_intrinsics(xs, ys, us, vs, off)
    assert g["cam_success"][0] == 1 and g["view_success"].all()
    # intrinsics_estimate_test.cpp:39-44 tolerances against the ground truth
    assert np.abs(g["kmtx"][0, :4] - intr[:4]).max() < 1e-6 and abs(g["kmtx"][0, 4]) < 1e-9
    assert rel(g["kmtx"][0, :4], o["kmtx"][:4]) < 1e-9
    assert rel(g["hmtx"], o["hmtx"]) < 1e-9 and rel(g["poses"], o["poses"]) < 1e-8
    for i, T in enumerate(c_se3_t):
        est = O.pose12_to_T(g["poses"][i])
        assert min(np.abs(est[:3, :3] - T[:3, :3]).max(), np.abs(est[:3, :3] + T[:3, :3]).max()) < 1e-6


def test_too_few_views_fails_like_reference():
    _, _, (xs, ys, us, vs, off) = intrinsics_estimate_scenario(seed=5, n_frames=3, rows=5, cols=7, spacing=0.04, k=(800.0, 805.0, 320.0, 240.0))
    g = capi.seed_intrinsics(xs, ys, us, vs, off)
    assert g["cam_success"][0] == 0 and np.allclose(g["poses"][:, :9], np.eye(3).ravel())


def noisy_multicam(n_cams=3, n_poses=150, seed=3):
    prob, _, _ = synth.make_bundle(seed=seed, n_cams=n_cams, n_poses=n_poses)
    rng = np.random.default_rng(seed)
    nb = prob.desc.n_blocks
    keep = rng.integers(20, 89, size=nb); keep[5] = 3; keep[17] = 4; keep[40] = 0   # < 4 points: failed views
    pick = lambda b: np.array([0, 7, 40, 85]) if keep[b] == 4 else np.arange(keep[b])   # 4 points in general position
    idx = np.concatenate([prob.block_offset[b] + pick(b) for b in range(nb)]).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(keep)])
    return prob.x[idx], prob.y[idx], prob.u[idx], prob.v[idx], off, np.asarray(prob.block_cam), n_cams


def test_noisy_multicamera_ragged_views_match_oracle():
    xs, ys, us, vs, off, cam, n_cams = noisy_multicam()
    g = capi.seed_intrinsics(xs, ys, us, vs, off, view_cam=cam, n_cams=n_cams)
    for c in range(n_cams):
        sel = np.flatnonzero(cam == c)
        pieces = [np.arange(off[k], off[k + 1]) for k in sel]
        idx = np.concatenate(pieces).astype(np.int64)
        o = O.estimate_intrinsics(xs[idx], ys[idx], us[idx], vs[idx], np.concatenate([[0], np.cumsum([len(p) for p in pieces])]))
        assert o["success"] and g["cam_success"][c] == 1
        assert np.array_equal(g["view_success"][sel], o["view_success"])
        assert rel(g["kmtx"][c], o["kmtx"]) < 1e-7
        ok = o["view_success"] == 1
        assert rel(g["hmtx"][sel][ok], o["hmtx"][ok]) < 1e-8
        # symmetric_rms_px sums ROOTS of the residuals (intrinsicsdlt.cpp:21-31): an exactly fitted 4-point view
        # returns the square root of rounding noise (~1e-7), hence the absolute term
        assert np.allclose(g["sym_rms"][sel][ok], o["sym_rms"][ok], rtol=1e-8, atol=1e-6)
        assert rel(g["poses"][sel], o["poses"]) < 1e-7
    assert (g["view_success"] == 0).sum() == 2   # the 3-point and the empty view


def test_bounds_sanitize_like_reference():
    xs, ys, us, vs, off, cam, n_cams = noisy_multicam(n_cams=1, n_poses=60)
    bounds = [200.0, 2000.0, 150.0, 2000.0, 100.0, 200.0, 50.0, 75.0, -1.0, 1.0]
    g = capi.seed_intrinsics(xs, ys, us, vs, off, bounds=bounds)
    o = O.estimate_intrinsics(xs, ys, us, vs, off, bounds10=bounds)
    assert g["kmtx"][0, 2] == 150.0 and g["kmtx"][0, 3] == 62.5     # principal point outside the bounds -> midpoint
    assert rel(g["kmtx"][0], o["kmtx"]) < 1e-7 and rel(g["poses"], o["poses"]) < 1e-7


def test_intrinsics_with_homography_ransac_matches_oracle():
    """IntrinsicsEstimOptions::homography_ransac: every view through the batched RANSAC kernel (10 % of the corners
    of every view replaced by gross outliers), then Zhang and the pose decomposition — same inlier sets, K and poses
    as the oracle's per-view ransac<HomographyEstimator>."""
    prob, _, _ = synth.make_bundle(seed=5, n_cams=1, n_poses=40)
    rng = np.random.default_rng(9)
    off = np.asarray(prob.block_offset)
    u, v = prob.u.copy(), prob.v.copy()
    bad = rng.random(len(u)) < 0.10
    u[bad] = rng.uniform(0, 1280, bad.sum()); v[bad] = rng.uniform(0, 720, bad.sum())
    ro = abi.RansacOptions.default()
    g = capi.seed_intrinsics(prob.x, prob.y, u, v, off, ransac=ro)
    o = O.estimate_intrinsics_ransac(prob.x, prob.y, u, v, off, ro)
    assert o["success"] and g["cam_success"][0] == 1 and np.array_equal(g["view_success"], o["view_success"])
    assert np.array_equal(g["inlier_mask"], o["inlier_mask"])
    assert (g["inlier_mask"][bad] == 0).mean() > 0.95 and (g["inlier_mask"][~bad] == 1).mean() > 0.95
    assert rel(g["kmtx"][0], o["kmtx"]) < 1e-7 and rel(g["hmtx"], o["hmtx"]) < 1e-8
    assert np.allclose(g["sym_rms"], o["sym_rms"], rtol=1e-8, atol=1e-9) and rel(g["poses"], o["poses"]) < 1e-7


def test_homography_ransac_with_ragged_views_matches_oracle():
    """Views of different sizes (partly detected boards): grouped by size, one launch of the batched kernel per
    group, results and inlier masks scattered back to the views' own CSR layout.  A view of three points stays
    unsuccessful (intrinsicsdlt.cpp:41-45)."""
    prob, _, _ = synth.make_bundle(seed=6, n_cams=2, n_poses=30)
    rng = np.random.default_rng(10)
    off0 = np.asarray(prob.block_offset); cam0 = np.asarray(prob.block_cam)
    keep_n = rng.choice([88, 88, 80, 61, 40, 33, 3], size=len(off0) - 1)
    keep_n[5] = 3
    sel = np.concatenate([off0[k] + np.sort(rng.choice(off0[k + 1] - off0[k], keep_n[k], replace=False)) for k in range(len(keep_n))])
    off = np.concatenate([[0], np.cumsum(keep_n)]).astype(np.int64)
    x, y, u, v = prob.x[sel], prob.y[sel], prob.u[sel].copy(), prob.v[sel].copy()
    bad = rng.random(len(u)) < 0.08
    u[bad] = rng.uniform(0, 1280, bad.sum()); v[bad] = rng.uniform(0, 720, bad.sum())
    ro = abi.RansacOptions.default()
    g = capi.seed_intrinsics(x, y, u, v, off, view_cam=cam0, n_cams=2, ransac=ro)
    for c in range(2):
        vs = np.flatnonzero(cam0 == c)
        idx = np.concatenate([np.arange(off[k], off[k + 1]) for k in vs])
        offc = np.concatenate([[0], np.cumsum(keep_n[vs])]).astype(np.int64)
        o = O.estimate_intrinsics_ransac(x[idx], y[idx], u[idx], v[idx], offc, ro)
        assert o["success"] and g["cam_success"][c] == 1
        assert np.array_equal(g["view_success"][vs], o["view_success"])
        assert np.array_equal(g["inlier_mask"][idx], o["inlier_mask"])
        assert rel(g["kmtx"][c], o["kmtx"]) < 1e-7 and rel(g["hmtx"][vs], o["hmtx"]) < 1e-8
        assert np.allclose(g["sym_rms"][vs], o["sym_rms"], rtol=1e-8, atol=1e-9) and rel(g["poses"][vs], o["poses"]) < 1e-7
    assert g["view_success"][5] == 0 and g["view_success"].sum() >= len(keep_n) - (keep_n < 4).sum() - 2


def test_planar_poses_match_oracle():
    xs, ys, us, vs, off, cam, n_cams = noisy_multicam()
    kmtx = np.array([[1000.0, 1005.0, 640.0, 360.0, 0.0], [1010.0, 1015.0, 640.0, 360.0, 0.5], [990.0, 1000.0, 630.0, 350.0, 0.0]])
    poses, ok = capi.seed_planar_poses(xs, ys, us, vs, off, kmtx, view_cam=cam)
    for k in range(len(off) - 1):
        s = slice(off[k], off[k + 1])
        T = O.estimate_planar_pose(xs[s], ys[s], us[s], vs[s], kmtx[cam[k]])
        ref = np.concatenate([T[:3, :3].ravel(), T[:3, 3]])
        assert rel(poses[k], ref) < 1e-8, k
        assert ok[k] == (off[k + 1] - off[k] >= 4)


def test_device_resident_inputs_and_scale():
    """100 000 views (8.8 M observations): device pointers are accepted as they are; K lands near the
    ground truth of the synthetic cameras, every view yields a pose in front of the camera."""
    import torch
    prob, _, xgt = synth.make_bundle(seed=137, n_cams=8, n_poses=12500)
    off = np.asarray(prob.block_offset); cam = np.asarray(prob.block_cam)
    host = capi.seed_intrinsics(prob.x, prob.y, prob.u, prob.v, off, view_cam=cam, n_cams=8)
    dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (prob.x, prob.y, prob.u, prob.v)]
    import ctypes as C
    nv = len(off) - 1
    kmtx = np.zeros((8, 5)); cam_ok = np.zeros(8, dtype=np.int32); poses = np.zeros((nv, 12))
    opts = abi.SeedOptions.from_bounds(None)
    rc = capi.lib().cal_seed_intrinsics(nv, abi.i64ptr(off.astype(np.int64)), abi.i32ptr(cam.astype(np.int32)),
                                        *[C.cast(C.c_void_p(d.data_ptr()), abi.c_double_p) for d in dev], 8, C.byref(opts), 0,
                                        abi.dptr(kmtx), abi.i32ptr(cam_ok), None, None, None, abi.dptr(poses))
    assert rc == 0 and np.array_equal(kmtx, host["kmtx"]) and np.array_equal(poses, host["poses"])
    assert host["cam_success"].all() and host["view_success"].all()
    k_gt = xgt[:80].reshape(8, 10)[:, :5]
    assert np.abs(host["kmtx"][:, :4] / k_gt[:, :4] - 1).max() < 0.05   # distortion + 0.2 px noise: a seed, not the answer
    assert (host["poses"][:, 11] > 0).all()


def test_json_to_calibration_chain(tmp_path):
    """The callers' path end to end (facades/intrinsics.cpp:80-135): PlanarDetections JSON -> columnar store ->
    estimate_intrinsics on the GPU -> optimize_intrinsics on the GPU.  Noise-free views of a distortion-free
    camera: the linear seed is already exact to 1e-6 and the refinement keeps it there (intrinsics_optimize_test
    tolerances)."""
    import json
    from calibration_b200 import geometry as G
    intr, c_se3_t, (xs, ys, us, vs, off) = intrinsics_estimate_scenario(n_frames=12, rows=7, cols=10)
    doc = {"sensor_id": "cam0", "images": [
        {"file": f"v{k}.png", "points": [{"x": float(us[i]), "y": float(vs[i]), "local_x": float(xs[i]), "local_y": float(ys[i])}
                                         for i in range(off[k], off[k + 1])]} for k in range(len(off) - 1)]}
    (tmp_path / "cam0.json").write_text(json.dumps(doc))
    nv, no = capi.Dataset.from_planar_json([str(tmp_path / "cam0.json")], str(tmp_path / "obs.calobs"), min_corners_per_view=8)
    assert (nv, no) == (int((np.diff(off) >= 8).sum()), len(xs))   # views the board left are dropped (collect_planar_views)
    with capi.Dataset(str(tmp_path / "obs.calobs"), pin=True) as ds:
        seed = capi.seed_intrinsics(ds.x, ds.y, ds.u, ds.v, ds.view_offset, view_cam=ds.view_cam, n_cams=ds.n_cams)
        assert seed["cam_success"][0] and np.abs(seed["kmtx"][0, :4] - intr[:4]).max() < 1e-6
        prob = abi.Problem(abi.KIND_INTRINSICS, abi.MODEL_PINHOLE_BC5, 1, ds.n_views, ds.x, ds.y, ds.u, ds.v, ds.view_offset,
                           np.zeros(ds.n_views, dtype=np.int32), huber_delta=1.0)
        intr0 = np.concatenate([seed["kmtx"][0], np.zeros(5)])
        x0 = G.pack_intrinsics(intr0, [O.pose12_to_T(p) for p in seed["poses"]])
        h = capi.RefineHandle(prob)
        try:
            x, res, _ = h.solve(x0)
            rms, g = h.view_errors(x)
        finally:
            h.close()
    assert res.success and np.abs(x[:4] - intr[:4]).max() < 1e-6 and np.abs(x[5:10]).max() < 1e-5 and g < 1e-6
