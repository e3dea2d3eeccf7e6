"""Oracle plane fit against the reference's tests (tests/unit/planefit_test.cpp) and, where the reference tree
or a prebuilt oracle/_ref exists, against the reference's own ransac<> loop bit for bit."""
import numpy as np
import pytest

import oracle_lib as O
from calibration_b200 import abi, synth


def align(p, ref):
    return -p if np.dot(p[:3], ref[:3]) < 0 else p          # planefit_test.cpp:18-20


def test_three_sample_stream_is_the_real_std_sample():
    for seed, n in [(1234567, 140), (1234567, 3), (7, 4), (2 ** 63 + 5, 1000), (99, 33)]:
        assert np.array_equal(O.sample_stream_k(seed, n, 3, 64), O.sample_stream_k(seed, n, 3, 64, real=True)), (seed, n)


def test_svd_matches_ideal_plane():
    # PlaneFit.SvdMatchesIdealPlane (:77-96): an 11 x 11 grid on the plane, 1e-9
    n = np.array([0.4, 0.1, 1.0]); n /= np.linalg.norm(n)
    gt = np.array([*n, -n @ np.array([0.5, -0.2, 0.8])])
    g = np.arange(-5, 6) * 0.1
    x, y = (a.ravel() for a in np.meshgrid(g, g, indexing="ij"))
    z = (-gt[3] - gt[0] * x - gt[1] * y) / gt[2]
    rc, p = O.fit_plane_svd(x, y, z)
    assert rc == 0 and np.abs(align(p, gt) - gt).max() < 1e-9
    assert O.fit_plane_svd(x[:2], y[:2], z[:2])[0] == 1      # "Not enough points to fit a plane" (:67-69)


def test_ransac_rejects_outliers():
    # PlaneFit.RansacRejectsOutliers (:22-75)
    gt, xyz = O.plane_testdata()
    opts = abi.RansacOptions.default(max_iters=2000, thresh=0.01, min_inliers=80, confidence=0.999)
    res, mask = O.ransac_plane(*xyz.T, opts)
    assert res.success and res.n_inliers >= 100 and mask.sum() == res.n_inliers
    p = np.array(res.plane)
    assert np.abs(align(p, gt) - gt).max() < 1e-3
    assert res.inlier_rms < 1e-3
    r = np.abs(xyz @ p[:3] + p[3])
    assert np.all(r[mask == 1] < opts.thresh)


def test_failure_and_tiny_inputs():
    res, mask = O.ransac_plane([0.0, 1.0], [0.0, 0.0], [0.0, 0.0])          # fewer than three points (:88-90)
    assert not res.success and list(res.plane) == [0.0] * 4 and res.iters == 0
    t = np.linspace(0, 1, 30)
    res, _ = O.ransac_plane(t, 2 * t, -t, abi.RansacOptions.default(min_inliers=3, max_iters=40))   # collinear: every sample degenerate
    assert not res.success and res.iters_run == 40
    rng = np.random.default_rng(0)
    res, _ = O.ransac_plane(*rng.uniform(-1, 1, (3, 50)), abi.RansacOptions.default(thresh=1e-3, min_inliers=20, max_iters=100))
    assert not res.success                                                   # no plane in noise


def test_batch_equals_single():
    x, y, z, _ = synth.synth_plane_ransac(seed=4, n_problems=6, n=120)
    opts = abi.RansacOptions.default(thresh=0.006, min_inliers=30)
    rb, mb = O.ransac_plane_batch(x, y, z, opts)
    for p in range(6):
        r1, m1 = O.ransac_plane(x[p], y[p], z[p], abi.RansacOptions.default(thresh=0.006, min_inliers=30, seed=1234567 + p))
        assert list(r1.plane) == list(rb[p].plane) and np.array_equal(m1, mb[p]) and r1.iters == rb[p].iters
        assert r1.success and 60 < r1.n_inliers <= 120


# ---- the reference's own loop (oracle/_ref) ----
needs_ref = pytest.mark.skipif(O.ref_lib() is None, reason="no /root/reference and no prebuilt oracle/_ref")


def same(opts, x, y, z):
    rr, mr = O.ref_ransac_plane(x, y, z, opts)
    ro, mo = O.ref_ransac_plane(x, y, z, opts, oracle_twin=True)
    assert bool(ro.success) == bool(rr.success) and ro.iters == rr.iters and ro.n_inliers == rr.n_inliers
    assert np.array_equal(mo, mr) and list(ro.plane) == list(rr.plane)
    assert ro.inlier_rms == rr.inlier_rms
    rs, ms = O.ransac_plane(x, y, z, opts)
    if rs.min_margin > 1e-12:
        assert bool(rs.success) == bool(rr.success) and np.array_equal(ms, mr)
    return ro


@needs_ref
@pytest.mark.parametrize("refit", [1, 0])
def test_plane_loop_matches_reference_template(refit):
    gt, xyz = O.plane_testdata()
    r = same(abi.RansacOptions.default(max_iters=2000, thresh=0.01, min_inliers=80, confidence=0.999, refit_on_inliers=refit), *xyz.T)
    assert r.success
    x, y, z, _ = synth.synth_plane_ransac(seed=9, n_problems=8, n=150)
    for p in range(8):
        for conf, mi in ((0.99, 1000), (0.0, 60), (1.0, 30), (0.999999, 400)):
            same(abi.RansacOptions.default(seed=31 * p + 1, thresh=0.006, min_inliers=20, confidence=conf, max_iters=mi,
                                           refit_on_inliers=refit), x[p], y[p], z[p])


@needs_ref
def test_plane_edge_cases_match_reference_template():
    t = np.linspace(0, 1, 30)
    same(abi.RansacOptions.default(min_inliers=3, max_iters=40), t, 2 * t, -t)
    same(abi.RansacOptions.default(min_inliers=1), [0.0, 1.0], [0.0, 0.0], [0.0, 0.0])
    q = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], float)
    r = same(abi.RansacOptions.default(min_inliers=3, thresh=1e-9), *q.T)
    assert r.success and r.n_inliers == 3
    same(abi.RansacOptions.default(min_inliers=3, thresh=1e-9, max_iters=25), *np.vstack([q, q[:2]]).T)
