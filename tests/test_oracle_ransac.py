"""Oracle RANSAC against the reference's homography tests (homography_test.cpp:50-160)."""
import numpy as np

import oracle_lib as O
from calibration_b200 import abi


def test_exact_homography_dlt():
    # HomographyTest.ExactHomography (:50-66): DLT on 50 exact correspondences
    H, xyuv = O.homography_testdata(50, 0.0, 0, 1)
    rc, Hd = O.homography_dlt(xyuv[:, 0], xyuv[:, 1], xyuv[:, 2], xyuv[:, 3])
    assert rc == 0
    assert np.allclose(Hd / Hd[2, 2], H, rtol=1e-6, atol=1e-6 * np.abs(H).max())


def test_ransac_recovers_homography_with_outliers():
    # :104-134 — 100 exact inliers (mt19937(42)) + 30 outliers (mt19937(7)); thresh 1, min_inliers 90, seed 123
    H, xyuv = O.homography_testdata(100, 0.0, 30, 7)
    opts = abi.RansacOptions.default(thresh=1.0, min_inliers=90, seed=123)
    res, mask = O.ransac(xyuv[:, 0], xyuv[:, 1], xyuv[:, 2], xyuv[:, 3], opts)
    assert res.success
    assert res.n_inliers >= 95 and mask.sum() == res.n_inliers
    assert res.symmetric_rms_px < 1e-3
    Hr = np.array(res.hmtx).reshape(3, 3)
    assert np.allclose(Hr / Hr[2, 2], H, rtol=1e-2, atol=1e-2)


def test_ransac_fails_with_too_few_inliers():
    # :137-160
    H, xyuv = O.homography_testdata(4, 0.0, 50, 3)
    opts = abi.RansacOptions.default(thresh=0.5, min_inliers=10, seed=42)
    res, mask = O.ransac(xyuv[:, 0], xyuv[:, 1], xyuv[:, 2], xyuv[:, 3], opts)
    assert not res.success and mask.sum() == 0


def test_fewer_than_four_points():
    res, mask = O.ransac([0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [10.0, 11.0, 10.0], [0.0, 0.0, 1.0])
    assert not res.success and res.iters == 0  # ransac.h:127-129


def test_precomputed_samples_equal_internal_stream():
    from calibration_b200 import synth
    x, y, u, v, _ = synth.synth_ransac(seed=3, n_problems=4, n=200)
    for p in range(4):
        opts = abi.RansacOptions.default(seed=1234567 + p)
        r1, m1 = O.ransac(x[p], y[p], u[p], v[p], opts)
        idx = O.sample_stream(1234567 + p, 200, 1000)
        r2, m2 = O.ransac(x[p], y[p], u[p], v[p], opts, sample_idx=idx)
        assert np.array_equal(m1, m2) and r1.iters == r2.iters and list(r1.hmtx) == list(r2.hmtx)
        assert r1.success and 100 < r1.n_inliers < 170
