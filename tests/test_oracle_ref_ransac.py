"""The oracle's RANSAC loop against the REFERENCE's own calib::ransac<> template.

oracle/_ref/libref_ransac.so is the reference's common/ransac.h:121-194 compiled unmodified from where it
lies in /root/reference (oracle/ref_ransac_harness.cpp, `make -C oracle ref`) around the oracle's estimator
hooks, so every difference below would be a difference in the loop itself: the std::sample stream, the
`continue`s, refit, tie-break, adaptive iteration count, result bookkeeping.  The library also holds
oracle/ransac.cpp compiled with the same flags (no FMA contraction), and that twin must agree with the
template BIT FOR BIT; the shipped liboracle.so (contraction on) must agree on everything discrete.
Skipped where neither the reference tree nor a prebuilt library exists."""
import numpy as np
import pytest

import oracle_lib as O
from calibration_b200 import abi, synth

pytestmark = pytest.mark.skipif(O.ref_lib() is None, reason="no /root/reference and no prebuilt oracle/_ref")


def same(opts, x, y, u, v):
    rr, mr = O.ref_ransac(x, y, u, v, opts)                       # the reference's template
    ro, mo = O.ref_ransac(x, y, u, v, opts, oracle_twin=True)     # the oracle's loop, same compilation flags
    assert bool(ro.success) == bool(rr.success)
    assert ro.iters == rr.iters and ro.n_inliers == rr.n_inliers
    assert np.array_equal(mo, mr)
    if ro.success:
        assert list(ro.hmtx) == list(rr.hmtx)          # bit for bit
        assert ro.inlier_rms == rr.inlier_rms
    rs, ms = O.ransac(x, y, u, v, opts)                           # the oracle as the other tests use it
    if rs.min_margin > 1e-9:
        assert bool(rs.success) == bool(rr.success) and np.array_equal(ms, mr) and rs.n_inliers == rr.n_inliers
        if rs.success and rs.inlier_rms > 1e-9:                   # an rms tie at rounding level may pick the other model
            assert rs.iters == rr.iters
            assert np.allclose(rs.hmtx, rr.hmtx, rtol=1e-9, atol=1e-9 * np.abs(rr.hmtx).max())
    return ro


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("refit", [1, 0])
def test_loop_matches_reference_template(seed, refit):
    x, y, u, v, _ = synth.synth_ransac(seed=10 + seed, n_problems=3, n=160)
    for p in range(3):
        opts = abi.RansacOptions.default(seed=99 + 7 * seed + p, refit_on_inliers=refit)
        r = same(opts, x[p], y[p], u[p], v[p])
        assert r.success


def test_reference_test_scenarios():
    # homography_test.cpp:104-134 and :137-160
    _, d = O.homography_testdata(100, 0.0, 30, 7)
    r = same(abi.RansacOptions.default(thresh=1.0, min_inliers=90, seed=123), *d.T)
    assert r.success and r.n_inliers >= 95
    _, d = O.homography_testdata(4, 0.0, 50, 3)
    r = same(abi.RansacOptions.default(thresh=0.5, min_inliers=10, seed=42), *d.T)
    assert not r.success


@pytest.mark.parametrize("conf,max_iters,min_inl", [(0.0, 40, 12), (0.999999, 300, 12), (0.5, 1000, 4), (0.99, 1, 4),
                                                    (0.99, 200, 150), (1.0, 50, 12)])
def test_iteration_control(conf, max_iters, min_inl):
    # calculate_iterations (ransac.h:64-78): confidence <= 0, clamp to iters_so_far, log(0) at confidence 1
    x, y, u, v, _ = synth.synth_ransac(seed=4, n_problems=2, n=200)
    for p in range(2):
        same(abi.RansacOptions.default(seed=5 + p, confidence=conf, max_iters=max_iters, min_inliers=min_inl),
             x[p], y[p], u[p], v[p])


def test_degenerate_and_tiny_inputs():
    rng = np.random.default_rng(0)
    # all object points on one line: every sample is degenerate (homographyestimator.cpp:100-119)
    t = rng.uniform(-1, 1, 40)
    r = same(abi.RansacOptions.default(min_inliers=4, max_iters=50), t, 2 * t + 1, rng.uniform(0, 100, 40), rng.uniform(0, 100, 40))
    assert not r.success
    # fewer than k_min_samples (ransac.h:127-129), exactly four, repeated points
    same(abi.RansacOptions.default(min_inliers=1), [0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [10.0, 11.0, 10.0], [0.0, 0.0, 1.0])
    q = np.array([[0, 0, 5, 5], [1, 0, 9, 5.5], [1, 1, 9.5, 10], [0, 1, 4.5, 9]], float)
    r = same(abi.RansacOptions.default(min_inliers=4, thresh=1e-6), *q.T)
    assert r.success and r.n_inliers == 4
    q2 = np.vstack([q, q[:2], q[:1]])
    same(abi.RansacOptions.default(min_inliers=4, thresh=1e-6, max_iters=30), *q2.T)


def test_ties_take_the_lower_rms():
    # is_better_model (ransac.h:113-117): equal counts -> strictly smaller rms wins; noise makes ties frequent
    x, y, u, v, _ = synth.synth_ransac(seed=21, n_problems=4, n=24)
    for p in range(4):
        same(abi.RansacOptions.default(seed=p, min_inliers=6, thresh=0.7, confidence=0.0, max_iters=400), x[p], y[p], u[p], v[p])
