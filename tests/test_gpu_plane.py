"""Parity of the batched plane RANSAC kernel (fit_plane_ransac, linear/planefit.cpp:86-104) with the CPU oracle:
bit-exact inlier masks and loop counters, planes to 1e-9."""
import numpy as np
import pytest

import oracle_lib as O
from calibration_b200 import abi, capi, synth

pytestmark = pytest.mark.gpu


def compare(x, y, z, opts, seed_per_problem=True, min_margin=1e-10):
    ro, mo = O.ransac_plane_batch(x, y, z, opts, seed_per_problem)
    rg, mg = capi.ransac_plane_batch(x, y, z, opts, seed_per_problem)
    n_checked = 0
    for p in range(x.shape[0]):
        if ro[p].min_margin <= min_margin:   # a distance within rounding of the threshold: not comparable bit for bit
            continue
        n_checked += 1
        assert rg[p].success == ro[p].success, p
        assert np.array_equal(mg[p], mo[p]), p                    # bit-exact inlier mask
        assert rg[p].n_inliers == ro[p].n_inliers and rg[p].iters == ro[p].iters and rg[p].iters_run == ro[p].iters_run, p
        Po, Pg = np.array(ro[p].plane), np.array(rg[p].plane)
        if ro[p].success:
            if np.dot(Pg[:3], Po[:3]) < 0 and np.sort(np.abs(Po[:3]))[-1] - np.sort(np.abs(Po[:3]))[-2] < 1e-9:
                Pg = -Pg                                           # the sign rule is ambiguous when two components tie
            assert np.abs(Pg - Po).max() <= 1e-9, p
            assert abs(np.linalg.norm(Pg[:3]) - 1.0) <= 1e-12
            assert abs(rg[p].inlier_rms - ro[p].inlier_rms) <= 1e-8 * ro[p].inlier_rms + 1e-12
        else:
            assert list(Pg) == [0.0] * 4 and not mg[p].any()       # PlaneRansacResult defaults (planefit.h:14-19)
    return n_checked, ro, rg


def test_batched_plane_ransac_bit_exact_masks():
    x, y, z, planes = synth.synth_plane_ransac(seed=23, n_problems=512, n=500)
    n_checked, ro, rg = compare(x, y, z, abi.RansacOptions.default(thresh=0.006))
    assert n_checked >= 0.98 * 512
    assert sum(r.success for r in rg) == 512
    assert 300 < np.mean([r.n_inliers for r in rg]) < 380
    for p in range(0, 512, 37):
        P = np.array(rg[p].plane)
        P = -P if np.dot(P[:3], planes[p, :3]) < 0 else P
        assert np.abs(P - planes[p]).max() < 1e-2                 # recovers the generating plane (outliers near the plane bias it)


def test_reference_test_scenario():
    # PlaneFit.RansacRejectsOutliers (planefit_test.cpp:22-75), replicated 8 times with the same seed
    gt, xyz = O.plane_testdata()
    x, y, z = (np.tile(c, (8, 1)) for c in xyz.T)
    opts = abi.RansacOptions.default(max_iters=2000, thresh=0.01, min_inliers=80, confidence=0.999)
    n_checked, ro, rg = compare(x, y, z, opts, seed_per_problem=False)
    assert n_checked == 8
    for r in rg:
        P = np.array(r.plane); P = -P if np.dot(P[:3], gt[:3]) < 0 else P
        assert r.success and r.n_inliers >= 100 and np.abs(P - gt).max() < 1e-3 and r.inlier_rms < 1e-3


@pytest.mark.parametrize("n", [3, 4, 31, 33, 130, 257])
def test_ragged_sizes(n):
    x, y, z, _ = synth.synth_plane_ransac(seed=n, n_problems=40, n=n, outlier_fraction=0.2)
    n_checked, _, _ = compare(x, y, z, abi.RansacOptions.default(thresh=0.006, min_inliers=min(12, n), max_iters=200))
    assert n_checked >= 36


def test_options_variants_and_failures():
    x, y, z, _ = synth.synth_plane_ransac(seed=5, n_problems=64, n=300, outlier_fraction=0.5)
    for opts in (abi.RansacOptions.default(thresh=0.006, refit_on_inliers=0), abi.RansacOptions.default(thresh=0.006, min_inliers=250, max_iters=150),
                 abi.RansacOptions.default(thresh=0.003, confidence=0.999), abi.RansacOptions.default(thresh=0.006, confidence=0.0, max_iters=40)):
        n_checked, _, _ = compare(x, y, z, opts)
        assert n_checked >= 60
    n_checked, _, _ = compare(x, y, z, abi.RansacOptions.default(thresh=0.006, seed=99), seed_per_problem=False)
    assert n_checked >= 60
    # collinear points: every sample degenerate; two points: below k_min_samples (planefit.cpp:88-90)
    t = np.linspace(0, 1, 40)[None].repeat(4, 0)
    _, ro, rg = compare(t, 2 * t, -t, abi.RansacOptions.default(min_inliers=3, max_iters=40))
    assert not any(r.success for r in rg) and all(r.iters_run == 40 for r in rg)
    _, ro, rg = compare(np.zeros((2, 2)), np.ones((2, 2)), np.zeros((2, 2)), abi.RansacOptions.default(min_inliers=1))
    assert not any(r.success for r in rg) and all(r.iters_run == 0 for r in rg)


def test_run_to_run_identical():
    x, y, z, _ = synth.synth_plane_ransac(seed=8, n_problems=256, n=500)
    a, ma = capi.ransac_plane_batch(x, y, z, abi.RansacOptions.default(thresh=0.006))
    b, mb = capi.ransac_plane_batch(x, y, z, abi.RansacOptions.default(thresh=0.006))
    assert bytes(a) == bytes(b) and np.array_equal(ma, mb)
