"""The product's RANSAC kernels — the very CUDA sources the GPU runs (csrc/ransac_kernel.cuh, ransac_plane_kernel.cuh,
ransac_sampler.cuh, dlt.cuh, plane_math.cuh) — compiled by g++ and executed on the CPU under a lock-step SIMT shim
(tests/host_emul/simt_shim.hpp: one OS thread per CUDA thread, warp collectives as rendezvous), against the oracle:
the same assertions as the GPU parity tests, at sizes the shim handles in seconds.  No GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from calibration_b200 import abi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_emul", "ransac_simt.cpp")
SO = os.path.join(ROOT, "tests", "host_emul", "_build", "libransac_simt.so")
CSRC = os.path.join(ROOT, "calibration_b200", "csrc")


@pytest.fixture(scope="module")
def simt():
    deps = [SRC, os.path.join(ROOT, "tests", "host_emul", "simt_shim.hpp")] + [
        os.path.join(CSRC, f) for f in ("ransac_kernel.cuh", "ransac_plane_kernel.cuh", "ransac_sampler.cuh", "ransac_iters.hpp", "dlt.cuh", "plane_math.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++20", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-o", SO, SRC], check=True)
    L = C.CDLL(SO)
    dp = abi.c_double_p
    L.simt_ransac_homography.argtypes = [C.c_int64, C.c_int32, dp, dp, dp, dp, C.POINTER(abi.RansacOptions), C.c_int,
                                         C.POINTER(abi.RansacResult), abi.c_uint8_p]
    L.simt_ransac_plane.argtypes = [C.c_int64, C.c_int32, dp, dp, dp, C.POINTER(abi.RansacOptions), C.c_int, C.POINTER(abi.PlaneResult),
                                    abi.c_uint8_p]
    return L


def run_h(L, x, y, u, v, opts, seed_per_problem=True):
    x, y, u, v = (abi.as_f64(a) for a in (x, y, u, v))
    npb, n = x.shape
    res = (abi.RansacResult * npb)(); mask = np.zeros((npb, n), dtype=np.uint8)
    assert L.simt_ransac_homography(npb, n, abi.dptr(x), abi.dptr(y), abi.dptr(u), abi.dptr(v), C.byref(opts), int(seed_per_problem), res,
                                    mask.ctypes.data_as(abi.c_uint8_p)) == 0
    return res, mask


def run_p(L, x, y, z, opts, seed_per_problem=True):
    x, y, z = (abi.as_f64(a) for a in (x, y, z))
    npb, n = x.shape
    res = (abi.PlaneResult * npb)(); mask = np.zeros((npb, n), dtype=np.uint8)
    assert L.simt_ransac_plane(npb, n, abi.dptr(x), abi.dptr(y), abi.dptr(z), C.byref(opts), int(seed_per_problem), res,
                               mask.ctypes.data_as(abi.c_uint8_p)) == 0
    return res, mask


def compare_h(L, x, y, u, v, opts, seed_per_problem=True, min_margin=1e-9):
    ro, mo = O.ransac_batch(x, y, u, v, opts, seed_per_problem)
    rg, mg = run_h(L, x, y, u, v, opts, seed_per_problem)
    n_checked = 0
    for p in range(x.shape[0]):
        if ro[p].min_margin <= min_margin:
            continue
        n_checked += 1
        assert rg[p].success == ro[p].success, p
        assert np.array_equal(mg[p], mo[p]), p
        assert rg[p].n_inliers == ro[p].n_inliers and rg[p].iters == ro[p].iters and rg[p].iters_run == ro[p].iters_run, p
        if ro[p].success:
            Ho, Hg = np.array(ro[p].hmtx), np.array(rg[p].hmtx)
            assert np.abs(Hg - Ho).max() <= 1e-7 * np.abs(Ho).max(), p
            assert abs(rg[p].inlier_rms - ro[p].inlier_rms) <= 1e-8 * ro[p].inlier_rms + 1e-9
            assert abs(rg[p].symmetric_rms_px - ro[p].symmetric_rms_px) <= 1e-8 * ro[p].symmetric_rms_px + 1e-5
    return n_checked, rg


def compare_p(L, x, y, z, opts, seed_per_problem=True, min_margin=1e-10):
    ro, mo = O.ransac_plane_batch(x, y, z, opts, seed_per_problem)
    rg, mg = run_p(L, x, y, z, opts, seed_per_problem)
    n_checked = 0
    for p in range(x.shape[0]):
        if ro[p].min_margin <= min_margin:
            continue
        n_checked += 1
        assert rg[p].success == ro[p].success and np.array_equal(mg[p], mo[p]), p
        assert rg[p].n_inliers == ro[p].n_inliers and rg[p].iters == ro[p].iters and rg[p].iters_run == ro[p].iters_run, p
        Po, Pg = np.array(ro[p].plane), np.array(rg[p].plane)
        if ro[p].success:
            assert np.abs(Pg - Po).max() <= 1e-9, p
            assert abs(rg[p].inlier_rms - ro[p].inlier_rms) <= 1e-8 * ro[p].inlier_rms + 1e-12
        else:
            assert list(Pg) == [0.0] * 4 and not mg[p].any()
    return n_checked, rg


def test_homography_kernel_source_matches_oracle(simt):
    x, y, u, v, _ = synth.synth_ransac(seed=17, n_problems=24, n=200)
    n_checked, rg = compare_h(simt, x, y, u, v, abi.RansacOptions.default())
    assert n_checked >= 23 and sum(r.success for r in rg) == 24


@pytest.mark.parametrize("n", [4, 5, 31, 33, 130])
def test_homography_kernel_ragged_sizes(simt, n):
    x, y, u, v, _ = synth.synth_ransac(seed=n, n_problems=6, n=n, outlier_fraction=0.2)
    n_checked, _ = compare_h(simt, x, y, u, v, abi.RansacOptions.default(min_inliers=min(12, n), max_iters=120))
    assert n_checked >= 5


def test_homography_kernel_option_variants(simt):
    x, y, u, v, _ = synth.synth_ransac(seed=5, n_problems=8, n=150, outlier_fraction=0.5)
    for opts in (abi.RansacOptions.default(refit_on_inliers=0), abi.RansacOptions.default(min_inliers=125, max_iters=100),
                 abi.RansacOptions.default(thresh=0.8, confidence=0.999), abi.RansacOptions.default(confidence=0.0, max_iters=40)):
        n_checked, _ = compare_h(simt, x, y, u, v, opts)
        assert n_checked >= 7
    n_checked, _ = compare_h(simt, x, y, u, v, abi.RansacOptions.default(seed=99), seed_per_problem=False)
    assert n_checked >= 7
    # homography_test.cpp:104-134 and :137-160
    _, d = O.homography_testdata(100, 0.0, 30, 7)
    _, rg = compare_h(simt, *[c[None] for c in d.T], abi.RansacOptions.default(thresh=1.0, min_inliers=90, seed=123), seed_per_problem=False)
    assert rg[0].success and rg[0].n_inliers >= 95
    _, d = O.homography_testdata(4, 0.0, 50, 3)
    _, rg = compare_h(simt, *[c[None] for c in d.T], abi.RansacOptions.default(thresh=0.5, min_inliers=10, seed=42), seed_per_problem=False)
    assert not rg[0].success


def test_plane_kernel_source_matches_oracle(simt):
    x, y, z, _ = synth.synth_plane_ransac(seed=23, n_problems=24, n=200)
    n_checked, rg = compare_p(simt, x, y, z, abi.RansacOptions.default(thresh=0.006))
    assert n_checked >= 23 and sum(r.success for r in rg) == 24
    gt, xyz = O.plane_testdata()      # planefit_test.cpp:22-75
    _, rg = compare_p(simt, *[c[None] for c in xyz.T], abi.RansacOptions.default(max_iters=2000, thresh=0.01, min_inliers=80, confidence=0.999),
                      seed_per_problem=False)
    assert rg[0].success and rg[0].n_inliers >= 100


@pytest.mark.parametrize("n", [3, 4, 31, 33, 130])
def test_plane_kernel_ragged_sizes_and_variants(simt, n):
    x, y, z, _ = synth.synth_plane_ransac(seed=n, n_problems=6, n=n, outlier_fraction=0.2)
    for kw in (dict(max_iters=120), dict(max_iters=120, refit_on_inliers=0), dict(confidence=0.0, max_iters=50)):
        n_checked, _ = compare_p(simt, x, y, z, abi.RansacOptions.default(thresh=0.006, min_inliers=min(12, n), **kw))
        assert n_checked >= 5


def test_plane_kernel_degenerate_inputs(simt):
    t = np.linspace(0, 1, 40)[None].repeat(2, 0)
    _, rg = compare_p(simt, t, 2 * t, -t, abi.RansacOptions.default(min_inliers=3, max_iters=40))
    assert not any(r.success for r in rg) and all(r.iters_run == 40 for r in rg)
    _, rg = compare_p(simt, np.zeros((2, 2)), np.ones((2, 2)), np.zeros((2, 2)), abi.RansacOptions.default(min_inliers=1))
    assert not any(r.success for r in rg) and all(r.iters_run == 0 for r in rg)
