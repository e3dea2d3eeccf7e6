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


# ---------------------------------------------------------------------------
# seeding kernels (csrc/seed_kernels.cuh + seed_host.hpp) under the same shim
# ---------------------------------------------------------------------------
SEED_SRC = os.path.join(ROOT, "tests", "host_emul", "seed_simt.cpp")
SEED_SO = os.path.join(ROOT, "tests", "host_emul", "_build", "libseed_simt.so")


@pytest.fixture(scope="module")
def seed_simt():
    deps = [SEED_SRC, os.path.join(ROOT, "tests", "host_emul", "simt_shim.hpp")] + [
        os.path.join(CSRC, f) for f in ("seed_kernels.cuh", "seed_host.hpp", "dlt.cuh")]
    if not os.path.exists(SEED_SO) or any(os.path.getmtime(d) > os.path.getmtime(SEED_SO) for d in deps):
        os.makedirs(os.path.dirname(SEED_SO), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++20", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-o", SEED_SO, SEED_SRC], check=True)
    L = C.CDLL(SEED_SO)
    dp, ip = abi.c_double_p, abi.c_int32_p
    L.simt_seed_intrinsics.argtypes = [C.c_int64, abi.c_int64_p, ip, dp, dp, dp, dp, C.c_int32, C.POINTER(abi.SeedOptions), dp, ip, ip, dp, dp, dp]
    L.simt_seed_planar_poses.argtypes = [C.c_int64, abi.c_int64_p, ip, dp, dp, dp, dp, dp, dp, ip]
    L.simt_gather_scatter_roundtrip.argtypes = [C.c_int64, C.c_int32, abi.c_int64_p, abi.c_int64_p, dp, dp, dp, dp, dp, dp, dp, dp,
                                                C.POINTER(abi.RansacResult), abi.c_uint8_p, C.POINTER(abi.RansacResult), abi.c_uint8_p]
    return L


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))


def simt_seed_intrinsics(L, x, y, u, v, off, cam=None, n_cams=1, bounds=None):
    x, y, u, v = (abi.as_f64(a) for a in (x, y, u, v))
    off = np.ascontiguousarray(off, dtype=np.int64); nv = len(off) - 1
    cam = np.zeros(nv, dtype=np.int32) if cam is None else np.ascontiguousarray(cam, dtype=np.int32)
    opts = abi.SeedOptions.from_bounds(bounds)
    kmtx = np.zeros((n_cams, 5)); cam_ok = np.zeros(n_cams, dtype=np.int32); ok = np.zeros(nv, dtype=np.int32)
    H = np.zeros((nv, 9)); rms = np.zeros(nv); poses = np.zeros((nv, 12))
    assert L.simt_seed_intrinsics(nv, abi.i64ptr(off), abi.i32ptr(cam), abi.dptr(x), abi.dptr(y), abi.dptr(u), abi.dptr(v), n_cams, C.byref(opts),
                                  abi.dptr(kmtx), abi.i32ptr(cam_ok), abi.i32ptr(ok), abi.dptr(H), abi.dptr(rms), abi.dptr(poses)) == 0
    return dict(kmtx=kmtx, cam_success=cam_ok, view_success=ok, hmtx=H.reshape(nv, 3, 3), sym_rms=rms, poses=poses)


def noisy_multicam(n_cams=2, n_poses=60, seed=3):
    prob, _, _ = synth.make_bundle(seed=seed, n_cams=n_cams, n_poses=n_poses)
    rng = np.random.default_rng(seed)
    nb = prob.desc.n_blocks
    keep = rng.integers(20, 89, size=nb); keep[5] = 3; keep[17] = 4; keep[40] = 0   # < 4 points: failed views
    pick = lambda b: np.array([0, 7, 40, 85]) if keep[b] == 4 else np.arange(keep[b])
    idx = np.concatenate([prob.block_offset[b] + pick(b) for b in range(nb)]).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(keep)])
    return prob.x[idx], prob.y[idx], prob.u[idx], prob.v[idx], off, np.asarray(prob.block_cam), n_cams


def test_seed_kernels_reference_scenario(seed_simt):
    from test_oracle_seed import intrinsics_estimate_scenario
    intr, c_se3_t, (xs, ys, us, vs, off) = intrinsics_estimate_scenario()
    g = simt_seed_intrinsics(seed_simt, xs, ys, us, vs, off)
    o = O.estimate_intrinsics(xs, ys, us, vs, off)
    assert g["cam_success"][0] == 1 and g["view_success"].all()
    assert np.abs(g["kmtx"][0, :4] - intr[:4]).max() < 1e-6 and abs(g["kmtx"][0, 4]) < 1e-9   # intrinsics_estimate_test.cpp:39-44
    assert rel(g["kmtx"][0, :4], o["kmtx"][:4]) < 1e-9 and rel(g["hmtx"], o["hmtx"]) < 1e-9 and rel(g["poses"], o["poses"]) < 1e-8
    _, _, (xs, ys, us, vs, off) = intrinsics_estimate_scenario(seed=5, n_frames=3, rows=5, cols=7, spacing=0.04, k=(800.0, 805.0, 320.0, 240.0))
    g = simt_seed_intrinsics(seed_simt, xs, ys, us, vs, off)
    assert g["cam_success"][0] == 0 and np.allclose(g["poses"][:, :9], np.eye(3).ravel())       # too few views


def test_seed_kernels_noisy_ragged_multicamera(seed_simt):
    xs, ys, us, vs, off, cam, n_cams = noisy_multicam()
    g = simt_seed_intrinsics(seed_simt, xs, ys, us, vs, off, cam=cam, n_cams=n_cams)
    for c in range(n_cams):
        sel = np.flatnonzero(cam == c)
        pieces = [np.arange(off[k], off[k + 1]) for k in sel]
        idx = np.concatenate(pieces).astype(np.int64)
        o = O.estimate_intrinsics(xs[idx], ys[idx], us[idx], vs[idx], np.concatenate([[0], np.cumsum([len(p) for p in pieces])]))
        assert o["success"] and g["cam_success"][c] == 1 and np.array_equal(g["view_success"][sel], o["view_success"])
        assert rel(g["kmtx"][c], o["kmtx"]) < 1e-7
        ok = o["view_success"] == 1
        assert rel(g["hmtx"][sel][ok], o["hmtx"][ok]) < 1e-8
        assert np.allclose(g["sym_rms"][sel][ok], o["sym_rms"][ok], rtol=1e-8, atol=1e-6)
        assert rel(g["poses"][sel], o["poses"]) < 1e-7
    assert (g["view_success"] == 0).sum() == 2
    bounds = [200.0, 2000.0, 150.0, 2000.0, 100.0, 200.0, 50.0, 75.0, -1.0, 1.0]
    xs, ys, us, vs, off, cam, _ = noisy_multicam(n_cams=1, n_poses=60)
    g = simt_seed_intrinsics(seed_simt, xs, ys, us, vs, off, bounds=bounds)
    o = O.estimate_intrinsics(xs, ys, us, vs, off, bounds10=bounds)
    assert g["kmtx"][0, 2] == 150.0 and g["kmtx"][0, 3] == 62.5 and rel(g["kmtx"][0], o["kmtx"]) < 1e-7


def test_planar_pose_kernel(seed_simt):
    xs, ys, us, vs, off, cam, n_cams = noisy_multicam()
    kmtx = np.array([[1000.0, 1005.0, 640.0, 360.0, 0.0], [1010.0, 1015.0, 640.0, 360.0, 0.5]])
    off = np.ascontiguousarray(off, dtype=np.int64); nv = len(off) - 1
    poses = np.zeros((nv, 12)); ok = np.zeros(nv, dtype=np.int32)
    xs, ys, us, vs = (abi.as_f64(a) for a in (xs, ys, us, vs)); cam = np.ascontiguousarray(cam, dtype=np.int32)
    assert seed_simt.simt_seed_planar_poses(nv, abi.i64ptr(off), abi.i32ptr(cam), abi.dptr(xs), abi.dptr(ys), abi.dptr(us), abi.dptr(vs),
                                            abi.dptr(kmtx), abi.dptr(poses), abi.i32ptr(ok)) == 0
    for k in range(nv):
        s = slice(off[k], off[k + 1])
        T = O.estimate_planar_pose(xs[s], ys[s], us[s], vs[s], kmtx[cam[k]])
        assert rel(poses[k], np.concatenate([T[:3, :3].ravel(), T[:3, 3]])) < 1e-8, k
        assert ok[k] == (off[k + 1] - off[k] >= 4)


def test_gather_scatter_kernels_for_ragged_ransac_views(seed_simt):
    rng = np.random.default_rng(0)
    sizes = rng.choice([5, 9, 9, 12], size=11); sizes[3] = 9; sizes[7] = 9
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    x, y, u, v = rng.normal(size=(4, off[-1]))
    ids = np.flatnonzero(sizes == 9).astype(np.int64); cnt, n = len(ids), 9
    g = np.zeros((4, cnt * n)); gres = (abi.RansacResult * cnt)(); gmask = rng.integers(0, 2, cnt * n).astype(np.uint8)
    for p in range(cnt):
        gres[p].iters = 100 + p; gres[p].success = 1
    res = (abi.RansacResult * len(sizes))(); mask = np.zeros(off[-1], dtype=np.uint8)
    assert seed_simt.simt_gather_scatter_roundtrip(cnt, n, abi.i64ptr(ids), abi.i64ptr(off), abi.dptr(x), abi.dptr(y), abi.dptr(u), abi.dptr(v),
                                                   abi.dptr(g[0]), abi.dptr(g[1]), abi.dptr(g[2]), abi.dptr(g[3]), gres,
                                                   gmask.ctypes.data_as(abi.c_uint8_p), res, mask.ctypes.data_as(abi.c_uint8_p)) == 0
    for p, k in enumerate(ids):
        s = slice(off[k], off[k + 1])
        assert np.array_equal(g[0, p * n:(p + 1) * n], x[s]) and np.array_equal(g[3, p * n:(p + 1) * n], v[s])
        assert res[k].iters == 100 + p and np.array_equal(mask[s], gmask[p * n:(p + 1) * n])
    others = np.setdiff1d(np.arange(len(sizes)), ids)
    assert all(res[k].success == 0 for k in others) and not any(mask[off[k]:off[k + 1]].any() for k in others)


# ---------------------------------------------------------------------------
# AX = XB kernels (csrc/axxb_kernels.cuh) under the same shim
# ---------------------------------------------------------------------------
AXXB_SRC = os.path.join(ROOT, "tests", "host_emul", "axxb_simt.cpp")
AXXB_SO = os.path.join(ROOT, "tests", "host_emul", "_build", "libaxxb_simt.so")


@pytest.fixture(scope="module")
def axxb_simt():
    deps = [AXXB_SRC, os.path.join(ROOT, "tests", "host_emul", "simt_shim.hpp")] + [os.path.join(CSRC, f) for f in ("axxb_kernels.cuh", "k1_math.cuh")]
    if not os.path.exists(AXXB_SO) or any(os.path.getmtime(d) > os.path.getmtime(AXXB_SO) for d in deps):
        os.makedirs(os.path.dirname(AXXB_SO), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++20", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-o", AXXB_SO, AXXB_SRC], check=True)
    L = C.CDLL(AXXB_SO)
    dp = abi.c_double_p
    L.simt_axxb_eval.argtypes = [C.c_int64, dp, dp, dp, dp, C.c_double, dp, C.c_int, dp, dp, dp]
    L.simt_axxb_eval_from_poses.restype = C.c_int64
    L.simt_axxb_eval_from_poses.argtypes = [C.c_int64, dp, dp, C.c_double, C.c_int, C.c_double, C.c_double, dp, C.c_int, dp, dp, dp]
    return L


def handeye_case(n, seed=2024, noise=True):
    from calibration_b200 import geometry as G
    bg, ct, X_gt = synth.make_handeye_poses(seed=seed, n=n)
    rng = np.random.default_rng(seed + 1)
    if noise:
        ct = [synth.perturb_pose(rng, T, 0.3, 0.002) for T in ct]
    X0 = synth.perturb_pose(rng, X_gt, 3.0, 0.01)
    return bg, ct, G.pack_handeye(X0)


@pytest.mark.parametrize("huber", [1.0, 0.002, -1.0])
def test_axxb_kernels_materialised_pairs(axxb_simt, huber):
    bg, ct, x7 = handeye_case(30)
    ra, rb, ta, tb = O.build_all_pairs(bg, ct, 1.0)
    n = len(ta)
    assert n > 300                                   # more than one CTA of 256 pairs
    c_o, g_o, H_o = O.axxb_eval(O.axxb_desc(ra, rb, ta, tb, huber), x7)
    cost = C.c_double(); g = np.zeros(6); H = np.zeros((6, 6))
    ra, rb, ta, tb = (abi.as_f64(a) for a in (ra, rb, ta, tb))
    assert axxb_simt.simt_axxb_eval(n, abi.dptr(ra), abi.dptr(rb), abi.dptr(ta), abi.dptr(tb), huber, abi.dptr(abi.as_f64(x7)), 1,
                                    C.cast(C.byref(cost), abi.c_double_p), abi.dptr(g), abi.dptr(H)) == 0
    assert abs(cost.value - c_o) <= 1e-12 * c_o
    assert np.abs(g - g_o).max() <= 1e-10 * np.abs(g_o).max() and np.abs(H - H_o).max() <= 1e-10 * np.abs(H_o).max()
    c2 = C.c_double()
    assert axxb_simt.simt_axxb_eval(n, abi.dptr(ra), abi.dptr(rb), abi.dptr(ta), abi.dptr(tb), huber, abi.dptr(abi.as_f64(x7)), 0,
                                    C.cast(C.byref(c2), abi.c_double_p), None, None) == 0
    assert abs(c2.value - c_o) <= 1e-12 * c_o       # the residual-only instantiation


@pytest.mark.parametrize("n_poses,min_angle", [(20, 0.5), (45, 8.0), (70, 1.0)])
def test_axxb_kernels_pairs_on_the_fly(axxb_simt, n_poses, min_angle):
    from calibration_b200 import geometry as G
    bg, ct, x7 = handeye_case(n_poses, seed=7)
    ra, rb, ta, tb = O.build_all_pairs(bg, ct, min_angle)
    c_o, g_o, H_o = O.axxb_eval(O.axxb_desc(ra, rb, ta, tb, 1.0), x7)
    bg12 = abi.as_f64(np.stack([G.pose_to_vec12(T) for T in bg])); ct12 = abi.as_f64(np.stack([G.pose_to_vec12(T) for T in ct]))
    cost = C.c_double(); g = np.zeros(6); H = np.zeros((6, 6))
    kept = axxb_simt.simt_axxb_eval_from_poses(n_poses, abi.dptr(bg12), abi.dptr(ct12), min_angle, 1, 1e-3, 1.0, abi.dptr(abi.as_f64(x7)), 1,
                                               C.cast(C.byref(cost), abi.c_double_p), abi.dptr(g), abi.dptr(H))
    assert kept == len(ta) and 0 < kept <= n_poses * (n_poses - 1) // 2          # is_good_pair keeps the same pairs
    assert abs(cost.value - c_o) <= 1e-11 * c_o
    assert np.abs(g - g_o).max() <= 1e-9 * np.abs(g_o).max() and np.abs(H - H_o).max() <= 1e-9 * np.abs(H_o).max()


# ---------------------------------------------------------------------------
# K1 itself (csrc/k1_kernel.cuh) with the layout / setup kernels, bundle kind, fused epilogue
# ---------------------------------------------------------------------------
K1_SRC = os.path.join(ROOT, "tests", "host_emul", "k1_simt.cpp")
K1_SO = os.path.join(ROOT, "tests", "host_emul", "_build", "libk1_simt.so")


@pytest.fixture(scope="module")
def k1_simt():
    deps = [K1_SRC, os.path.join(ROOT, "tests", "host_emul", "simt_shim.hpp")] + [
        os.path.join(CSRC, f) for f in ("k1_kernel.cuh", "k1_roles.hpp", "k1_math.cuh", "refine_kernels.cuh", "refine_setup_kernels.cuh", "refine_schur_kernels.cuh",
                                       "refine_model.hpp")] + [os.path.join(ROOT, "include", "calib_b200.h")]
    if not os.path.exists(K1_SO) or any(os.path.getmtime(d) > os.path.getmtime(K1_SO) for d in deps):
        os.makedirs(os.path.dirname(K1_SO), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O1", "-std=c++20", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-I/usr/local/cuda/include", "-o", K1_SO, K1_SRC],
                       check=True)
    L = C.CDLL(K1_SO)
    dp = abi.c_double_p
    L.simt_bundle_eval.argtypes = [C.POINTER(abi.ProblemDesc), dp, dp, dp, dp, abi.c_int32_p, abi.c_int32_p]
    L.simt_tangent_count.restype = C.c_int64
    L.simt_tangent_count.argtypes = [C.POINTER(abi.ProblemDesc)]
    L.simt_views_eval.argtypes = [C.POINTER(abi.ProblemDesc), dp, dp, dp, dp, abi.c_int32_p]
    return L


def k1_eval(L, prob, x0):
    n = L.simt_tangent_count(C.byref(prob.desc))
    cost = C.c_double(); g = np.zeros(n); H = np.zeros((n, n)); roles = C.c_int32(); uni = C.c_int32()
    assert L.simt_bundle_eval(C.byref(prob.desc), abi.dptr(abi.as_f64(x0)), C.cast(C.byref(cost), abi.c_double_p), abi.dptr(g), abi.dptr(H),
                              C.byref(roles), C.byref(uni)) == 0
    return cost.value, g, H, roles.value, uni.value


K1_CASES = {   # n_poses = 32: whole tiles of full-depth blocks, the predicate-free loop; 21 / 37: padded tiles, the predicated loop
    "default_uniform": (dict(n_cams=3, n_poses=32), 2),
    "default_padded": (dict(n_cams=3, n_poses=21), 2),
    "fixed_intrinsics": (dict(n_cams=2, n_poses=32, optimize_intrinsics=False), 1),
    "skew": (dict(n_cams=2, n_poses=32, optimize_skew=True), 3),
    "skew_padded": (dict(n_cams=2, n_poses=37, optimize_skew=True), 3),
    "only_target": (dict(n_cams=2, n_poses=32, optimize_intrinsics=False, optimize_hand_eye=False), 1),
    "only_handeye": (dict(n_cams=2, n_poses=21, optimize_intrinsics=False, optimize_target_pose=False), 1),
    "only_intrinsics": (dict(n_cams=2, n_poses=32, optimize_hand_eye=False, optimize_target_pose=False), 2),
    "scheimpflug": (dict(n_cams=2, n_poses=32, model=abi.MODEL_SCHEIMPFLUG_BC5), 3),
    "scheimpflug_skew": (dict(n_cams=2, n_poses=21, model=abi.MODEL_SCHEIMPFLUG_BC5, optimize_skew=True), 3),
    "scheimpflug_fixed_intrinsics": (dict(n_cams=2, n_poses=32, model=abi.MODEL_SCHEIMPFLUG_BC5, optimize_intrinsics=False), 1),
    "no_loss": (dict(n_cams=3, n_poses=32, huber_delta=-1.0), 2),
}


@pytest.mark.parametrize("name", sorted(K1_CASES))
def test_k1_source_matches_oracle(k1_simt, name):
    kw, n_roles = K1_CASES[name]
    prob, x0, _ = synth.make_bundle(**kw)
    c_o, g_o, H_o = O.refine_eval(prob, x0)
    c, g, H, roles, uni = k1_eval(k1_simt, prob, x0)
    assert roles == n_roles                                       # 1, 2 and 3 role warps per tile are all exercised
    assert uni == kw["n_cams"] * (kw["n_poses"] // 32)             # whole tiles of full-depth blocks take the predicate-free loop
    assert len(g) == len(g_o)
    assert abs(c - c_o) <= 1e-12 * abs(c_o)
    assert np.abs(g - g_o).max() <= 1e-10 * np.abs(g_o).max()
    assert np.abs(H - H_o).max() <= 1e-10 * np.abs(H_o).max()


@pytest.mark.parametrize("n_poses", [21, 32])
def test_repack_kernel_shared_board_form_is_bitwise_the_per_observation_form(k1_simt, n_poses):
    """cal_problem_desc with board_n > 0 (one board instead of object points per observation): k_repack reads the
    board modulo the block length and must lay out exactly the same device observations, so the pass over them
    returns bit-identical cost, gradient and Hessian."""
    prob, x0, _ = synth.make_bundle(n_cams=2, n_poses=n_poses)
    pb = prob.with_shared_board()
    assert pb.desc.board_n == 88 and not pb.desc.obj_x and not pb.desc.obj_y
    c, g, H, _, _ = k1_eval(k1_simt, prob, x0)
    cb, gb, Hb, _, _ = k1_eval(k1_simt, pb, x0)
    assert c == cb and np.array_equal(g, gb) and np.array_equal(H, Hb)
    # views that do not share a board are refused by the helper (the descriptor form requires equal blocks)
    prob.x[100] += 1e-9
    with pytest.raises(ValueError):
        prob.with_shared_board()


VIEW_CASES = {
    "intrinsics": lambda: synth.make_intrinsics()[:2],                                   # C1: 20 views, 9 x 6 board, one role pair
    "intrinsics_skew": lambda: synth.make_intrinsics(optimize_skew=True)[:2],
    "intrinsics_no_loss": lambda: synth.make_intrinsics(huber_delta=-1.0)[:2],
    "intrinsics_scheimpflug": lambda: synth.make_intrinsics(model=abi.MODEL_SCHEIMPFLUG_BC5)[:2],   # three roles; the pass itself is well defined
    "extrinsics": lambda: synth.make_extrinsics(n_views=24)[:2],                          # 2 cameras: camera 0 and view 0 are the gauge
    "extrinsics_ragged": lambda: synth.make_extrinsics(n_views=21, drop_fraction=0.3)[:2],
    "extrinsics_fixed_intrinsics": lambda: synth.make_extrinsics(n_views=20, optimize_intrinsics=False)[:2],
    "extrinsics_poses_only_three_cams": lambda: synth.make_extrinsics(n_cams=3, n_views=16, optimize_intrinsics=False)[:2],
}


@pytest.mark.parametrize("name", sorted(VIEW_CASES))
def test_k1_view_store_epilogue_and_gather_match_oracle(k1_simt, name):
    """The kinds with per-view pose blocks: K1's VIEW_STORE epilogue (per-block H_vv, g_v, E_vc = Q T_c, E_vi through the
    chain rule T_b) and k_view_gather under the shim, assembled densely by the product's host model — cost, gradient and
    the full tangent-space normal matrix against the oracle's forward-mode evaluation."""
    prob, x0 = VIEW_CASES[name]()
    c_o, g_o, H_o = O.refine_eval(prob, x0)
    n = k1_simt.simt_tangent_count(C.byref(prob.desc))
    assert n == len(g_o)
    cost = C.c_double(); g = np.zeros(n); H = np.zeros((n, n)); roles = C.c_int32()
    assert k1_simt.simt_views_eval(C.byref(prob.desc), abi.dptr(abi.as_f64(x0)), C.cast(C.byref(cost), abi.c_double_p), abi.dptr(g), abi.dptr(H),
                                   C.byref(roles)) == 0
    assert abs(cost.value - c_o) <= 1e-12 * abs(c_o)
    assert np.abs(g - g_o).max() <= 1e-10 * np.abs(g_o).max()
    assert np.abs(H - H_o).max() <= 1e-10 * np.abs(H_o).max()


def test_k1_source_ragged_blocks(k1_simt):
    prob, x0, _ = synth.make_bundle(n_cams=2, n_poses=37)
    rng = np.random.default_rng(11)
    nb = prob.desc.n_blocks
    keep = rng.integers(1, 89, size=nb); keep[0] = 1
    idx = np.concatenate([np.arange(prob.block_offset[b], prob.block_offset[b] + keep[b]) for b in range(nb)])
    off = np.concatenate([[0], np.cumsum(keep)])
    p2 = abi.Problem(abi.KIND_BUNDLE, abi.MODEL_PINHOLE_BC5, 2, 0, prob.x[idx], prob.y[idx], prob.u[idx], prob.v[idx], off,
                     prob.block_cam, block_b_se3_g=prob.block_b_se3_g, optimize_intrinsics=True, huber_delta=1.0)
    c_o, g_o, H_o = O.refine_eval(p2, x0)
    c, g, H, roles, uni = k1_eval(k1_simt, p2, x0)
    assert uni == 0 and abs(c - c_o) <= 1e-12 * abs(c_o)
    assert np.abs(g - g_o).max() <= 1e-10 * np.abs(g_o).max() and np.abs(H - H_o).max() <= 1e-10 * np.abs(H_o).max()


# K2 kernels (csrc/refine_schur_kernels.cuh) under the same shim
# ---------------------------------------------------------------------------
K2_SRC = os.path.join(ROOT, "tests", "host_emul", "k2_simt.cpp")
K2_SO = os.path.join(ROOT, "tests", "host_emul", "_build", "libk2_simt.so")


@pytest.fixture(scope="module")
def k2_simt():
    deps = [K2_SRC, os.path.join(ROOT, "tests", "host_emul", "simt_shim.hpp")] + [
        os.path.join(CSRC, f) for f in ("refine_schur_kernels.cuh", "refine_kernels.cuh", "k1_math.cuh")]
    if not os.path.exists(K2_SO) or any(os.path.getmtime(d) > os.path.getmtime(K2_SO) for d in deps):
        os.makedirs(os.path.dirname(K2_SO), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O1", "-std=c++20", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-I/usr/local/cuda/include", "-o", K2_SO, K2_SRC],
                       check=True)
    L = C.CDLL(K2_SO)
    dp, ip = abi.c_double_p, abi.c_int32_p
    L.simt_k2_step.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, ip, ip, ip, ip, ip, ip, dp, dp, dp, dp, C.c_int, dp, dp, dp, C.c_double,
                               dp, dp, dp, dp, dp, ip, dp]
    L.simt_k2_plus_norms.argtypes = [C.c_int, dp, dp, dp, ip, C.c_double, dp, dp, dp]
    L.simt_k2_cov.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, ip, ip, ip, ip, ip, ip, dp, dp, dp, dp, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp]
    return L


def _k2_problem(seed, n_views, PI, radius, fixed_views=(0,), drop=0.15, n_cams=2):
    """A random normal-equation system with the block structure of the per-view kinds: per view a 6-dof pose block,
    two cameras (camera 0's pose is the gauge, camera 1's is free), PI free intrinsics per camera; every residual
    block couples ONE view with ONE camera.  Returns the per-block products K1's epilogue stores and the dense system."""
    rng = np.random.default_rng(seed)
    # shared columns: [intrinsics of camera 0 | pose (3 + 3) and intrinsics of camera 1 | ... of camera 2 | ...]
    base = [PI + (c - 1) * (PI + 6) for c in range(n_cams)]
    cq = np.array([-1] + base[1:], dtype=np.int32); ct = np.array([-1] + [b + 3 for b in base[1:]], dtype=np.int32)
    ci = np.array([0] + [b + 6 for b in base[1:]], dtype=np.int32)
    ns = n_cams * PI + 6 * (n_cams - 1)
    blocks = [(v, c) for v in range(n_views) for c in range(n_cams) if c == v % n_cams or rng.random() > drop]
    nb = len(blocks)
    view_free = np.ones(n_views, dtype=np.int32); view_free[list(fixed_views)] = 0
    Hvv = np.zeros((21, nb)); gv = np.zeros((6, nb)); Evc = np.zeros((36, nb)); Evi = np.zeros((6 * max(PI, 1), nb))
    pcol = {v: 6 * k for k, v in enumerate(np.flatnonzero(view_free))}           # dense column of a free view's block
    n_free = len(pcol)
    N = 6 * n_free + ns
    H = np.zeros((N, N)); g = np.zeros(N)
    iu = [(i, j) for i in range(6) for j in range(i, 6)]
    for b, (v, c) in enumerate(blocks):
        m = 24
        J = rng.normal(size=(m, 12 + PI)) * rng.uniform(0.5, 30.0, size=12 + PI)     # [view twist | camera pose | intrinsics], graded columns
        r = rng.normal(size=m)
        scols = np.array([-1 if cq[c] < 0 else cq[c] + j for j in range(3)] + [-1 if ct[c] < 0 else ct[c] + j for j in range(3)]
                         + [ci[c] + j for j in range(PI)])
        keep = scols >= 0
        Js = J[:, 6:][:, keep]; sc = 6 * n_free + scols[keep]
        H[np.ix_(sc, sc)] += Js.T @ Js; g[sc] += Js.T @ r
        if view_free[v]:
            Jv = J[:, :6]; pc = pcol[v] + np.arange(6)
            A = Jv.T @ Jv
            Hvv[:, b] = [A[i, j] for i, j in iu]; gv[:, b] = Jv.T @ r
            E = Jv.T @ J[:, 6:]
            Evc[:, b] = E[:, :6].reshape(-1)                       # garbage in the columns of a constant camera pose is ignored
            if PI: Evi[:, b] = E[:, 6:].reshape(-1)
            H[np.ix_(pc, pc)] += A; g[pc] += Jv.T @ r
            H[np.ix_(pc, sc)] += Jv.T @ Js; H[np.ix_(sc, pc)] += Js.T @ Jv
        else:
            Evc[:, b] = rng.normal(size=36)                        # never read for a held view
    return dict(n_views=n_views, n_cams=n_cams, PI=PI, ns=ns, blocks=blocks, view_free=view_free, cq=cq, ct=ct, ci=ci, Hvv=Hvv, gv=gv, Evc=Evc,
                Evi=Evi, H=H, g=g, pcol=pcol, n_free=n_free, radius=radius)


@pytest.mark.parametrize("case", [dict(seed=1, n_views=37, PI=9, radius=1e4), dict(seed=2, n_views=90, PI=10, radius=3.0),
                                  dict(seed=3, n_views=21, PI=0, radius=1e2), dict(seed=4, n_views=70, PI=11, radius=1e-2, fixed_views=(0, 5, 69)),
                                  dict(seed=5, n_views=45, PI=9, radius=10.0, n_cams=7),    # ns = 99: the 5 x 5-tile SYRK blocks, six warps, a one-view last step
                                  dict(seed=6, n_views=14, PI=9, radius=1.0, n_cams=12)])   # ns = 174: fifteen warps (5 x 5 block grid), a two-view last step, reduced solve on the host
def test_k2_source_matches_a_dense_solve(k2_simt, case):
    """One LM iteration of the per-view kinds through the product's K2 kernels — Jacobi scaling, clamped LM diagonal,
    per-view Cholesky, Schur complement (tiled SYRK over several CTAs), back-substitution, the per-view terms of the model
    cost change — against numpy's dense solve of the same damped, scaled normal equations."""
    P = _k2_problem(**case)
    nv, ns, nf, rad = P["n_views"], P["ns"], P["n_free"], P["radius"]
    H, g = P["H"], P["g"]
    s = 1.0 / (1.0 + np.sqrt(np.diag(H)))
    Hs = H * np.outer(s, s)
    D = np.clip(np.diag(Hs), 1e-6, 1e32)
    y = np.linalg.solve(Hs + np.diag(D / rad), g * s)
    step = -y
    sh = slice(6 * nf, 6 * nf + ns)
    ss = np.ascontiguousarray(s[sh]); Hsd = np.ascontiguousarray((Hs + np.diag(D / rad))[sh, sh]); gss = np.ascontiguousarray((g * s)[sh])
    bcam = np.array([c for _, c in P["blocks"]], dtype=np.int32); bview = np.array([v for v, _ in P["blocks"]], dtype=np.int32)
    Cm = np.zeros((ns, ns)); c = np.zeros(ns); ys = np.zeros(ns); dlt = np.zeros((nv, 6)); red = np.zeros(4); fail = np.zeros(1, dtype=np.int32); sp = np.zeros((nv, 6))
    arrs = [np.ascontiguousarray(P[k]) for k in ("Hvv", "gv", "Evc", "Evi")]
    rc = k2_simt.simt_k2_step(nv, P["n_cams"], P["PI"], len(bcam), abi.i32ptr(bcam), abi.i32ptr(bview), abi.i32ptr(P["view_free"]), abi.i32ptr(P["cq"]),
                              abi.i32ptr(P["ct"]), abi.i32ptr(P["ci"]), *[abi.dptr(a) for a in arrs], ns, abi.dptr(ss), abi.dptr(Hsd), abi.dptr(gss), rad,
                              abi.dptr(Cm), abi.dptr(c), abi.dptr(ys), abi.dptr(dlt), abi.dptr(red), abi.i32ptr(fail), abi.dptr(sp))
    assert rc == 0 and fail[0] == 0
    # Schur complement of the view blocks: E_s^T (A + D_p / radius)^-1 E_s and E_s^T (..)^-1 g_p, scaled
    pv = slice(0, 6 * nf)
    App = (Hs + np.diag(D / rad))[pv, pv]; Eps = Hs[pv, sh]
    C_ref = Eps.T @ np.linalg.solve(App, Eps) if nf else np.zeros((ns, ns))
    c_ref = Eps.T @ np.linalg.solve(App, (g * s)[pv]) if nf else np.zeros(ns)
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    if ns:
        assert rel(Cm, C_ref) < 1e-11 and rel(c, c_ref) < 1e-11
        assert rel(ys, y[sh]) < 1e-9
    for v, col in P["pcol"].items():
        assert rel(sp[v], s[col:col + 6]) < 1e-15
    d_ref = np.zeros((nv, 6))
    for v, col in P["pcol"].items():
        d_ref[v] = step[col:col + 6] * s[col:col + 6]
    assert rel(dlt, d_ref) < 1e-9
    assert np.all(dlt[P["view_free"] == 0] == 0.0)
    sg_ref = float(step[pv] @ (g * s)[pv])
    quad_ref = float(step[pv] @ Hs[pv, pv] @ step[pv] + 2.0 * step[pv] @ Hs[pv, sh] @ step[sh])
    assert abs(red[0] - sg_ref) <= 1e-9 * abs(sg_ref) and abs(red[1] - quad_ref) <= 1e-8 * max(abs(quad_ref), abs(sg_ref))


def test_k2_source_flags_a_rank_deficient_view(k2_simt):
    """a view block that is not positive definite raises the failure flag (the LM then treats the step as invalid)"""
    P = _k2_problem(seed=5, n_views=12, PI=9, radius=1e4)
    b0 = [b for b, (v, _) in enumerate(P["blocks"]) if v == 3]
    P["Hvv"][:, b0] = 0.0
    P["Hvv"][0, b0[0]] = -1.0
    nv, ns = P["n_views"], P["ns"]
    bcam = np.array([c for _, c in P["blocks"]], dtype=np.int32); bview = np.array([v for v, _ in P["blocks"]], dtype=np.int32)
    Cm = np.zeros((ns, ns)); c = np.zeros(ns); ys = np.zeros(ns); dlt = np.zeros((nv, 6)); red = np.zeros(4); fail = np.zeros(1, dtype=np.int32); sp = np.zeros((nv, 6))
    arrs = [np.ascontiguousarray(P[k]) for k in ("Hvv", "gv", "Evc", "Evi")]
    ss = np.ones(ns); Hsd = np.eye(ns); gss = np.zeros(ns)
    rc = k2_simt.simt_k2_step(nv, P["n_cams"], P["PI"], len(bcam), abi.i32ptr(bcam), abi.i32ptr(bview), abi.i32ptr(P["view_free"]), abi.i32ptr(P["cq"]),
                              abi.i32ptr(P["ct"]), abi.i32ptr(P["ci"]), *[abi.dptr(a) for a in arrs], ns, abi.dptr(ss), abi.dptr(Hsd), abi.dptr(gss), 1e4,
                              abi.dptr(Cm), abi.dptr(c), abi.dptr(ys), abi.dptr(dlt), abi.dptr(red), abi.i32ptr(fail), abi.dptr(sp))
    assert rc == 0 and fail[0] == 1


# segment layout (small problems) under the same shim
# ---------------------------------------------------------------------------
SEG_SRC = os.path.join(ROOT, "tests", "host_emul", "seg_simt.cpp")
SEG_SO = os.path.join(ROOT, "tests", "host_emul", "_build", "libseg_simt.so")


@pytest.fixture(scope="module")
def seg_simt():
    deps = [SEG_SRC, os.path.join(ROOT, "tests", "host_emul", "simt_shim.hpp"), os.path.join(ROOT, "include", "calib_b200.h")] + [
        os.path.join(CSRC, f) for f in ("k1_kernel.cuh", "k1_roles.hpp", "k1_math.cuh", "refine_kernels.cuh", "refine_setup_kernels.cuh",
                                       "refine_schur_kernels.cuh", "refine_assemble_kernels.cuh", "refine_cost_kernel.cuh", "tile_stage.cuh", "refine_model.hpp")]
    if not os.path.exists(SEG_SO) or any(os.path.getmtime(d) > os.path.getmtime(SEG_SO) for d in deps):
        os.makedirs(os.path.dirname(SEG_SO), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O1", "-std=c++20", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-I/usr/local/cuda/include", "-o", SEG_SO, SEG_SRC],
                       check=True)
    L = C.CDLL(SEG_SO)
    dp = abi.c_double_p
    L.simt_segment_eval.argtypes = [C.POINTER(abi.ProblemDesc), dp, C.c_int, C.c_int, C.c_int, dp, dp, dp, abi.c_int32_p]
    L.simt_seg_tangent_count.restype = C.c_int64
    L.simt_seg_tangent_count.argtypes = [C.POINTER(abi.ProblemDesc)]
    return L


SEG_CASES = {   # (problem, corners per segment, columns per reduction chunk)
    "c1_intrinsics": (lambda: synth.make_intrinsics()[:2], 8, 8192),                       # BASELINE configs[0]: 54 corners -> 7 segments per view
    "c1_intrinsics_skew_small_chunks": (lambda: synth.make_intrinsics(n_views=12, optimize_skew=True)[:2], 8, 48),
    "c3_extrinsics": (lambda: synth.make_extrinsics(n_views=24)[:2], 8, 8192),              # configs[2] shape: 88 corners -> 11 segments
    "extrinsics_ragged_one_segment": (lambda: synth.make_extrinsics(n_views=21, drop_fraction=0.3)[:2], 88, 64),
    "extrinsics_fixed_intrinsics": (lambda: synth.make_extrinsics(n_views=20, optimize_intrinsics=False)[:2], 8, 100),
    "bundle": (lambda: synth.make_bundle(n_cams=2, n_poses=12)[:2], 8, 8192),
}


@pytest.mark.parametrize("name", sorted(SEG_CASES))
def test_segment_layout_source_matches_oracle(seg_simt, name):
    """The layout small problems run (residual blocks cut into segments of 8-27 corners): K1's NOT_FUSED epilogue, the
    per-block Huber weight over a block's segments, k_view_part's chain rule, and the chunked fixed-order column sums,
    in the launch order of device_pass / launch_assemble — cost, gradient and normal matrix against the oracle."""
    mk, target, chunk = SEG_CASES[name]
    prob, x0 = mk()
    c_o, g_o, H_o = O.refine_eval(prob, x0)
    n = seg_simt.simt_seg_tangent_count(C.byref(prob.desc))
    assert n == len(g_o)
    cost = C.c_double(); g = np.zeros(n); H = np.zeros((n, n)); nseg = C.c_int32()
    assert seg_simt.simt_segment_eval(C.byref(prob.desc), abi.dptr(abi.as_f64(x0)), target, chunk, 0, C.cast(C.byref(cost), abi.c_double_p), abi.dptr(g),
                                      abi.dptr(H), C.byref(nseg)) == 0
    if target < 54:
        assert nseg.value > prob.desc.n_blocks          # blocks really are cut into several segments
    assert abs(cost.value - c_o) <= 1e-12 * abs(c_o)
    assert np.abs(g - g_o).max() <= 1e-10 * np.abs(g_o).max()
    assert np.abs(H - H_o).max() <= 1e-10 * np.abs(H_o).max()


COST_CASES = {
    "intrinsics_segments": (lambda: synth.make_intrinsics()[:2], 8),
    "extrinsics_ragged": (lambda: synth.make_extrinsics(n_views=21, drop_fraction=0.3)[:2], 27),
    "bundle_whole_blocks": (lambda: synth.make_bundle(n_cams=2, n_poses=24)[:2], 88),          # one segment per block: the fused layout's tiles, 11 chunks deep
    "bundle_scheimpflug_no_loss": (lambda: synth.make_bundle(n_cams=2, n_poses=10, model=abi.MODEL_SCHEIMPFLUG_BC5, huber_delta=-1.0)[:2], 20),
}


@pytest.mark.parametrize("name", sorted(COST_CASES))
def test_residual_only_pass_source_matches_oracle(seg_simt, name):
    """k_cost — the pass every candidate step of the LM runs: the per-warp two-stage TMA ring of tile_stage.cuh (as a
    memcpy plus completion counter here), projection and squared residual per corner, then the per-block Huber weight
    from the segments' sums and the fixed-order cost reduction: cost and per-block sums of squares against the oracle."""
    mk, target = COST_CASES[name]
    prob, x0 = mk()
    c_o, _, _ = O.refine_eval(prob, x0)
    ssr_o = O.block_ssr(prob, x0)
    cost = C.c_double(); ssr = np.zeros(prob.desc.n_blocks); nseg = C.c_int32()
    assert seg_simt.simt_segment_eval(C.byref(prob.desc), abi.dptr(abi.as_f64(x0)), target, 8192, 1, C.cast(C.byref(cost), abi.c_double_p), abi.dptr(ssr),
                                      None, C.byref(nseg)) == 0
    assert abs(cost.value - c_o) <= 1e-12 * abs(c_o)
    assert np.abs(ssr - ssr_o).max() <= 1e-12 * np.abs(ssr_o).max()


@pytest.mark.parametrize("case", [dict(seed=11, n_views=19, PI=9, radius=1.0), dict(seed=12, n_views=40, PI=10, radius=1.0, fixed_views=(0, 7))])
def test_covariance_kernels_source_matches_the_dense_inverse(k2_simt, case):
    """Block-structured covariance of the per-view kinds (k_cov_view_prep, k_cov_vv after launch_schur with an infinite
    radius): the view x view blocks of the inverse normal matrix, un-scaled and lifted to ambient coordinates with the
    quaternion plus-Jacobian, against numpy's inverse of the whole dense system."""
    P = _k2_problem(**case)
    nv, ns, nf = P["n_views"], P["ns"], P["n_free"]
    H = P["H"]
    s = 1.0 / (1.0 + np.sqrt(np.diag(H)))
    Hs = H * np.outer(s, s)
    sh = slice(6 * nf, 6 * nf + ns)
    ss = np.ascontiguousarray(s[sh]); Hss = np.ascontiguousarray(Hs[sh, sh])
    rng = np.random.default_rng(case["seed"])
    q = rng.normal(size=(nv, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    bcam = np.array([c for _, c in P["blocks"]], dtype=np.int32); bview = np.array([v for v, _ in P["blocks"]], dtype=np.int32)
    W = np.zeros((ns, ns)); Z = np.zeros((nv, 6, ns)); G = np.zeros((nv, 6, ns)); Ainv = np.zeros((nv, 6, 6)); cov = np.zeros((7 * nv, 7 * nv))
    arrs = [np.ascontiguousarray(P[k]) for k in ("Hvv", "gv", "Evc", "Evi")]
    rc = k2_simt.simt_k2_cov(nv, P["n_cams"], P["PI"], len(bcam), abi.i32ptr(bcam), abi.i32ptr(bview), abi.i32ptr(P["view_free"]), abi.i32ptr(P["cq"]),
                             abi.i32ptr(P["ct"]), abi.i32ptr(P["ci"]), *[abi.dptr(a) for a in arrs], ns, abi.dptr(ss), abi.dptr(Hss), abi.dptr(q),
                             abi.dptr(W), abi.dptr(Z), abi.dptr(G), abi.dptr(Ainv), abi.dptr(cov))
    assert rc == 0
    Cs = np.linalg.inv(Hs)                       # scaled covariance; tangent covariance = s C s = inv(H)
    Ct = np.linalg.inv(H)
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    assert rel(W, Cs[sh, sh]) < 1e-9
    def plus_jacobian(qq):                        # d (dq * q) / d delta at delta = 0, dq = [1, delta] (Ceres QuaternionManifold)
        w, x, y, z = qq
        return np.array([[-x, -y, -z], [w, z, -y], [-z, w, x], [y, -x, w]])
    exp = np.zeros_like(cov)
    for v, cv in P["pcol"].items():
        assert rel(Z[v], -Cs[cv:cv + 6, sh] @ np.linalg.inv(Cs[sh, sh])) < 1e-8      # C_vs = -Z_v W
        Jv = np.zeros((7, 6)); Jv[:4, :3] = plus_jacobian(q[v]); Jv[4:, 3:] = np.eye(3)
        rows = np.r_[4 * v + np.arange(4), 4 * nv + 3 * v + np.arange(3)]
        for w_, cw in P["pcol"].items():
            Jw = np.zeros((7, 6)); Jw[:4, :3] = plus_jacobian(q[w_]); Jw[4:, 3:] = np.eye(3)
            cols = np.r_[4 * w_ + np.arange(4), 4 * nv + 3 * w_ + np.arange(3)]
            exp[np.ix_(rows, cols)] = Jv @ Ct[cv:cv + 6, cw:cw + 6] @ Jw.T
    assert rel(cov, exp) < 1e-8
    held = np.flatnonzero(P["view_free"] == 0)
    for v in held:                                 # held views: zero rows and columns
        rows = np.r_[4 * v + np.arange(4), 4 * nv + 3 * v + np.arange(3)]
        assert np.all(cov[rows] == 0.0) and np.all(cov[:, rows] == 0.0)


def _ceres_quaternion_plus(q, d):
    """ceres::QuaternionManifold::Plus: x_plus = [cos|d|, sin|d| / |d| d] * x (Hamilton product, w first)"""
    n = np.linalg.norm(d)
    dq = np.array([1.0, 0, 0, 0]) if n == 0 else np.r_[np.cos(n), np.sin(n) / n * d]
    w1, x1, y1, z1 = dq; w2, x2, y2, z2 = q
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2])


def test_view_plus_and_norms_kernels_match_the_ceres_manifold(k2_simt):
    """k_view_plus / k_view_norms / k_reduce_views: the per-view part of the candidate point, of the step norm, of |x| and
    of the projected-gradient max norm, against the QuaternionManifold formulas (un-normalised quaternions, as the LM keeps them)."""
    rng = np.random.default_rng(21)
    nv, t = 1100, 0.7                                   # more views than the 1024 threads of the reduction
    q = rng.normal(size=(nv, 4)) * rng.uniform(0.8, 1.2, size=(nv, 1)); tr = rng.normal(size=(nv, 3))
    x = np.r_[q.reshape(-1), tr.reshape(-1)]
    delta = rng.normal(size=(nv, 6)) * 0.05; delta[3] = 0.0
    gp = rng.normal(size=(nv, 6)) * 1e-3
    vfree = np.ones(nv, dtype=np.int32); vfree[[0, 17]] = 0
    xc = np.zeros(7 * nv); rp = np.zeros(4); rn = np.zeros(4)
    assert k2_simt.simt_k2_plus_norms(nv, abi.dptr(x), abi.dptr(np.ascontiguousarray(delta)), abi.dptr(np.ascontiguousarray(gp)), abi.i32ptr(vfree), t,
                                      abi.dptr(xc), abi.dptr(rp), abi.dptr(rn)) == 0
    q_ref = np.array([_ceres_quaternion_plus(q[v], t * delta[v, :3]) for v in range(nv)])
    t_ref = tr + t * delta[:, 3:]
    assert np.abs(xc[:4 * nv].reshape(nv, 4) - q_ref).max() < 1e-15 and np.abs(xc[4 * nv:].reshape(nv, 3) - t_ref).max() < 1e-15
    dx2 = ((q_ref - q) ** 2).sum() + ((t_ref - tr) ** 2).sum()
    assert abs(rp[2] - dx2) <= 1e-12 * dx2
    assert abs(rn[0] - (x ** 2).sum()) <= 1e-12 * (x ** 2).sum()
    gm = 0.0
    for v in np.flatnonzero(vfree):
        gm = max(gm, np.abs(q[v] - _ceres_quaternion_plus(q[v], -gp[v, :3])).max(), np.abs(gp[v, 3:]).max())
    assert abs(rn[3] - gm) <= 1e-12 * gm


@pytest.mark.parametrize("n", [1, 7, 15, 16, 17, 31, 32, 33, 114, 145, 160])
def test_reduced_solve_source_matches_a_dense_solve(k2_simt, n):
    """k_reduced_solve (blocked Cholesky of the shared block in one CTA, right-hand side as an extra row, blocked back
    substitution) against a dense host solve, at sizes around the panel width (16), at the c5-size extrinsics problem
    (114) and at the widest block the kernel takes (160).  Replaces the host factorisation of
    ceres-style normal equations the reference leaves to Ceres (ceresutils.h:27-43)."""
    rng = np.random.default_rng(100 + n)
    M = rng.standard_normal((n, n + 3))
    S = M @ M.T + 0.5 * np.eye(n)
    Cm = 0.1 * (lambda Q: Q @ Q.T)(rng.standard_normal((n, 2)))      # a positive semi-definite Schur complement to subtract
    Sm = np.ascontiguousarray(S + Cm)
    g, c = rng.standard_normal(n), rng.standard_normal(n)
    y, err = np.zeros(n), np.zeros(1)
    rc = k2_simt.simt_k2_reduced_solve(n, abi.dptr(Sm), abi.dptr(np.ascontiguousarray(Cm)), abi.dptr(g), abi.dptr(c), abi.dptr(y), abi.dptr(err))
    assert rc == 0, rc
    assert err[0] <= 1e-11, err[0]
    assert np.abs(S @ y - (g - c)).max() <= 1e-9 * (np.abs(g - c).max() + np.abs(S).max() * np.abs(y).max())


def test_reduced_solve_source_flags_an_indefinite_block(k2_simt):
    n = 40
    rng = np.random.default_rng(7)
    M = rng.standard_normal((n, n))
    S = M @ M.T + np.eye(n)
    S[25, 25] = -1.0                                                  # pivot 25 (second panel) turns negative
    y, err = np.zeros(n), np.zeros(1)
    rc = k2_simt.simt_k2_reduced_solve(n, abi.dptr(np.ascontiguousarray(S)), abi.dptr(np.zeros((n, n))), abi.dptr(rng.standard_normal(n)), abi.dptr(np.zeros(n)),
                                       abi.dptr(y), abi.dptr(err))
    assert rc == 1, rc
