"""The product's per-observation math (k1_math.cuh) and host assembly (refine_model.hpp), compiled with g++
and run on the CPU in the dataflow of K1's fused epilogue, against the oracle: pinhole / Scheimpflug,
every intrinsics mode, fixed / free block combinations, Huber on and off, ragged views.  No GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from calibration_b200 import abi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_emul", "emul.cpp")
SO = os.path.join(ROOT, "tests", "host_emul", "_build", "libemul.so")


@pytest.fixture(scope="module")
def emul():
    deps = [SRC] + [os.path.join(ROOT, "calibration_b200", "csrc", f) for f in ("k1_math.cuh", "refine_model.hpp", "refine_kernels.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-I/usr/local/cuda/include", "-o", SO, SRC], check=True)
    L = C.CDLL(SO)
    L.emul_bundle_eval.argtypes = [C.POINTER(abi.ProblemDesc), abi.c_double_p, abi.c_double_p, abi.c_double_p, abi.c_double_p]
    L.emul_tangent_count.argtypes = [C.POINTER(abi.ProblemDesc)]
    L.emul_views_eval.argtypes = L.emul_bundle_eval.argtypes
    return L


CASES = {
    "default": dict(n_cams=3, n_poses=30),
    "fixed_intrinsics": dict(n_cams=2, n_poses=30, optimize_intrinsics=False),
    "skew": dict(n_cams=2, n_poses=30, optimize_skew=True),
    "only_target": dict(n_cams=2, n_poses=30, optimize_intrinsics=False, optimize_hand_eye=False),
    "only_handeye": dict(n_cams=2, n_poses=30, optimize_intrinsics=False, optimize_target_pose=False),
    "only_intrinsics": dict(n_cams=2, n_poses=30, optimize_hand_eye=False, optimize_target_pose=False),
    "scheimpflug": dict(n_cams=2, n_poses=30, model=abi.MODEL_SCHEIMPFLUG_BC5),
    "scheimpflug_skew": dict(n_cams=2, n_poses=24, model=abi.MODEL_SCHEIMPFLUG_BC5, optimize_skew=True),
    "scheimpflug_fixed_intrinsics": dict(n_cams=2, n_poses=24, model=abi.MODEL_SCHEIMPFLUG_BC5, optimize_intrinsics=False),
    "no_loss": dict(n_cams=3, n_poses=24, huber_delta=-1.0),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_emulated_pass_matches_oracle(emul, name):
    prob, x0, _ = synth.make_bundle(**CASES[name])
    n = emul.emul_tangent_count(C.byref(prob.desc))
    c_o, g_o, H_o = O.refine_eval(prob, x0)
    assert n == len(g_o)
    cost = C.c_double(); g = np.zeros(n); H = np.zeros((n, n))
    assert emul.emul_bundle_eval(C.byref(prob.desc), abi.dptr(abi.as_f64(x0)), C.cast(C.byref(cost), abi.c_double_p), abi.dptr(g), abi.dptr(H)) == 0
    assert abs(cost.value - c_o) <= 1e-12 * abs(c_o)
    assert np.abs(g - g_o).max() <= 1e-10 * np.abs(g_o).max()
    assert np.abs(H - H_o).max() <= 1e-10 * np.abs(H_o).max()


def test_emulated_pass_ragged_views(emul):
    prob, x0, _ = synth.make_bundle(n_cams=2, n_poses=37)
    rng = np.random.default_rng(11)
    nb = prob.desc.n_blocks
    keep = rng.integers(1, 89, size=nb); keep[0] = 1
    idx = np.concatenate([np.arange(prob.block_offset[b], prob.block_offset[b] + keep[b]) for b in range(nb)])
    off = np.concatenate([[0], np.cumsum(keep)])
    p2 = abi.Problem(abi.KIND_BUNDLE, abi.MODEL_PINHOLE_BC5, 2, 0, prob.x[idx], prob.y[idx], prob.u[idx], prob.v[idx], off,
                     prob.block_cam, block_b_se3_g=prob.block_b_se3_g, optimize_intrinsics=True, huber_delta=1.0)
    n = emul.emul_tangent_count(C.byref(p2.desc))
    c_o, g_o, H_o = O.refine_eval(p2, x0)
    cost = C.c_double(); g = np.zeros(n); H = np.zeros((n, n))
    assert emul.emul_bundle_eval(C.byref(p2.desc), abi.dptr(abi.as_f64(x0)), C.cast(C.byref(cost), abi.c_double_p), abi.dptr(g), abi.dptr(H)) == 0
    assert abs(cost.value - c_o) <= 1e-12 * abs(c_o) and np.abs(H - H_o).max() <= 1e-10 * np.abs(H_o).max()


VIEW_CASES = {
    "intrinsics": lambda: synth.make_intrinsics(),
    "intrinsics_skew": lambda: synth.make_intrinsics(optimize_skew=True),
    "intrinsics_no_loss": lambda: synth.make_intrinsics(huber_delta=-1.0),
    "intrinsics_scheimpflug_skew": lambda: synth.make_intrinsics(model=abi.MODEL_SCHEIMPFLUG_BC5, optimize_skew=True),
    "extrinsics": lambda: synth.make_extrinsics(n_views=30),
    "extrinsics_missing_views": lambda: synth.make_extrinsics(n_cams=3, n_views=40, drop_fraction=0.3),
    "extrinsics_fixed_intrinsics": lambda: synth.make_extrinsics(n_cams=3, n_views=25, optimize_intrinsics=False),
    "extrinsics_fixed_extrinsics": lambda: synth.make_extrinsics(n_views=25, optimize_extrinsics=False),
    "extrinsics_skew": lambda: synth.make_extrinsics(n_views=25, optimize_skew=True),
}


@pytest.mark.parametrize("name", sorted(VIEW_CASES))
def test_emulated_per_view_kinds_match_oracle(emul, name):
    """compose_intrinsics / compose_extrinsics, the view- and camera-type chain-rule transforms and the dense
    assembly (per-view 6x6 blocks + couplings) of the kinds the Schur kernels serve."""
    prob, x0, _ = VIEW_CASES[name]()
    n = emul.emul_tangent_count(C.byref(prob.desc))
    c_o, g_o, H_o = O.refine_eval(prob, x0)
    assert n == len(g_o)
    cost = C.c_double(); g = np.zeros(n); H = np.zeros((n, n))
    assert emul.emul_views_eval(C.byref(prob.desc), abi.dptr(abi.as_f64(x0)), C.cast(C.byref(cost), abi.c_double_p), abi.dptr(g), abi.dptr(H)) == 0
    assert abs(cost.value - c_o) <= 1e-12 * abs(c_o)
    assert np.abs(g - g_o).max() <= 1e-10 * np.abs(g_o).max()
    assert np.abs(H - H_o).max() <= 1e-10 * np.abs(H_o).max()


def test_k1_role_tables_invariants():
    """K1's compile-time split of the local system over roles and the per-tile value map (k1_roles.hpp): every entry
    owned once, dense slots, the epilogue's role-locality, and a value map that is a bijection onto the camera row."""
    src = os.path.join(ROOT, "tests", "host_emul", "roles_check.cpp")
    so = os.path.join(ROOT, "tests", "host_emul", "_build", "libroles.so")
    deps = [src] + [os.path.join(ROOT, "calibration_b200", "csrc", f) for f in ("k1_roles.hpp", "k1_math.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, src], check=True)
    L = C.CDLL(so)
    assert L.roles_check() == 0
    assert [L.roles_of(0, i) for i in range(3)] == [1, 2, 3] and [L.roles_of(1, i) for i in range(3)] == [1, 3, 3]


# ---- plane RANSAC: the product's per-hypothesis math (plane_math.cuh) in the kernel's batch order ----
PLANE_SRC = os.path.join(ROOT, "tests", "host_emul", "plane_emul.cpp")
PLANE_SO = os.path.join(ROOT, "tests", "host_emul", "_build", "libplane_emul.so")


@pytest.fixture(scope="module")
def plane_emul():
    deps = [PLANE_SRC] + [os.path.join(ROOT, "calibration_b200", "csrc", f) for f in ("plane_math.cuh", "ransac_iters.hpp")]
    if not os.path.exists(PLANE_SO) or any(os.path.getmtime(d) > os.path.getmtime(PLANE_SO) for d in deps):
        os.makedirs(os.path.dirname(PLANE_SO), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", PLANE_SO, PLANE_SRC], check=True)
    L = C.CDLL(PLANE_SO)
    L.emul_ransac_plane.argtypes = [C.c_int32, abi.c_double_p, abi.c_double_p, abi.c_double_p, C.POINTER(abi.RansacOptions),
                                    abi.c_int32_p, C.POINTER(abi.PlaneResult), abi.c_uint8_p]
    return L


def run_plane_emul(L, x, y, z, opts):
    x, y, z = (abi.as_f64(a) for a in (x, y, z))
    idx = O.sample_stream_k(opts.seed, len(x), 3, max(opts.max_iters, 1))
    res = abi.PlaneResult(); mask = np.zeros(len(x), dtype=np.uint8)
    L.emul_ransac_plane(len(x), abi.dptr(x), abi.dptr(y), abi.dptr(z), C.byref(opts), idx.ctypes.data_as(abi.c_int32_p),
                        C.byref(res), mask.ctypes.data_as(abi.c_uint8_p))
    return res, mask


def check_plane(L, x, y, z, opts, margin=1e-10):
    ro, mo = O.ransac_plane(x, y, z, opts)
    re_, me = run_plane_emul(L, x, y, z, opts)
    if ro.min_margin <= margin:
        return False
    assert bool(re_.success) == bool(ro.success)
    assert np.array_equal(me, mo) and re_.n_inliers == ro.n_inliers
    assert re_.iters == ro.iters and re_.iters_run == ro.iters_run
    if ro.success:
        assert np.abs(np.array(re_.plane) - np.array(ro.plane)).max() <= 1e-9
        assert abs(re_.inlier_rms - ro.inlier_rms) <= 1e-9 * ro.inlier_rms + 1e-13
    else:
        assert list(re_.plane) == [0.0] * 4
    return True


def test_plane_math_replays_oracle_loop(plane_emul):
    x, y, z, _ = synth.synth_plane_ransac(seed=12, n_problems=48, n=200)
    n_ok = 0
    for p in range(48):
        for kw in (dict(), dict(refit_on_inliers=0), dict(confidence=0.0, max_iters=70), dict(min_inliers=190, max_iters=90)):
            n_ok += check_plane(plane_emul, x[p], y[p], z[p], abi.RansacOptions.default(seed=1234567 + p, thresh=0.006, **kw))
    assert n_ok >= 0.97 * 48 * 4


def test_plane_math_reference_scenario_and_edges(plane_emul):
    gt, xyz = O.plane_testdata()     # planefit_test.cpp:22-75 (noise-free plane: the scatter matrix is singular)
    assert check_plane(plane_emul, *xyz.T, abi.RansacOptions.default(max_iters=2000, thresh=0.01, min_inliers=80, confidence=0.999))
    t = np.linspace(0, 1, 40)
    assert check_plane(plane_emul, t, 2 * t, -t, abi.RansacOptions.default(min_inliers=3, max_iters=40))      # all samples degenerate
    assert check_plane(plane_emul, [0.0, 1.0], [0.0, 0.0], [0.0, 0.0], abi.RansacOptions.default(min_inliers=1))
    q = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0.3, 0.3, 0.5]], float)
    assert check_plane(plane_emul, *q.T, abi.RansacOptions.default(min_inliers=3, thresh=1e-6, max_iters=30))
    for n in (3, 31, 33, 65):
        x, y, z, _ = synth.synth_plane_ransac(seed=n, n_problems=1, n=n, outlier_fraction=0.2)
        check_plane(plane_emul, x[0], y[0], z[0], abi.RansacOptions.default(thresh=0.006, min_inliers=min(12, n), max_iters=150))
