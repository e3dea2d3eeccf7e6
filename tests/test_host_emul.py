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
