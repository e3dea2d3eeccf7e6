"""The product's refinement path — ALL of it: cal_refine_create's layout construction, device_pass, the Levenberg–Marquardt
loop with its Schur steps, the covariance — executed on the CPU from the product's own sources.

calibration_b200/csrc/refine_host.cu is compiled UNMODIFIED by g++ against a host-only stand-in of the CUDA runtime
(tests/host_emul/fake_cuda/cuda_runtime.h: device memory is host memory, streams are tokens) and linked with
tests/host_emul/refine_product_simt.cpp, which implements every launch_* of refine_kernels.cu / k1_fused.cu with
simt::launch on the same kernel sources under the lock-step SIMT shim.  The result, tests/host_emul/_build/
libcalib_b200_simt.so, exports the refinement part of the C ABI; the ctypes layer is pointed at it for the duration of a
test.  TEST INFRASTRUCTURE: nothing in the package can load it, and the product still has no CPU path — this is how the
host logic that the GPU suite exercises on the device is checked when there is none: the solves below must reproduce the
oracle's iteration counts, parameters, final cost and covariance."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from calibration_b200 import abi, capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "host_emul")
CSRC = os.path.join(ROOT, "calibration_b200", "csrc")
OUT = os.path.join(EMUL, "_build")
SO = os.path.join(OUT, "libcalib_b200_simt.so")


def _build():
    deps = [os.path.join(EMUL, f) for f in ("refine_product_simt.cpp", "simt_shim.hpp", os.path.join("fake_cuda", "cuda_runtime.h"))] + [
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp", ".h")) or f == "refine_host.cu"] + [
        os.path.join(ROOT, "include", "calib_b200.h")]
    if os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    os.makedirs(OUT, exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    fake = os.path.join(EMUL, "fake_cuda")
    o1, o2 = os.path.join(OUT, "refine_product_simt.o"), os.path.join(OUT, "refine_host_cpu.o")
    p1 = subprocess.Popen([cxx, "-O1", "-std=c++20", "-fPIC", "-pthread", "-Wno-unknown-pragmas", "-I", fake, "-c",
                           os.path.join(EMUL, "refine_product_simt.cpp"), "-o", o1])
    p2 = subprocess.Popen([cxx, "-O2", "-std=c++17", "-fPIC", "-x", "c++", "-I", fake, "-c", os.path.join(CSRC, "refine_host.cu"), "-o", o2])
    assert p1.wait() == 0 and p2.wait() == 0
    subprocess.run([cxx, "-shared", "-pthread", "-o", SO, o1, o2], check=True)
    return SO


@pytest.fixture()
def product_on_cpu(monkeypatch):
    """points calibration_b200.capi at the CPU build of the product's refinement sources for one test"""
    so = _build()
    monkeypatch.setattr(capi._build, "LIB", so)
    monkeypatch.setattr(capi, "_lib", None)
    assert capi.device_count() == 1          # the stand-in runtime's one "device"
    yield capi
    # monkeypatch restores capi._lib / LIB: the next capi.lib() is the real library again


CASES = {
    "bundle": lambda: synth.make_bundle(n_cams=2, n_poses=12),
    "intrinsics_c1_shape": lambda: synth.make_intrinsics(n_views=8),
    "extrinsics_stereo": lambda: synth.make_extrinsics(n_views=8),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_product_lm_solve_reproduces_the_oracle(product_on_cpu, monkeypatch, name):
    monkeypatch.setenv("CALIB_B200_FUSED", "1")   # the layout of the benchmark; a few hundred OS threads per pass instead of tens of thousands
    prob, x0, _ = CASES[name]()
    h = product_on_cpu.RefineHandle(prob)
    c, g, H = h.eval(x0)
    c_o, g_o, H_o = O.refine_eval(prob, x0)
    assert abs(c - c_o) <= 1e-12 * abs(c_o) and np.abs(g - g_o).max() <= 1e-10 * np.abs(g_o).max() and np.abs(H - H_o).max() <= 1e-10 * np.abs(H_o).max()
    x, r, cov = h.solve(x0)
    rms_blocks, rms = h.view_errors(x)
    h.close()
    x_o, r_o, cov_o = O.refine_solve(prob, x0)
    assert r.success and r_o.success
    assert (r.iterations, r.num_jac_evals, r.num_cost_evals) == (r_o.iterations, r_o.num_jac_evals, r_o.num_cost_evals)
    assert np.abs(x - x_o).max() <= 1e-8 * np.abs(x_o).max()
    assert abs(r.final_cost - r_o.final_cost) <= 1e-10 * r_o.final_cost and abs(r.initial_cost - r_o.initial_cost) <= 1e-12 * r_o.initial_cost
    assert r.report.decode() .split("Termination")[1] == r_o.report.decode().split("Termination")[1]
    assert r.covariance_ok and r_o.covariance_ok and np.abs(cov - cov_o).max() <= 1e-6 * np.abs(cov_o).max()
    ssr = O.block_ssr(prob, x)
    n_pts = np.diff(prob.block_offset)
    assert np.abs(rms_blocks - np.sqrt(ssr / (2 * n_pts))).max() <= 1e-10      # view_errors: sqrt(sum r^2 / (2 points)) per block
    assert abs(rms - np.sqrt(ssr.sum() / (2 * n_pts.sum()))) <= 1e-10


def test_product_segment_layout_pass_and_validation(product_on_cpu):
    """the default layout of small problems (no CALIB_B200_FUSED), one fused pass and one residual-only pass; and
    cal_refine_create's validation in front of it"""
    prob, x0, _ = synth.make_intrinsics(n_views=4)
    h = product_on_cpu.RefineHandle(prob)
    info = h.layout_info()
    assert info["n_segments"] > prob.desc.n_blocks          # blocks are cut into segments
    c, g, H = h.eval(x0)
    c2, ssr = h.cost(x0, want_block_ssr=True)
    h.close()
    c_o, g_o, H_o = O.refine_eval(prob, x0)
    assert abs(c - c_o) <= 1e-12 * abs(c_o) and abs(c2 - c_o) <= 1e-12 * abs(c_o)
    assert np.abs(g - g_o).max() <= 1e-10 * np.abs(g_o).max() and np.abs(H - H_o).max() <= 1e-10 * np.abs(H_o).max()
    assert np.abs(ssr - O.block_ssr(prob, x0)).max() <= 1e-12 * ssr.max()
    prob3, _, _ = synth.make_intrinsics(n_views=3)
    with pytest.raises(ValueError, match="at least 4"):
        product_on_cpu.RefineHandle(prob3)


def test_the_real_library_is_back_afterwards():
    if capi.device_count() == 0:
        prob, _, _ = synth.make_bundle(n_cams=1, n_poses=8)
        with pytest.raises(capi.CalibCudaError):
            capi.RefineHandle(prob)


@pytest.mark.parametrize("fused", ["0", "1"])
def test_shared_board_form_through_the_product_host_code(product_on_cpu, monkeypatch, fused):
    """cal_problem_desc.board_n > 0 through cal_refine_create itself — the staging-buffer arithmetic and copies of the
    shared-board form, then k_repack — must give bit-identical passes to the per-observation form in both layouts."""
    monkeypatch.setenv("CALIB_B200_FUSED", fused)
    prob, x0, _ = synth.make_bundle(n_cams=1, n_poses=4, optimize_intrinsics=fused == "1")   # small: the segment layout's 2-D reduction grids are slow under the shim
    out = []
    for p in (prob, prob.with_shared_board()):
        h = product_on_cpu.RefineHandle(p)
        out.append(h.eval(x0) + (h.cost(x0),))
        h.close()
    (c, g, H, c1), (cb, gb, Hb, c1b) = out
    assert c == cb and c1 == c1b and np.array_equal(g, gb) and np.array_equal(H, Hb)
    pb = prob.with_shared_board()
    pb.desc.board_n = 80
    with pytest.raises(ValueError, match="exactly board_n"):
        product_on_cpu.RefineHandle(pb)


def test_every_parameter_block_frozen(product_on_cpu, monkeypatch):
    """optimize_bundle with all three optimize_* flags false — Ceres finds no free parameter block and
    returns CONVERGENCE after 0 iterations with final_cost = the fixed cost; the product's LM must do the same."""
    monkeypatch.setenv("CALIB_B200_FUSED", "1")
    prob, x0, _ = synth.make_bundle(n_cams=1, n_poses=8, optimize_intrinsics=False, optimize_hand_eye=False, optimize_target_pose=False)
    h = product_on_cpu.RefineHandle(prob)
    assert h.n_tan == 0
    x, r, cov = h.solve(x0)
    h.close()
    x_o, r_o, _ = O.refine_solve(prob, x0)
    assert r.success and r.iterations == 0 == r_o.iterations
    assert abs(r.final_cost - r_o.final_cost) <= 1e-12 * r_o.final_cost and r.final_cost == r.initial_cost
    assert np.array_equal(x, x0)


@pytest.mark.parametrize("name", ["intrinsics_c1_shape", "extrinsics_stereo"])
def test_speculative_pass_of_the_per_view_kinds_changes_nothing(product_on_cpu, monkeypatch, name):
    """An LM candidate of the per-view kinds is evaluated with the full fused pass into the handle's second set of
    per-block / per-view buffers (one pass per accepted step instead of a residual-only pass plus a Jacobian pass).  The
    solve must take the same steps as with CALIB_B200_NO_SPECULATION: same counters (the reference's meaning, ceresutils.h
    :27-43 — one cost evaluation per candidate, one Jacobian evaluation per accepted point), same parameters and
    covariance, fewer launches."""
    monkeypatch.setenv("CALIB_B200_FUSED", "1")
    prob, x0, _ = CASES[name]()
    out = []
    for nospec in (False, True):
        if nospec: monkeypatch.setenv("CALIB_B200_NO_SPECULATION", "1")
        h = product_on_cpu.RefineHandle(prob)
        x, r, cov = h.solve(x0)
        out.append((x, (r.iterations, r.num_jac_evals, r.num_cost_evals), r.final_cost, cov, h.launch_count()))
        h.close()
    (xa, ca, fa, cova, la), (xb, cb, fb, covb, lb) = out
    assert ca == cb and la < lb
    assert np.abs(xa - xb).max() <= 1e-12 * np.abs(xb).max() and abs(fa - fb) <= 1e-12 * fb
    assert np.abs(cova - covb).max() <= 1e-9 * np.abs(covb).max()
