"""The C++ host layer (include/calib_b200_adapter.hpp, stand-alone mode) on the CPU: it compiles as C++20 with
-Wall -Wextra -Werror, the reference's unit tests re-expressed on it (tests/cpp/reference_tests.cpp) pass when the
C ABI underneath is answered by the CPU stand-in (packing, block order, unpacking, error mapping), and against the
REAL library its argument validation maps to the reference's exception types while every computing call fails
loudly without a device.  The parity run of the same binary is tests/test_gpu_zzz_cpp_host.py."""
import re

import cpp_host_build as B

VALIDATION = ["InsufficientViewsThrow", "InputValidation", "MismatchedPoseVectorsThrow", "InsufficientPoints", "TooFewPointsFail"]


def _summary(out):
    m = re.search(r"(\d+) tests ran, (\d+) failed", out.stdout)
    assert m, out.stdout + out.stderr
    return int(m.group(1)), int(m.group(2))


import pytest


@pytest.mark.parametrize("per_observation", [False, True])
def test_reference_unit_tests_pass_on_the_adapter_over_the_cpu_standin(per_observation):
    """both descriptor forms the adapter can emit: one shared board (default) and object_xy per observation"""
    exe, env = B.build_standin()
    if per_observation:
        env = dict(env, CALIB_B200_PER_OBSERVATION="1")
    out = B.run(exe, env)
    ran, failed = _summary(out)
    assert out.returncode == 0 and failed == 0, out.stdout[-4000:]
    assert ran >= 30


def test_argument_validation_through_the_real_library_without_a_device():
    exe, env = B.build_real()
    out = B.run(exe, env, *VALIDATION)
    ran, failed = _summary(out)
    assert out.returncode == 0 and failed == 0 and ran == len(VALIDATION), out.stdout[-4000:]


def test_computing_calls_fail_loudly_without_a_device():
    from calibration_b200 import capi
    if capi.device_count() > 0:
        import pytest
        pytest.skip("a CUDA device is present")
    exe, env = B.build_real()
    for name in ("OptimizeBundle.SingleCameraHandEye", "OptimizeIntrinsics.RecoversSkew", "CeresAXXBRefine.ImprovesOverInitializer",
                 "HomographyTest.RansacRecoversHomographyWithOutliers", "PlaneFit.RansacRejectsOutliers",
                 "EstimateIntrinsics.RecoversCameraMatrix", "PlanarPoseTest.DLTEstimation"):
        out = B.run(exe, env, name)
        assert out.returncode != 0, name
        assert "no CUDA device: calib_b200 has no CPU fallback" in out.stdout, out.stdout[-2000:]


def test_cpp_example_of_the_adapter_builds_and_runs_over_the_standin():
    """examples/cpp_adapter_example.cpp: estimate_intrinsics -> estimate_planar_pose -> optimize_intrinsics on noisy
    data, the way a user of the reference writes it"""
    exe, env = B.build_example(real=False)
    out = B.run(exe, env)
    assert out.returncode == 0 and "CONVERGENCE" in out.stdout, out.stdout + out.stderr
    exe, env = B.build_example(real=True)     # links against the product; without a device it reports the error and exits 2
    from calibration_b200 import capi
    if capi.device_count() == 0:
        out = B.run(exe, env)
        assert out.returncode == 2 and "no CUDA device" in out.stdout


def test_reference_unit_tests_on_the_adapter_over_the_products_own_host_code_and_kernels():
    """A subset of the C++ tests with cal_refine_* answered by the product itself on the CPU — refine_host.cu compiled
    against the host-only CUDA runtime stand-in, kernels under the SIMT shim (tests/test_product_on_cpu.py) — in the
    adapter's default shared-board form: one test per refinement kind and model, plus the validation paths."""
    exe, env = B.build_product_on_cpu()
    names = ["RecoversIntrinsicsNoSkew", "SingleCameraHandEye", "RecoverAllParameters", "HandeyeWithFixedIntrinsics", "InputValidation",
             "InsufficientViewsThrow", "MismatchedPoseVectorsThrow"]
    out = B.run(exe, dict(env, CALIB_B200_FUSED="1"), *names, timeout=900)
    ran, failed = _summary(out)
    assert out.returncode == 0 and failed == 0 and ran == len(names), out.stdout[-4000:]
