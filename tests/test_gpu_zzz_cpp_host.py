"""The reference's unit tests for the refinement path, re-expressed in C++ on the drop-in adapter
(tests/cpp/reference_tests.cpp over include/calib_b200_adapter.hpp), linked against libcalib_b200.so and run on the
GPU: optimize_intrinsics / optimize_extrinsics / optimize_bundle (pinhole and Scheimpflug) / optimize_handeye /
estimate_homography (DLT and RANSAC) / fit_plane_ransac / estimate_intrinsics / estimate_planar_pose with the
reference's seeds, scenes and tolerances.  No Python and no torch in that process.  (Named zzz so it runs last.)"""
import re

import pytest

import cpp_host_build as B

pytestmark = pytest.mark.gpu


def _run(env):
    exe, base = B.build_real()
    out = B.run(exe, dict(base, **env), timeout=900)
    m = re.search(r"(\d+) tests ran, (\d+) failed", out.stdout)
    assert m, out.stdout[-4000:] + out.stderr[-2000:]
    assert out.returncode == 0 and int(m.group(2)) == 0 and int(m.group(1)) >= 30, out.stdout[-6000:]


def test_reference_unit_tests_pass_on_the_adapter_over_the_cuda_library():
    """per-observation form of cal_problem_desc (object_xy uploaded for every observation)"""
    _run({"CALIB_B200_PER_OBSERVATION": "1"})


def test_reference_unit_tests_pass_with_the_shared_board_form():
    """the adapter's default: one board while the views share it (cal_problem_desc.board_n > 0)"""
    _run({})


def test_cpp_example_of_the_adapter_runs():
    exe, env = B.build_example(real=True)
    out = B.run(exe, env, timeout=300)
    assert out.returncode == 0 and "CONVERGENCE" in out.stdout, out.stdout + out.stderr


def test_optimize_bundle_from_the_reference_input_type_at_scale():
    """examples/cpp_bundle_e2e.cpp: calib::optimize_bundle(std::vector<BundleObservation>) with 8 cameras x 3 000 poses x 88 corners
    (2.1 M observations: the staged, threaded packing into page-locked memory), twice; the program checks convergence and fx itself."""
    import json
    exe, env = B.build_bundle_e2e()
    out = B.run(exe, env, "3000", "2", timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-1000:]
    rec = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert rec["observations"] == 8 * 3000 * 88 and all(r.get("converged") and r["covariance_rows"] == 143 for r in rec["runs"])
    assert rec["runs"][0]["final_cost"] == rec["runs"][1]["final_cost"]   # run-to-run identical
