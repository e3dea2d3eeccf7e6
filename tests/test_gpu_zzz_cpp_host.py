"""The reference's unit tests for the refinement path, re-expressed in C++ on the drop-in adapter
(tests/cpp/reference_tests.cpp over include/calib_b200_adapter.hpp), linked against libcalib_b200.so and run on the
GPU: optimize_intrinsics / optimize_extrinsics / optimize_bundle (pinhole and Scheimpflug) / optimize_handeye /
estimate_homography (DLT and RANSAC) / fit_plane_ransac / estimate_intrinsics / estimate_planar_pose with the
reference's seeds, scenes and tolerances.  No Python and no torch in that process.  (Named zzz so it runs last.)"""
import re

import pytest

import cpp_host_build as B

pytestmark = pytest.mark.gpu


def test_reference_unit_tests_pass_on_the_adapter_over_the_cuda_library():
    exe, env = B.build_real()
    out = B.run(exe, env, timeout=900)
    m = re.search(r"(\d+) tests ran, (\d+) failed", out.stdout)
    assert m, out.stdout[-4000:] + out.stderr[-2000:]
    assert out.returncode == 0 and int(m.group(2)) == 0 and int(m.group(1)) >= 29, out.stdout[-6000:]
