"""The C-ABI shared library loads and exports every symbol include/calib_b200.h declares
(no compute calls: this runs without a GPU), and fails loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

from calibration_b200 import abi, build, capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "calib_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cal_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_declared_symbols():
    path = build.build()
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layouts_match_header_sizes():
    assert ctypes.sizeof(abi.ProblemDesc) == 4 * 4 + 2 * 8 + 8 * 8 + 6 * 4 + 8 + 2 * 8 + 2 * 4
    assert ctypes.sizeof(abi.OptimOptions) == 32
    assert ctypes.sizeof(abi.OptimResult) == 6 * 4 + 2 * 8 + 256
    assert ctypes.sizeof(abi.RansacOptions) == 40
    assert ctypes.sizeof(abi.RansacResult) == 16 + 72 + 24


@pytest.mark.skipif(capi.device_count() > 0, reason="needs a machine WITHOUT a GPU")
def test_no_cpu_fallback():
    prob, x0, _ = synth.make_bundle(n_cams=1, n_poses=8)
    with pytest.raises(capi.CalibCudaError):
        capi.RefineHandle(prob)


def test_validation_errors_mirror_reference():
    # std::invalid_argument cases (intrinsics.cpp:92-96, bundle.cpp:136-145, *residual.h create())
    prob, x0, _ = synth.make_intrinsics(n_views=3)
    with pytest.raises(ValueError, match="at least 4"):
        capi.RefineHandle(prob)
    prob, x0, _ = synth.make_bundle(n_cams=1, n_poses=4)
    prob.desc.n_cams = 0
    with pytest.raises(ValueError, match="No camera intrinsics"):
        capi.RefineHandle(prob)
    prob, x0, _ = synth.make_bundle(n_cams=1, n_poses=4)
    prob.block_offset[2] = prob.block_offset[1]  # an empty view
    with pytest.raises(ValueError, match="No observations"):
        capi.RefineHandle(prob)


def test_shared_board_form_is_validated_before_the_device_is_touched():
    """cal_problem_desc.board_n > 0: one board instead of object points per observation.  Its argument checks run
    before the device check, so they are testable here; a valid descriptor then fails with 'no CUDA device'."""
    prob, x0, _ = synth.make_bundle(n_cams=1, n_poses=8)
    pb = prob.with_shared_board()
    if capi.device_count() == 0:
        with pytest.raises(capi.CalibCudaError):
            capi.RefineHandle(pb)
    pb.desc.board_n = 80          # blocks hold 88 observations
    with pytest.raises(ValueError, match="exactly board_n"):
        capi.RefineHandle(pb)
    pb.desc.board_n = 88
    pb.desc.board_y = None
    with pytest.raises(ValueError, match="null observation arrays"):
        capi.RefineHandle(pb)
    pb.desc.board_n = -1
    with pytest.raises(ValueError, match="board_n"):
        capi.RefineHandle(pb)


def test_header_is_plain_c_and_a_c_client_links(tmp_path):
    """include/calib_b200.h is a C header (what cgo / JNI / ctypes bind to): it compiles as C99 and as C++17, and
    the C example links against the shared library (running it needs a GPU: done in the gpu suite)."""
    import subprocess
    inc = os.path.join(ROOT, "include")
    (tmp_path / "t.c").write_text('#include "calib_b200.h"\nint main(void) { return (int)sizeof(cal_problem_desc) == 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", str(tmp_path / "t.c")], check=True)
    (tmp_path / "t.cpp").write_text('#include "calib_b200.h"\nint main() { return 0; }\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", str(tmp_path / "t.cpp")], check=True)
    libdir = os.path.dirname(build.build())
    exe = os.path.join(ROOT, "examples", "_build", "c_api_example")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I", inc, os.path.join(ROOT, "examples", "c_api_example.c"), "-L", libdir, "-lcalib_b200",
                    "-lm", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    assert os.path.exists(exe)
