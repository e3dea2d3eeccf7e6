"""Builders of the C++ host-test binaries (tests/cpp/reference_tests.cpp).  TEST INFRASTRUCTURE.

  real     linked against calibration_b200/_build/libcalib_b200.so — the product; computes only on a GPU
  standin  linked against tests/cpp/abi_standin.cpp -DSTANDIN_SIMT: the C ABI answered by the product's seeding /
           RANSAC kernels under the CPU SIMT shim and by the CPU oracle for the LM solves (CPU suite only)
"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
OUT = os.path.join(CPP, "_build")
INC = os.path.join(ROOT, "include")
EMUL = os.path.join(ROOT, "tests", "host_emul")
CSRC = os.path.join(ROOT, "calibration_b200", "csrc")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
FLAGS = ["-std=c++20", "-O1", "-Wall", "-Wextra", "-Werror", "-I", INC, "-I", CPP]
SOURCES = [os.path.join(CPP, f) for f in ("reference_tests.cpp", "ref_sim.hpp", "mini_gtest.hpp")] + [
    os.path.join(INC, f) for f in ("calib_b200.h", "calib_b200_adapter.hpp", "calib_b200_mini.hpp")]


def _stale(target, deps):
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def build_real():
    from calibration_b200 import build
    lib = build.build()
    libdir = os.path.dirname(lib)
    exe = os.path.join(OUT, "host_tests")
    if _stale(exe, SOURCES + [lib]):
        os.makedirs(OUT, exist_ok=True)
        subprocess.run([CXX, *FLAGS, os.path.join(CPP, "reference_tests.cpp"), "-o", exe, "-L", libdir, "-lcalib_b200",
                        "-Wl,-rpath," + libdir], check=True)
    return exe, dict(os.environ, LD_LIBRARY_PATH=libdir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))


def _simt_lib(name, src, deps):
    so = os.path.join(EMUL, "_build", name)
    src = os.path.join(EMUL, src)
    deps = [src, os.path.join(EMUL, "simt_shim.hpp")] + [os.path.join(CSRC, d) for d in deps]
    if _stale(so, deps):  # the same command tests/test_simt_kernels.py uses
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.run([CXX, "-O2", "-std=c++20", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-o", so, src], check=True)
    return so


def build_standin():
    import oracle_lib
    oracle_lib.lib()
    odir = os.path.join(ROOT, "oracle", "_build")
    sdir = os.path.join(EMUL, "_build")
    libs = [_simt_lib("libransac_simt.so", "ransac_simt.cpp", ("ransac_kernel.cuh", "ransac_plane_kernel.cuh", "ransac_sampler.cuh",
                                                                "ransac_iters.hpp", "dlt.cuh", "plane_math.cuh")),
            _simt_lib("libseed_simt.so", "seed_simt.cpp", ("seed_kernels.cuh", "seed_host.hpp", "dlt.cuh"))]
    exe = os.path.join(OUT, "host_tests_standin")
    standin = os.path.join(CPP, "abi_standin.cpp")
    if _stale(exe, SOURCES + [standin, os.path.join(odir, "liboracle.so")] + libs):
        os.makedirs(OUT, exist_ok=True)
        subprocess.run([CXX, *FLAGS, "-DSTANDIN_SIMT", "-I", os.path.join(ROOT, "oracle"), os.path.join(CPP, "reference_tests.cpp"), standin,
                        "-o", exe, "-L", odir, "-loracle", "-L", sdir, "-lransac_simt", "-lseed_simt", "-pthread",
                        "-Wl,-rpath," + odir, "-Wl,-rpath," + sdir], check=True)
    return exe, dict(os.environ, LD_LIBRARY_PATH=os.pathsep.join([odir, sdir, os.environ.get("LD_LIBRARY_PATH", "")]))


def build_product_on_cpu():
    """reference_tests.cpp over the CPU build of the product's own refinement sources (tests/test_product_on_cpu.py):
    cal_refine_* = refine_host.cu + the kernels under the shim; the linear stage as in build_standin; AX = XB from the oracle"""
    import test_product_on_cpu
    so = test_product_on_cpu._build()
    _, env = build_standin()
    odir, sdir = os.path.join(ROOT, "oracle", "_build"), os.path.join(EMUL, "_build")
    exe = os.path.join(OUT, "host_tests_product")
    standin = os.path.join(CPP, "abi_standin.cpp")
    if _stale(exe, SOURCES + [standin, so]):
        subprocess.run([CXX, *FLAGS, "-DSTANDIN_SIMT", "-DSTANDIN_NO_REFINE", "-I", os.path.join(ROOT, "oracle"), os.path.join(CPP, "reference_tests.cpp"),
                        standin, "-o", exe, "-L", odir, "-loracle", "-L", sdir, "-lransac_simt", "-lseed_simt", "-lcalib_b200_simt", "-pthread",
                        "-Wl,-rpath," + odir, "-Wl,-rpath," + sdir], check=True)
    return exe, env


EXAMPLE = os.path.join(ROOT, "examples", "cpp_adapter_example.cpp")


def build_example(real):
    """examples/cpp_adapter_example.cpp against the product library (real) or the CPU stand-in."""
    if real:
        from calibration_b200 import build
        lib = build.build()
        libdir = os.path.dirname(lib)
        exe = os.path.join(ROOT, "examples", "_build", "cpp_adapter_example")
        if _stale(exe, [EXAMPLE, lib] + SOURCES[3:]):
            os.makedirs(os.path.dirname(exe), exist_ok=True)
            subprocess.run([CXX, *FLAGS, EXAMPLE, "-o", exe, "-L", libdir, "-lcalib_b200", "-Wl,-rpath," + libdir], check=True)
        return exe, dict(os.environ, LD_LIBRARY_PATH=libdir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    _, env = build_standin()
    odir, sdir = os.path.join(ROOT, "oracle", "_build"), os.path.join(EMUL, "_build")
    exe = os.path.join(OUT, "cpp_adapter_example_standin")
    standin = os.path.join(CPP, "abi_standin.cpp")
    if _stale(exe, [EXAMPLE, standin] + SOURCES[3:]):
        subprocess.run([CXX, *FLAGS, "-DSTANDIN_SIMT", "-I", os.path.join(ROOT, "oracle"), EXAMPLE, standin, "-o", exe, "-L", odir, "-loracle",
                        "-L", sdir, "-lransac_simt", "-lseed_simt", "-pthread", "-Wl,-rpath," + odir, "-Wl,-rpath," + sdir], check=True)
    return exe, env


def run(exe, env, *filters, timeout=600):
    return subprocess.run([exe, *filters], capture_output=True, text=True, timeout=timeout, env=env)


E2E_EXAMPLE = os.path.join(ROOT, "examples", "cpp_bundle_e2e.cpp")


def build_bundle_e2e():
    """examples/cpp_bundle_e2e.cpp (optimize_bundle from std::vector<BundleObservation> at BASELINE configs[4] scale) against the product library."""
    from calibration_b200 import build
    lib = build.build()
    libdir = os.path.dirname(lib)
    exe = os.path.join(ROOT, "examples", "_build", "cpp_bundle_e2e")
    if _stale(exe, [E2E_EXAMPLE, lib] + SOURCES[3:]):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.run([CXX, "-std=c++20", "-O2", "-Wall", "-Wextra", "-Werror", "-pthread", "-I", INC, E2E_EXAMPLE, "-o", exe, "-L", libdir, "-lcalib_b200",
                        "-Wl,-rpath," + libdir], check=True)
    return exe, dict(os.environ, LD_LIBRARY_PATH=libdir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
