"""Builds libcalib_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with
the repository snapshot.  `python -m calibration_b200.build` rebuilds.
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# CALIB_B200_BUILD_DIR / CALIB_B200_NVCC_EXTRA: a second build of the library beside the default one (kernel experiments
# are compared in ONE process sequence on the same GPU box: tools/k1_probe.py); the default build ignores both
OUT_DIR = os.path.join(_HERE, os.environ.get("CALIB_B200_BUILD_DIR", "_build"))
LIB = os.path.join(OUT_DIR, "libcalib_b200.so")
NVCC_EXTRA = os.environ.get("CALIB_B200_NVCC_EXTRA", "").split()
SOURCES = ["k1_fused.cu", "refine_kernels.cu", "refine_host.cu", "axxb.cu", "ransac.cu", "ransac_plane.cu", "seed.cu", "comm.cpp", "comm_peer.cu", "dataset.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
    "-diag-suppress", "170",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_HERE, "..", "include", "calib_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    env = dict(os.environ)
    # the image exports CXX=/opt/gcc/bin/g++ (no libgomp); nvcc should use the distro g++
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(OUT_DIR, src.rsplit(".", 1)[0] + ".o")
        cmd = [_nvcc(), *ccbin, *NVCC_FLAGS, *NVCC_EXTRA, "-c", path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas"); cmd.insert(2, "-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode(errors="replace"))
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [_nvcc(), *ccbin, "-shared", "-o", LIB, *objs, "-lcudart", "-ldl"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
