"""Compare the device code of the current build with the one a git revision produces, function by function.

    python tools/sass_diff.py <git-rev>

Builds the kernel translation units of <git-rev> in a scratch directory with build.py's flags (nvcc cross-compiles
without a GPU) and compares `cuobjdump -sass` per function with calibration_b200/_build/*.o (hashes of anonymous
namespaces, which depend on the source path, are normalised).  Used when kernels are moved between files: the move
must not change a single instruction.
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from calibration_b200 import build  # noqa: E402

UNITS = ["refine_kernels", "k1_fused", "axxb", "ransac", "ransac_plane", "seed", "comm_peer"]


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    norm = lambda t: re.sub(r"_GLOBAL__N__[0-9a-f]{8}_", "_GLOBAL__N__X_", t)
    d, name, buf = {}, None, []
    for ln in out.split("\n"):
        m = re.search(r"Function : (\S+)", ln)
        if m:
            if name:
                d[name] = "\n".join(buf)
            name, buf = norm(m.group(1)), []
        elif name and not re.match(r"^\s*/\*[0-9a-f]{4}\*/\s*$", ln):
            buf.append(norm(ln.rstrip()))
    if name:
        d[name] = "\n".join(buf)
    return d


def main(rev):
    build.build()
    rc = 0
    with tempfile.TemporaryDirectory() as tmp:
        tar = subprocess.run(["git", "-C", ROOT, "archive", rev, "calibration_b200/csrc", "include"], capture_output=True, check=True).stdout
        subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
        csrc = os.path.join(tmp, "calibration_b200", "csrc")
        ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
        procs = []
        for u in UNITS:
            if os.path.exists(os.path.join(csrc, u + ".cu")):
                procs.append((u, subprocess.Popen(["nvcc", *ccbin, *build.NVCC_FLAGS, "-c", u + ".cu", "-o", os.path.join(tmp, u + ".o")], cwd=csrc,
                                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)))
        for u, p in procs:
            if p.wait() != 0:
                print(f"{u}: nvcc failed at {rev}"); rc = 1; continue
            a, b = functions(os.path.join(tmp, u + ".o")), functions(os.path.join(build.OUT_DIR, u + ".o"))
            gone, new = sorted(set(a) - set(b)), sorted(set(b) - set(a))
            diff = sorted(k for k in a if k in b and a[k] != b[k])
            print(f"{u}: {len(a)} -> {len(b)} functions, changed: {diff or 'none'}, removed: {gone or 'none'}, added: {new or 'none'}")
            rc |= bool(diff)
    return rc


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "HEAD~1"))
