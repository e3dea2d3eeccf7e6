"""registers / stack / spills / static shared memory of every kernel of a .cu (nvcc -Xptxas -v), one line per kernel, in the
format of profiles/r2_ptxas_v.txt.   python tools/ptxas_table.py refine_kernels.cu [...]"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "calibration_b200", "csrc")
for src in sys.argv[1:]:
    cmd = ["nvcc", "-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
           "-diag-suppress", "170", "-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", "/tmp/_ptxas_table.o"]
    err = subprocess.run(cmd, capture_output=True, text=True).stderr
    cur = None
    for line in err.splitlines():
        m = re.search(r"Function properties for (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m: stack, ss, sl = m.groups(); continue
        m = re.search(r"Used (\d+) registers(.*)", line)
        if m and cur:
            sm = re.search(r"(\d+) bytes smem", m.group(2))
            print(f"{cur} | {m.group(1)} | {stack} | {ss} | {sl} | {sm.group(1) if sm else 0}")
            cur = None
