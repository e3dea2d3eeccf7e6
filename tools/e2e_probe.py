"""Wall-clock breakdown of the end-to-end call (create + solve + destroy) at C5 scale."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from calibration_b200 import abi, capi, synth
n_poses = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
prob, x0, xgt = synth.make_bundle(seed=137, n_cams=8, n_poses=n_poses, pinned=True)
capi.RefineHandle(synth.make_bundle(n_cams=2, n_poses=64)[0]).close()  # context + module load
for rep in range(3):
    t0 = time.perf_counter(); h = capi.RefineHandle(prob); t1 = time.perf_counter()
    x, res, cov = h.solve(x0); t2 = time.perf_counter()
    h.close(); t3 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms  solve {1e3*(t2-t1):.1f} ms ({res.iterations} it, {res.num_jac_evals} jac, {res.num_cost_evals} cost)  destroy {1e3*(t3-t2):.1f} ms  total {1e3*(t3-t0):.1f} ms")
