"""optimize_extrinsics at scale (per-view pose unknowns -> Schur path): timing breakdown of a solve."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from calibration_b200 import abi, capi, synth
n_cams = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_views = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
t0 = time.time(); prob, x0, xgt = synth.make_extrinsics(n_cams=n_cams, n_views=n_views); gen = time.time() - t0
capi.RefineHandle(synth.make_bundle(n_cams=2, n_poses=64)[0]).close()
t0 = time.perf_counter(); h = capi.RefineHandle(prob); t1 = time.perf_counter()
for _ in range(2): h.bench_pass(x0, reps=1, jacobian=True)
ms, k1, _ = h.bench_pass(x0, reps=5, jacobian=True)
t2 = time.perf_counter(); x, res, _ = h.solve(x0, abi.OptimOptions.default(compute_covariance=0)); t3 = time.perf_counter()
rms, g = h.view_errors(x)
print(f"n_obs={prob.desc.n_obs} n_blocks={prob.desc.n_blocks} n_tan={h.n_tan} gen {gen:.1f}s create {1e3*(t1-t0):.1f} ms  pass {ms/5:.3f} ms (k1 {k1/5:.3f})  "
      f"solve {1e3*(t3-t2):.1f} ms: {res.report.decode()} jac {res.num_jac_evals} cost {res.num_cost_evals} launches {h.launch_count()}  global rms {g:.4f} px")
h.close()
