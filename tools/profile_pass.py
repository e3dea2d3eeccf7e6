"""Small driver for ncu: a few fused passes of the bundle hot path on a C5-shaped problem."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from calibration_b200 import capi, synth

n_poses = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fixed = len(sys.argv) > 3 and sys.argv[3] == "fixed"
prob, x0, _ = synth.make_bundle(seed=137, n_cams=8, n_poses=n_poses, optimize_intrinsics=not fixed)
h = capi.RefineHandle(prob)
for _ in range(2):
    h.bench_pass(x0, reps=1, jacobian=True)
ms, k1, cost = h.bench_pass(x0, reps=reps, jacobian=True)
msc, kc, _ = h.bench_pass(x0, reps=reps, jacobian=False)
print(f"n_obs={prob.desc.n_obs} pass {ms/reps:.3f} ms (k1 {k1/reps:.3f} ms) cost-pass {msc/reps:.3f} ms (kernel {kc/reps:.3f}) cost={cost:.6e}")
