"""K1 timing probe: device time of the fused pass (CUDA events inside cal_refine_bench_pass) for a bundle of
n_cams x n_poses views of a rows x cols board.  Used to A/B kernel variants built into different directories:
    CALIB_B200_BUILD_DIR=_build_x python tools/k1_probe.py 8 100000 8 11
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calibration_b200 import capi, synth  # noqa: E402

n_cams, n_poses, rows, cols = (int(a) for a in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
prob, x0, _ = synth.make_bundle(seed=137, n_cams=n_cams, n_poses=n_poses, rows=rows, cols=cols, chunk=12500)
h = capi.RefineHandle(prob, device=0)
for _ in range(3):
    h.bench_pass(x0, reps=1, jacobian=True)
ms, k1, cost = h.bench_pass(x0, reps=reps, jacobian=True)
n = int(prob.desc.n_obs)
print(json.dumps({"build": os.environ.get("CALIB_B200_BUILD_DIR", "_build"), "n_obs": n, "corners": rows * cols, "pass_ms": ms / reps, "k1_ms": k1 / reps,
                  "k1_ps_per_obs": 1e9 * k1 / reps / n, "cost": cost, "info": h.layout_info()}))
h.close()
