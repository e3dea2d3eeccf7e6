"""Wall time of the full solves of the small named shapes (C1 intrinsics, C3 extrinsics, C4 bundle), with and without
covariance; CALIB_B200_TRACE=1 prints the phases of cal_refine_create / cal_refine_solve on stderr."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from calibration_b200 import abi, capi, synth

mk = {"c1": lambda: synth.make_intrinsics(), "c3": lambda: synth.make_extrinsics(n_cams=2, n_views=1000),
      "c4": lambda: synth.make_bundle(seed=137, n_cams=4, n_poses=5000)}
for name in (sys.argv[1:] or ["c1", "c3", "c4"]):
    prob, x0, _ = mk[name]()
    for cov in (0, 1):
        opts = abi.OptimOptions.default(compute_covariance=cov)
        h = capi.RefineHandle(prob); h.solve(x0, opts, want_cov=bool(cov)); h.close()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); h = capi.RefineHandle(prob); t1 = time.perf_counter()
            x, res, c = h.solve(x0, opts, want_cov=bool(cov)); t2 = time.perf_counter(); n = h.launch_count(); h.close()
            ts.append((1e3 * (t1 - t0), 1e3 * (t2 - t1)))
        print(json.dumps({"workload": name, "covariance": cov, "create_ms": [round(a, 3) for a, _ in ts], "solve_ms": [round(b, 3) for _, b in ts],
                          "iterations": int(res.iterations), "launches": int(n), "n_obs": int(prob.desc.n_obs)}), flush=True)
