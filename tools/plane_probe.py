"""Throughput probe of the batched plane RANSAC kernel (no torch: device buffers through cuda-python's runtime
bindings); prints one JSON line.  usage: plane_probe.py [n_problems] [cpu]"""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cuda.bindings import runtime as rt
from calibration_b200 import abi, capi, synth


def ck(r):
    assert int(r[0]) == 0, r[0]
    return r[1] if len(r) > 1 else None


npb, n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000, 500
t0 = time.time(); x, y, z, _ = synth.synth_plane_ransac(seed=23, n_problems=npb, n=n); gen = time.time() - t0
L = capi.lib()
nb = npb * n * 8
dev = []
for a in (x, y, z):
    p = ck(rt.cudaMalloc(nb)); ck(rt.cudaMemcpy(p, np.ascontiguousarray(a).ctypes.data, nb, rt.cudaMemcpyKind.cudaMemcpyHostToDevice)); dev.append(p)
res = ck(rt.cudaMalloc(npb * C.sizeof(abi.PlaneResult))); mask = ck(rt.cudaMalloc(npb * n))
opts = abi.RansacOptions.default(thresh=0.006)
ms = C.c_float(); times = []
for rep in range(4):
    rc = L.cal_ransac_plane_batch_dev(npb, n, *[C.c_void_p(int(d)) for d in dev], C.byref(opts), 1, C.c_void_p(int(res)), C.c_void_p(int(mask)), C.byref(ms))
    assert rc == 0, L.cal_last_error()
    times.append(ms.value)
host = (abi.PlaneResult * npb)()
ck(rt.cudaMemcpy(C.addressof(host), res, C.sizeof(host), rt.cudaMemcpyKind.cudaMemcpyDeviceToHost))
runs = np.array([h.iters_run for h in host]); inl = np.array([h.n_inliers for h in host]); ok = sum(h.success for h in host)
t = min(times[1:]) * 1e-3
out = {"kernel": "k_ransac_plane", "problems": npb, "n": n, "ms": t * 1e3, "ms_all": times, "problems_per_s": npb / t,
       "hypotheses": int(runs.sum()), "mean_iters_run": float(runs.mean()), "hyp_point_scores_per_s": float(runs.sum()) * n * 2 / t,
       "success": int(ok), "mean_inliers": float(inl.mean()), "hbm_read_GBps": 24.0 * npb * n / t / 1e9, "gen_s": gen}
print(json.dumps(out))
if len(sys.argv) > 2 and sys.argv[2] == "cpu":
    import oracle_lib as O
    k = 4000
    t0 = time.perf_counter(); O.ransac_plane_batch(x[:k], y[:k], z[:k], opts); dt = time.perf_counter() - t0
    print(json.dumps({"cpu_oracle_problems_per_s": k / dt, "cores": os.cpu_count(), "sample": k}))
