"""Throughput probes for the secondary kernels (RANSAC K3, AX=XB) at BASELINE sizes; prints JSON lines."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from calibration_b200 import abi, capi, synth

what = sys.argv[1] if len(sys.argv) > 1 else "ransac"
if what == "ransac":
    npb, n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000, 500
    t0 = time.time(); x, y, u, v, _ = synth.synth_ransac(seed=17, n_problems=npb, n=n); gen = time.time() - t0
    dev = [torch.from_numpy(a).cuda() for a in (x, y, u, v)]
    res = torch.empty(npb * C.sizeof(abi.RansacResult), dtype=torch.uint8, device="cuda")
    mask = torch.empty(npb * n, dtype=torch.uint8, device="cuda")
    opts = abi.RansacOptions.default()
    ms = C.c_float()
    L = capi.lib()
    for rep in range(3):
        rc = L.cal_ransac_homography_batch_dev(npb, n, *[C.c_void_p(d.data_ptr()) for d in dev], C.byref(opts), 1,
                                               C.c_void_p(res.data_ptr()), C.c_void_p(mask.data_ptr()), C.byref(ms))
        assert rc == 0, L.cal_last_error()
    r = np.frombuffer(res.cpu().numpy().tobytes(), dtype=np.dtype([("success", "i4"), ("iters", "i4"), ("n_inliers", "i4"), ("iters_run", "i4"),
                                                                    ("h", "f8", 9), ("rms", "f8"), ("sym", "f8"), ("mm", "f8")]))
    hyp = int(r["iters_run"].sum())
    t0 = time.perf_counter(); capi.ransac_homography_batch(x, y, u, v, opts, want_mask=True); e2e = time.perf_counter() - t0
    out = {"kernel": "k_ransac", "problems": npb, "n": n, "ms": ms.value, "problems_per_s": npb / (ms.value * 1e-3),
           "hypotheses": hyp, "mean_iters_run": hyp / npb, "hyp_point_scores_per_s": hyp * n * 2 / (ms.value * 1e-3),
           "success": int(r["success"].sum()), "mean_inliers": float(r["n_inliers"].mean()),
           "hbm_read_GBps": 32.0 * npb * n / (ms.value * 1e-3) / 1e9, "e2e_host_call_s": e2e, "gen_s": gen}
    print(json.dumps(out))
    if len(sys.argv) > 3 and sys.argv[3] == "cpu":
        import oracle_lib as O
        k = 2000
        t0 = time.perf_counter(); O.ransac_batch(x[:k], y[:k], u[:k], v[:k], opts); dt = time.perf_counter() - t0
        print(json.dumps({"cpu_oracle_problems_per_s": k / dt, "cores": os.cpu_count(), "sample": k}))
elif what == "seed":
    # seeding stage at C5 scale: estimate_intrinsics for 8 cameras x n_poses views, then estimate_planar_pose
    n_poses = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    prob, _, xgt = synth.make_bundle(seed=137, n_cams=8, n_poses=n_poses)
    off = np.asarray(prob.block_offset); cam = np.asarray(prob.block_cam)
    dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (prob.x, prob.y, prob.u, prob.v)]
    nv = len(off) - 1
    kmtx = np.zeros((8, 5)); cam_ok = np.zeros(8, dtype=np.int32); poses = np.zeros((nv, 12)); ok = np.zeros(nv, dtype=np.int32)
    opts = abi.SeedOptions.from_bounds(None)
    dptrs = [C.cast(C.c_void_p(d.data_ptr()), abi.c_double_p) for d in dev]
    off64 = off.astype(np.int64); cam32 = cam.astype(np.int32)
    L = capi.lib()
    def run_intr():
        rc = L.cal_seed_intrinsics(nv, abi.i64ptr(off64), abi.i32ptr(cam32), *dptrs, 8, C.byref(opts), 0, abi.dptr(kmtx), abi.i32ptr(cam_ok),
                                   abi.i32ptr(ok), None, None, abi.dptr(poses))
        assert rc == 0, L.cal_last_error()
    def run_pose():
        rc = L.cal_seed_planar_poses(nv, abi.i64ptr(off64), abi.i32ptr(cam32), *dptrs, 8, abi.dptr(kmtx), 0, abi.dptr(poses), abi.i32ptr(ok))
        assert rc == 0, L.cal_last_error()
    run_intr(); run_pose()
    t0 = time.perf_counter(); run_intr(); ti = time.perf_counter() - t0
    t0 = time.perf_counter(); run_pose(); tp = time.perf_counter() - t0
    k_gt = xgt[:80].reshape(8, 10)[:, :5]
    print(json.dumps({"stage": "seed", "views": nv, "observations": int(off[-1]), "estimate_intrinsics_ms": ti * 1e3,
                      "estimate_planar_pose_ms": tp * 1e3, "views_per_s_intrinsics": nv / ti, "views_per_s_pose": nv / tp,
                      "note": "observations resident on the device; includes D2H of 96 B/view poses and the host Zhang step",
                      "max_rel_K_error_vs_gt": float(np.abs(kmtx[:, :4] / k_gt[:, :4] - 1).max())}))
elif what == "solves":
    # wall time of the full solves of the small BASELINE configs (latency-bound)
    from calibration_b200 import geometry as G
    capi.RefineHandle(synth.make_bundle(n_cams=2, n_poses=64)[0]).close()
    out = {}
    only = int(sys.argv[2]) if len(sys.argv) > 2 else -1
    for idx, (name, mk) in enumerate((("C1 intrinsics 20 views x 54", lambda: synth.make_intrinsics()),
                     ("C3 extrinsics 2 cams x 1000 views x 88", lambda: synth.make_extrinsics(n_cams=2, n_views=1000)),
                     ("C4 bundle 4 cams x 5000 poses x 88", lambda: synth.make_bundle(n_cams=4, n_poses=5000)))):
        if only >= 0 and idx != only:
            continue
        prob, x0, _ = mk()
        print("running", name, flush=True)
        h = capi.RefineHandle(prob); h.solve(x0); h.close()
        t0 = time.perf_counter(); h = capi.RefineHandle(prob); t1 = time.perf_counter(); x, res, cov = h.solve(x0); t2 = time.perf_counter(); h.close()
        out[name] = {"create_ms": 1e3 * (t1 - t0), "solve_ms": 1e3 * (t2 - t1), "iterations": int(res.iterations), "n_obs": int(prob.desc.n_obs),
                     "ms_per_iteration": 1e3 * (t2 - t1) / max(int(res.iterations), 1)}
    print(json.dumps(out))
elif what == "axxb_otf":
    # optimize_handeye at C4 scale straight from the poses: pairs formed on the fly (cal_axxb_create_from_poses)
    from calibration_b200 import geometry as G
    n_poses = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    bg, ct, X_gt = synth.make_handeye_poses(seed=3, n=n_poses)
    rng = np.random.default_rng(0)
    ct = [synth.perturb_pose(rng, T, 0.3, 0.002) for T in ct]
    t0 = time.perf_counter(); h = capi.AxxbHandle.from_poses(bg, ct, 0.05); create = time.perf_counter() - t0
    x0 = G.pack_handeye(synth.perturb_pose(rng, X_gt, 2.0, 0.01))
    h.eval(x0)
    t0 = time.perf_counter()
    for _ in range(5): h.eval(x0)
    dt = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter(); xs, res, cov = h.solve(x0); ts = time.perf_counter() - t0
    print(json.dumps({"kernel": "k_axxb_otf", "poses": n_poses, "pairs_kept": h.n_pairs, "eval_ms_incl_sync": dt * 1e3,
                      "pairs_per_s": h.n_pairs / dt, "create_s": create, "solve_s": ts, "report": res.report.decode(),
                      "param_err": float(np.abs(xs - G.pack_handeye(X_gt)).max())}))
else:
    import oracle_lib as O
    n_poses = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    rng = np.random.default_rng(1)
    # synthetic motion pairs directly (pair construction is a host-side "next" row)
    npairs = n_poses * (n_poses - 1) // 2
    from calibration_b200 import geometry as G
    X = G.make_pose([0.02, -0.01, 0.09], [0.2, 0.9, 0.1], 0.17)
    A = synth.random_poses(rng, npairs, 25.0, (-0.1, 0.1), 0.1)
    Xi = G.inv_pose(X)
    B = Xi[None] @ A @ X[None]
    ra = np.ascontiguousarray(A[:, :3, :3].reshape(npairs, 9)); rb = np.ascontiguousarray(B[:, :3, :3].reshape(npairs, 9))
    ta = np.ascontiguousarray(A[:, :3, 3]) + rng.normal(scale=1e-3, size=(npairs, 3)); tb = np.ascontiguousarray(B[:, :3, 3])
    t0 = time.perf_counter(); h = capi.AxxbHandle(ra, rb, ta, tb, 1.0); create = time.perf_counter() - t0
    x0 = G.pack_handeye(synth.perturb_pose(rng, X, 2.0, 0.01))
    h.eval(x0)
    t0 = time.perf_counter()
    for _ in range(5): h.eval(x0)
    dt = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter(); xs, res, cov = h.solve(x0); ts = time.perf_counter() - t0
    print(json.dumps({"kernel": "k_axxb", "pairs": npairs, "eval_ms_incl_sync": dt * 1e3, "pairs_per_s": npairs / dt,
                      "hbm_GBps": 192.0 * npairs / dt / 1e9, "create_s": create, "solve_s": ts, "report": res.report.decode(),
                      "param_err": float(np.abs(xs - G.pack_handeye(X)).max())}))
