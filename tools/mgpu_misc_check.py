"""Two (or more) ranks, launched with torch.distributed.run: the multi-GPU paths besides the bundle pass.
  1. covariance of a per-view kind (optimize_extrinsics) sharded BY VIEWS: every rank's [na][na] matrix equals the
     corresponding rows / columns of the single-GPU covariance of the whole problem;
  2. optimize_handeye (AX = XB) with the pair tiles sharded over the ranks and the 28 sums all-reduced: same normal
     equations and the same solution as one GPU;
  3. batched RANSAC split by problem (no communication): every rank's slice equals the slice of the whole batch.
"""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from calibration_b200 import abi, capi, geometry as G, sharding, synth


def watchdog(limit):
    def run():
        time.sleep(limit); sys.stderr.write("WATCHDOG\n"); sys.stderr.flush(); os._exit(3)
    threading.Thread(target=run, daemon=True).start()


def main():
    watchdog(float(os.environ.get("MAX_SECONDS", "150")))
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = capi.Comm(uid[0], rank, world, local)

    def all_gather(b):
        out = [None] * world; dist.all_gather_object(out, b); return out
    comm.enable_peer(all_gather)

    def report(name, ok, detail):
        res = all_gather((bool(ok), detail))
        if rank == 0:
            print(f"{name}: {'OK' if all(r[0] for r in res) else 'FAILED'} {res}", flush=True)

    # ---- 1. covariance under view sharding ----
    n_views = 400
    prob, x0, _ = synth.make_extrinsics(n_cams=2, n_views=n_views)
    opts = abi.OptimOptions.default(compute_covariance=1)
    h = capi.RefineHandle(prob, device=local); x_ref, r_ref, cov_ref = h.solve(x0, opts); h.close()
    sub, xl, vr = sharding.shard_views(prob, x0, rank, world)
    h = capi.RefineHandle(sub, device=local); h.attach_comm(comm)
    x_loc, r_loc, cov_loc = h.solve(xl, opts); h.close()
    n_shared = len(x0) - 7 * n_views
    v0, v1 = vr
    idx = np.concatenate([np.arange(n_shared), n_shared + 4 * v0 + np.arange(4 * (v1 - v0)), n_shared + 4 * n_views + 3 * v0 + np.arange(3 * (v1 - v0))])
    sub_ref = cov_ref[np.ix_(idx, idx)]
    err = float(np.abs(cov_loc - sub_ref).max() / np.abs(sub_ref).max())
    report("covariance of optimize_extrinsics sharded by views", bool(r_loc.covariance_ok) and bool(r_ref.covariance_ok) and err <= 1e-6,
           {"rel_err": err, "iters": (int(r_loc.iterations), int(r_ref.iterations)), "views": vr})

    # ---- 2. AX = XB, pair tiles sharded ----
    bg, ct, X_gt = synth.make_handeye_poses(seed=3, n=1500)
    rng = np.random.default_rng(0)
    ct = [synth.perturb_pose(rng, T, 0.3, 0.002) for T in ct]
    xh0 = G.pack_handeye(synth.perturb_pose(rng, X_gt, 2.0, 0.01))
    h1 = capi.AxxbHandle.from_poses(bg, ct, 0.05, device=local)
    c1, g1, H1 = h1.eval(xh0); x1, r1, cov1 = h1.solve(xh0); ms1 = h1.bench_pass(xh0, 20) / 20
    hs = capi.AxxbHandle.from_poses(bg, ct, 0.05, device=local); hs.attach_comm(comm)
    cs, gs, Hs = hs.eval(xh0); xs, rs, covs = hs.solve(xh0)
    dist.barrier(); mss = hs.bench_pass(xh0, 20) / 20
    ok = (abs(cs - c1) <= 1e-12 * c1 and np.abs(gs - g1).max() <= 1e-10 * np.abs(g1).max() and np.abs(Hs - H1).max() <= 1e-10 * np.abs(H1).max()
          and np.abs(xs - x1).max() <= 1e-9 and rs.iterations == r1.iterations and np.abs(covs - cov1).max() <= 1e-6 * np.abs(cov1).max())
    report("optimize_handeye with sharded pair tiles", ok, {"dx": float(np.abs(xs - x1).max()), "iters": (int(rs.iterations), int(r1.iterations)),
                                                            "pass_ms_one_gpu": ms1, "pass_ms_sharded": mss, "pairs": int(h1.n_pairs)})
    hs.close(); h1.close()

    # ---- 3. RANSAC split by problem ----
    x, y, u, v, _ = synth.synth_ransac(seed=31, n_problems=4000, n=500)
    ro = abi.RansacOptions.default()
    r_all, m_all = capi.ransac_homography_batch(x, y, u, v, ro, device=local)
    per = (4000 + world - 1) // world
    p0, p1 = rank * per, min(4000, (rank + 1) * per)
    o = abi.RansacOptions.default(); o.seed = ro.seed + p0
    r_my, m_my = capi.ransac_homography_batch(x[p0:p1], y[p0:p1], u[p0:p1], v[p0:p1], o, device=local)
    same = np.array_equal(m_my, m_all[p0:p1]) and bytes(r_my) == bytes(r_all)[p0 * len(bytes(r_all)) // 4000:p1 * len(bytes(r_all)) // 4000]
    r_multi, m_multi = capi.ransac_homography_batch(x, y, u, v, ro, devices=list(range(world)))
    same = same and np.array_equal(m_multi, m_all) and bytes(r_multi) == bytes(r_all)
    report("batched RANSAC split by problem (per rank and cal_ransac_homography_batch_multi)", same, {"slice": (p0, p1)})
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
