"""Two (or more) ranks: optimize_extrinsics sharded BY VIEWS — per-view pose blocks, their 6x6 factors and the
back-substitution stay on the owning GPU, the Schur complement and the per-view scalars are all-reduced — must
reproduce the single-GPU solve of the whole problem.  Launch with torch.distributed.run."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from calibration_b200 import abi, capi, sharding, synth


def watchdog(limit):
    def run():
        time.sleep(limit); sys.stderr.write("WATCHDOG\n"); sys.stderr.flush(); os._exit(3)
    threading.Thread(target=run, daemon=True).start()


def main():
    watchdog(float(os.environ.get("MAX_SECONDS", "100")))
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = capi.Comm(uid[0], rank, world, local)

    def all_gather(b):
        out = [None] * world; dist.all_gather_object(out, b); return out
    comm.enable_peer(all_gather)
    opts = abi.OptimOptions.default(compute_covariance=0)
    for n_views in (600, 9000):
        prob, x0, _ = synth.make_extrinsics(n_cams=2, n_views=n_views)
        h = capi.RefineHandle(prob, device=local); x_ref, r_ref, _ = h.solve(x0, opts); h.close()
        sub, xl, vr = sharding.shard_views(prob, x0, rank, world)
        h = capi.RefineHandle(sub, device=local); h.attach_comm(comm)
        t0 = time.perf_counter(); x_loc, r_loc, _ = h.solve(xl, opts); dt = time.perf_counter() - t0
        h.close()
        x_mine = sharding.gather_views(x_loc, x_ref, prob, vr)   # other ranks' views are taken from the reference
        err = float(np.abs(x_mine - x_ref).max() / np.abs(x_ref).max())
        ok = err <= 1e-8 and r_loc.iterations == r_ref.iterations and bool(r_loc.success) == bool(r_ref.success)
        res = [None] * world
        dist.all_gather_object(res, (ok, err, int(r_loc.iterations), int(r_ref.iterations), dt))
        if rank == 0:
            print(f"n_views={n_views} blocks={prob.desc.n_blocks} peer={getattr(comm, 'peer', False)} per-rank (ok, rel err, iters, ref iters, s): {res}", flush=True)
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
