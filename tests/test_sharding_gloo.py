"""The N > 1 path on CPU: two gloo ranks each evaluate their shard with the oracle, sum-allreduce the
normal equations (the one exchange step of the path) and must reproduce the single-process result;
a replicated Gauss-Newton step computed from the reduced system is then identical on both ranks."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle_lib as O
    from calibration_b200 import sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kw = dict(seed=137, n_cams=3, n_poses=64, chunk=16)
    # (a) chunk sharding as bench.py does it: the shard holds exactly the bytes of the full problem
    mine, x0, _ = synth.make_bundle(chunks=sharding.chunk_shard(64, 16, rank, world), **kw)
    full, _, _ = synth.make_bundle(**kw)
    c, g, H = O.refine_eval(mine, x0, threads=1)
    buf = torch.from_numpy(np.concatenate([[c], g, H.ravel()]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    cf, gf, Hf = O.refine_eval(full, x0, threads=1)
    ref = np.concatenate([[cf], gf, Hf.ravel()])
    err = float(np.abs(buf.numpy() - ref).max() / np.abs(ref).max())
    # (b) observation-balanced block partition of an existing problem
    part = sharding.shard_problem(full, rank, world)
    c2, g2, H2 = O.refine_eval(part, x0, threads=1)
    buf2 = torch.from_numpy(np.concatenate([[c2], g2, H2.ravel()]))
    dist.all_reduce(buf2, op=dist.ReduceOp.SUM)
    err2 = float(np.abs(buf2.numpy() - ref).max() / np.abs(ref).max())
    # replicated step from the reduced system: bitwise identical across ranks (allreduce result is)
    n = len(g)
    step = np.linalg.solve(buf.numpy()[1 + n:].reshape(n, n) + 1e-6 * np.eye(n), buf.numpy()[1:1 + n])
    gathered = [torch.zeros(n, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(step))
    same = all(torch.equal(gathered[0], t) for t in gathered)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([err, err2, float(same), part.desc.n_obs]))
    dist.destroy_process_group()


def test_two_rank_allreduce_reproduces_single_process(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"r{r}.npy") for r in range(world)]
    for r in res:
        assert r[0] < 1e-12 and r[1] < 1e-12 and r[2] == 1.0
    assert sum(r[3] for r in res) == 3 * 64 * 88


def test_partition_blocks_balanced_and_contiguous():
    from calibration_b200 import sharding
    rng = np.random.default_rng(0)
    lens = rng.integers(1, 200, size=1000)
    off = np.concatenate([[0], np.cumsum(lens)])
    for world in (1, 2, 4, 8):
        b = sharding.partition_blocks(off, world)
        assert b[0] == 0 and b[-1] == 1000 and (np.diff(b) >= 0).all()
        per = np.diff(off[b])
        assert per.sum() == off[-1] and per.max() - per.min() <= 2 * lens.max()
    with pytest.raises(ValueError):
        sharding.chunk_shard(100, 16, 0, 4)


def test_view_shards_partition_the_per_view_kinds():
    """shard_views: every residual block lands in exactly one shard together with its view's pose block, the shared
    blocks are replicated, and the oracle's cost is additive over the shards (the per-view kinds' exchange step adds
    the Schur complement on top, checked on two GPUs by tools/mgpu_views_check.py)."""
    import sys
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle_lib as O
    from calibration_b200 import sharding, synth
    for mk, n_shared in ((lambda: synth.make_extrinsics(n_cams=3, n_views=41, drop_fraction=0.3), 3 * 17), (lambda: synth.make_intrinsics(n_views=23), 10)):
        prob, x0, _ = mk()
        c_full, _, _ = O.refine_eval(prob, x0)
        for world in (2, 3):
            tot, blocks, views = 0.0, 0, []
            for r in range(world):
                sub, xl, (v0, v1) = sharding.shard_views(prob, x0, r, world)
                assert sub.desc.view_base == v0 and sub.desc.n_views == v1 - v0 and len(xl) == n_shared + 7 * (v1 - v0)
                assert np.array_equal(xl[:n_shared], x0[:n_shared])
                back = sharding.gather_views(xl, np.zeros_like(x0), prob, (v0, v1))
                nv = prob.desc.n_views
                assert np.array_equal(back[n_shared + 4 * v0:n_shared + 4 * v1], x0[n_shared + 4 * v0:n_shared + 4 * v1])
                assert np.array_equal(back[n_shared + 4 * nv + 3 * v0:n_shared + 4 * nv + 3 * v1], x0[n_shared + 4 * nv + 3 * v0:n_shared + 4 * nv + 3 * v1])
                c, _, _ = O.refine_eval(sub, xl)
                tot += c; blocks += int(sub.desc.n_blocks); views.append((v0, v1))
            assert blocks == prob.desc.n_blocks and views[0][0] == 0 and views[-1][1] == prob.desc.n_views
            assert all(views[i][1] == views[i + 1][0] for i in range(world - 1))
            assert abs(tot - c_full) <= 1e-12 * c_full
