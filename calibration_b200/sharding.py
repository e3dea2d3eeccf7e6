"""View sharding for one-process-per-GPU runs.

Residual blocks are independent given the parameters, so a rank owns a
contiguous range of blocks; the only exchange is the sum-allreduce of the
per-camera normal-equation blocks after every pass.
"""
import numpy as np


def chunk_shard(n_units, chunk, rank, world):
    """Chunk ids (of `chunk` units each) owned by `rank`; the chunk count must divide evenly."""
    n_chunks = (n_units + chunk - 1) // chunk
    if n_chunks % world != 0:
        raise ValueError(f"{n_chunks} chunks are not divisible by {world} ranks")
    per = n_chunks // world
    return list(range(rank * per, (rank + 1) * per))


def partition_blocks(block_offset, world):
    """Contiguous block ranges balanced by observation count: returns world + 1 block boundaries."""
    block_offset = np.asarray(block_offset, dtype=np.int64)
    n_obs = int(block_offset[-1])
    targets = (np.arange(1, world) * n_obs) / world
    cuts = np.searchsorted(block_offset, targets, side="left")
    bounds = np.concatenate([[0], cuts, [len(block_offset) - 1]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def shard_problem(problem, rank, world):
    """A bundle Problem restricted to the blocks of `rank` (parameters are shared by all ranks)."""
    from . import abi
    d = problem.desc
    if d.kind != abi.KIND_BUNDLE:
        raise ValueError("only the bundle kind (no per-view unknowns) is sharded through this helper")
    b = partition_blocks(problem.block_offset, world)
    b0, b1 = int(b[rank]), int(b[rank + 1])
    o0, o1 = int(problem.block_offset[b0]), int(problem.block_offset[b1])
    return abi.Problem(d.kind, d.model, d.n_cams, 0, problem.x[o0:o1], problem.y[o0:o1], problem.u[o0:o1], problem.v[o0:o1],
                       problem.block_offset[b0:b1 + 1] - o0, problem.block_cam[b0:b1], block_b_se3_g=problem.block_b_se3_g[b0:b1],
                       optimize_intrinsics=bool(d.optimize_intrinsics), optimize_skew=bool(d.optimize_skew),
                       optimize_target_pose=bool(d.optimize_target_pose), optimize_hand_eye=bool(d.optimize_hand_eye),
                       huber_delta=d.huber_delta)


def shard_views(problem, x0, rank, world):
    """Shard of a problem with per-view pose blocks (intrinsics / extrinsics kinds): a contiguous range of VIEWS
    (balanced by observation count) with all their residual blocks, the shared parameter blocks replicated.
    Returns (sub-problem, its start vector, (first view, one past the last view)).  The gauge of the extrinsics
    kind fixes GLOBAL view 0 (extrinsics.cpp:133-139): the shard carries its view offset in desc.view_base."""
    from . import abi
    d = problem.desc
    if d.kind == abi.KIND_BUNDLE:
        raise ValueError("the bundle kind has no per-view unknowns: use shard_problem")
    nv, nc = int(d.n_views), int(d.n_cams)
    P = 12 if d.model == abi.MODEL_SCHEIMPFLUG_BC5 else 10
    bview = np.arange(d.n_blocks, dtype=np.int32) if d.kind == abi.KIND_INTRINSICS else np.asarray(problem.block_view)
    blen = np.diff(np.asarray(problem.block_offset))
    per_view = np.bincount(bview, weights=blen, minlength=nv)
    bounds = partition_blocks(np.concatenate([[0], np.cumsum(per_view)]).astype(np.int64), world)
    v0, v1 = int(bounds[rank]), int(bounds[rank + 1])
    sel = np.flatnonzero((bview >= v0) & (bview < v1))
    off = np.asarray(problem.block_offset)
    idx = np.concatenate([np.arange(off[b], off[b + 1]) for b in sel]).astype(np.int64) if len(sel) else np.zeros(0, dtype=np.int64)
    new_off = np.concatenate([[0], np.cumsum(blen[sel])]).astype(np.int64)
    sub = abi.Problem(d.kind, d.model, nc, v1 - v0, problem.x[idx], problem.y[idx], problem.u[idx], problem.v[idx], new_off,
                      np.asarray(problem.block_cam)[sel], block_view=(bview[sel] - v0).astype(np.int32),
                      optimize_intrinsics=bool(d.optimize_intrinsics), optimize_skew=bool(d.optimize_skew),
                      optimize_extrinsics=bool(d.optimize_extrinsics), huber_delta=d.huber_delta)
    sub.desc.view_base = v0
    sub.desc.n_views_total = nv   # the reference's minimum-view check applies to the whole problem, not to a shard
    x0 = np.asarray(x0)
    n_shared = P if d.kind == abi.KIND_INTRINSICS else (P + 7) * nc
    q = x0[n_shared:n_shared + 4 * nv].reshape(nv, 4)[v0:v1]
    t = x0[n_shared + 4 * nv:].reshape(nv, 3)[v0:v1]
    return sub, np.concatenate([x0[:n_shared], q.ravel(), t.ravel()]), (v0, v1)


def gather_views(x_local, x_full_like, problem, view_range):
    """Writes a shard's converged view poses back into a full-size parameter vector (shared part from x_local)."""
    from . import abi
    d = problem.desc
    nv, nc = int(d.n_views), int(d.n_cams)
    P = 12 if d.model == abi.MODEL_SCHEIMPFLUG_BC5 else 10
    n_shared = P if d.kind == abi.KIND_INTRINSICS else (P + 7) * nc
    v0, v1 = view_range
    out = np.array(x_full_like, dtype=np.float64)
    out[:n_shared] = x_local[:n_shared]
    nl = v1 - v0
    out[n_shared + 4 * v0:n_shared + 4 * v1] = x_local[n_shared:n_shared + 4 * nl]
    out[n_shared + 4 * nv + 3 * v0:n_shared + 4 * nv + 3 * v1] = x_local[n_shared + 4 * nl:]
    return out
