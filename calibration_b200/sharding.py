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
