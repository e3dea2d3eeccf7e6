"""Seeded synthetic checkerboard data for the BASELINE.json configurations.

One host generator produces the bytes both the CUDA path and the CPU oracle
consume.  Shapes:
  C1  optimize_intrinsics, 1 camera x 20 views x 9x6 corners
  C3  optimize_extrinsics, 2 cameras x 1000 views x 11x8 corners
  C4  optimize_bundle, 4 cameras x 5000 robot poses x 11x8 corners
  C5  optimize_bundle, 8 cameras x 100000 views x 11x8 corners (~70.4 M observations)
(C2, batched RANSAC, lives in synth_ransac below.)
Everything is vectorised numpy; C5 is generated in chunks of views.
"""
import numpy as np

from . import abi
from . import geometry as G

K_GT = np.array([1000.0, 1005.0, 640.0, 360.0, 0.0])
DIST_GT = np.array([-0.12, 0.02, 0.0005, -0.0007, 0.001])  # k1, k2, k3, p1, p2 (bundle_test.cpp:165)


def grid(rows, cols, spacing):
    x0, y0 = -0.5 * (cols - 1) * spacing, -0.5 * (rows - 1) * spacing
    c, r = np.meshgrid(np.arange(cols), np.arange(rows))
    return np.stack([x0 + c.ravel() * spacing, y0 + r.ravel() * spacing], axis=1)


def rodrigues(rvec):
    """(n, 3) rotation vectors -> (n, 3, 3)."""
    rvec = np.atleast_2d(rvec)
    th = np.linalg.norm(rvec, axis=1)
    k = rvec / np.maximum(th, 1e-300)[:, None]
    K = np.zeros((len(rvec), 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -k[:, 2], k[:, 1], k[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -k[:, 0], -k[:, 1], k[:, 0]
    s, c = np.sin(th)[:, None, None], np.cos(th)[:, None, None]
    return np.eye(3)[None] + s * K + (1 - c) * (K @ K)


def random_poses(rng, n, max_tilt_deg, z_range, xy_range):
    """Target poses in front of a camera: (n, 4, 4)."""
    ax = rng.normal(size=(n, 3)); ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = np.deg2rad(rng.uniform(0.0, max_tilt_deg, size=n))
    T = np.tile(np.eye(4), (n, 1, 1))
    T[:, :3, :3] = rodrigues(ax * ang[:, None])
    T[:, 0, 3] = rng.uniform(-xy_range, xy_range, size=n)
    T[:, 1, 3] = rng.uniform(-xy_range, xy_range, size=n)
    T[:, 2, 3] = rng.uniform(*z_range, size=n)
    return T


def inv_poses(T):
    Ti = np.tile(np.eye(4), (len(T), 1, 1))
    Rt = np.transpose(T[:, :3, :3], (0, 2, 1))
    Ti[:, :3, :3] = Rt
    Ti[:, :3, 3] = -(Rt @ T[:, :3, 3:4])[:, :, 0]
    return Ti


def perturb_pose(rng, T, rot_deg, trans):
    ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
    D = np.eye(4); D[:3, :3] = G.angle_axis_to_R(ax, np.deg2rad(rot_deg)); D[:3, 3] = rng.uniform(-trans, trans, size=3)
    return T @ D


def render(intr, c_se3_t, obj, rng=None, noise=0.0):
    """Pixels of the board under poses (n, 4, 4): returns (n, m, 2)."""
    P = np.einsum("nij,mj->nmi", c_se3_t[:, :3, :2], obj) + c_se3_t[:, None, :3, 3]
    uv = G.project(intr, P)
    if noise > 0:
        uv = uv + rng.normal(scale=noise, size=uv.shape)
    return uv


def _soa(obj, uv):
    n, m = uv.shape[:2]
    xs = np.tile(obj[:, 0], n); ys = np.tile(obj[:, 1], n)
    return xs, ys, uv[..., 0].ravel(), uv[..., 1].ravel(), np.arange(n + 1, dtype=np.int64) * m


def make_intrinsics(seed=7, n_views=20, rows=6, cols=9, spacing=0.03, noise=0.2, huber_delta=1.0, optimize_skew=False,
                    model=abi.MODEL_PINHOLE_BC5):
    """C1: single-camera planar intrinsic calibration."""
    rng = np.random.default_rng(seed)
    intr_gt = np.concatenate([K_GT, DIST_GT])
    if model == abi.MODEL_SCHEIMPFLUG_BC5:
        intr_gt = np.concatenate([intr_gt, [0.02, -0.015]])
    obj = grid(rows, cols, spacing)
    poses = random_poses(rng, n_views, 35.0, (0.5, 0.9), 0.08)
    uv = render(intr_gt, poses, obj, rng, noise)
    xs, ys, us, vs, off = _soa(obj, uv)
    prob = abi.Problem(abi.KIND_INTRINSICS, model, 1, n_views, xs, ys, us, vs, off, np.zeros(n_views, dtype=np.int32),
                       optimize_skew=optimize_skew, huber_delta=huber_delta)
    intr0 = intr_gt.copy(); intr0[0] *= 0.97; intr0[1] *= 1.03; intr0[2] += 5.0; intr0[3] -= 4.0; intr0[5:10] = 0.0
    init = [perturb_pose(rng, poses[i], 1.0, 0.005) for i in range(n_views)]
    return prob, G.pack_intrinsics(intr0, init), G.pack_intrinsics(intr_gt, list(poses))


def make_extrinsics(seed=5, n_cams=2, n_views=1000, rows=8, cols=11, spacing=0.02, noise=0.2, huber_delta=1.0,
                    optimize_intrinsics=True, optimize_extrinsics=True, optimize_skew=False, drop_fraction=0.0):
    """C3: stereo / multi-camera extrinsic refinement (joint intrinsics + relative pose)."""
    rng = np.random.default_rng(seed)
    obj = grid(rows, cols, spacing)
    intr_gt, cam_gt = [], []
    for c in range(n_cams):
        k = np.concatenate([K_GT, DIST_GT]); k[0] *= 1.0 + 0.01 * c; k[1] *= 1.0 + 0.01 * c
        intr_gt.append(k)
        T = np.eye(4)
        if c > 0:
            T = G.make_pose([0.12 * c, 0.0, 0.0], [0, 1, 0], np.deg2rad(2.0 * c))
        cam_gt.append(T)
    tgt = random_poses(rng, n_views, 30.0, (0.8, 1.5), 0.15)
    xs, ys, us, vs, bcam, bview, lens = [], [], [], [], [], [], []
    for c in range(n_cams):
        uv = render(intr_gt[c], cam_gt[c][None] @ tgt, obj, rng, noise)
        keep = rng.uniform(size=n_views) >= drop_fraction if (drop_fraction > 0 and c > 0) else np.ones(n_views, bool)
        for v in np.nonzero(keep)[0]:
            xs.append(obj[:, 0]); ys.append(obj[:, 1]); us.append(uv[v, :, 0]); vs.append(uv[v, :, 1])
            bcam.append(c); bview.append(v); lens.append(len(obj))
    order = np.lexsort((bcam, bview))  # reference order: view-major, camera inner (extrinsics.cpp:91-106)
    cat = lambda a: np.concatenate([a[i] for i in order])
    off = np.concatenate([[0], np.cumsum([lens[i] for i in order])])
    prob = abi.Problem(abi.KIND_EXTRINSICS, abi.MODEL_PINHOLE_BC5, n_cams, n_views, cat(xs), cat(ys), cat(us), cat(vs), off,
                       np.asarray(bcam)[order], np.asarray(bview)[order], optimize_intrinsics=optimize_intrinsics,
                       optimize_extrinsics=optimize_extrinsics, optimize_skew=optimize_skew, huber_delta=huber_delta)
    intr0 = [k.copy() for k in intr_gt]
    if optimize_intrinsics:
        for k in intr0:
            k[0] *= 0.99; k[1] *= 1.01; k[2] += 2.0; k[3] -= 1.5
    cam0 = [cam_gt[0]] + [perturb_pose(rng, T, 2.0, 0.01) for T in cam_gt[1:]]
    tgt0 = [tgt[0]] + [perturb_pose(rng, tgt[v], 2.0, 0.01) for v in range(1, n_views)]
    return prob, G.pack_extrinsics(intr0, cam0, tgt0), G.pack_extrinsics(intr_gt, cam_gt, list(tgt))


def rig(n_cams):
    """Hand-eye poses g_se3_c of a small-baseline rig (bundle_test.cpp:13-16 for camera 0)."""
    out = []
    for k in range(n_cams):
        T = G.make_pose([0.03 + 0.02 * (k % 4), 0.015 * (k // 4), 0.12], [0, 1, 0], np.deg2rad(8.0))
        roll = np.eye(4); roll[:3, :3] = G.angle_axis_to_R([0, 0, 1], np.deg2rad(90.0 * (k % 4)))
        out.append(T @ roll)
    return out


def make_bundle(seed=137, n_cams=4, n_poses=5000, rows=8, cols=11, spacing=0.02, noise=0.2, huber_delta=1.0,
                optimize_intrinsics=True, optimize_skew=False, optimize_target_pose=True, optimize_hand_eye=True,
                model=abi.MODEL_PINHOLE_BC5, chunk=12500, chunks=None, pinned=False):
    """C4 / C5: hand-eye bundle (one block per (robot pose, camera), blocks in pose-major order).

    Poses are generated in chunks of `chunk` robot poses, chunk c from its own
    generator seeded by (seed, c), so a shard (`chunks` = iterable of chunk ids)
    holds exactly the bytes the full problem holds for those poses; the
    parameters (ground truth and start) depend on `seed` only.
    """
    prng = np.random.default_rng(seed)
    obj = grid(rows, cols, spacing)
    m = len(obj)
    g_gt = rig(n_cams)
    b_gt = G.make_pose([0.5, -0.1, 0.8], [1, 0, 0], np.deg2rad(14.0))
    intr_gt = []
    for c in range(n_cams):
        k = np.concatenate([K_GT, DIST_GT]); k[0] *= 1.0 + 0.005 * c; k[1] *= 1.0 + 0.005 * c
        if model == abi.MODEL_SCHEIMPFLUG_BC5:
            k = np.concatenate([k, [0.02 - 0.004 * c, -0.015 + 0.003 * c]])
        intr_gt.append(k)
    intr0 = [k.copy() for k in intr_gt]
    if optimize_intrinsics:
        for k in intr0:
            k[0] *= 0.99; k[1] *= 1.01; k[2] += 2.0; k[3] -= 1.5
            if model == abi.MODEL_SCHEIMPFLUG_BC5:
                k[10] += 0.005; k[11] -= 0.005
    g0 = [perturb_pose(prng, T, 1.0, 0.005) for T in g_gt] if optimize_hand_eye else g_gt
    b0 = perturb_pose(prng, b_gt, 1.0, 0.005) if optimize_target_pose else b_gt

    n_chunks = (n_poses + chunk - 1) // chunk
    chunk_ids = list(range(n_chunks)) if chunks is None else list(chunks)
    sizes = [min(chunk, n_poses - c * chunk) for c in chunk_ids]
    n_local = int(sum(sizes))
    nb = n_cams * n_local
    n_obs = nb * m
    if pinned:
        import torch
        keep = [torch.empty(n_obs, dtype=torch.float64).pin_memory() for _ in range(4)]
        bufs = [t.numpy() for t in keep]
    else:
        keep = None
        bufs = [np.empty(n_obs) for _ in range(4)]
    xs, ys, us, vs = bufs
    bTg = np.empty((nb, 12))
    g_inv = np.stack([G.inv_pose(T) for T in g_gt])
    p0 = 0
    for cid, n in zip(chunk_ids, sizes):
        rng = np.random.default_rng([seed, cid])
        # camera 0 sees the board under a random pose; the robot pose follows from the chain
        c0_t = random_poses(rng, n, 30.0, (0.8, 1.5), 0.12)
        b_g = b_gt[None] @ inv_poses(c0_t) @ g_inv[0][None]       # b_se3_g = b_se3_t (c_se3_t)^-1 (g_se3_c)^-1
        g_b = inv_poses(b_g)
        sl = slice(p0 * n_cams * m, (p0 + n) * n_cams * m)
        X = xs[sl].reshape(n, n_cams, m); Y = ys[sl].reshape(n, n_cams, m)
        U = us[sl].reshape(n, n_cams, m); V = vs[sl].reshape(n, n_cams, m)
        X[:] = obj[:, 0]; Y[:] = obj[:, 1]
        for c in range(n_cams):
            c_t = g_inv[c][None] @ g_b @ b_gt[None]
            uv = render(intr_gt[c], c_t, obj, rng, noise)
            U[:, c, :] = uv[..., 0]; V[:, c, :] = uv[..., 1]
        bTg[p0 * n_cams:(p0 + n) * n_cams] = np.repeat(G.pose_to_vec12(b_g), n_cams, axis=0)
        p0 += n
    bcam = np.tile(np.arange(n_cams, dtype=np.int32), n_local)
    off = np.arange(nb + 1, dtype=np.int64) * m
    prob = abi.Problem(abi.KIND_BUNDLE, model, n_cams, 0, xs, ys, us, vs, off, bcam, block_b_se3_g=bTg,
                       optimize_intrinsics=optimize_intrinsics, optimize_skew=optimize_skew,
                       optimize_target_pose=optimize_target_pose, optimize_hand_eye=optimize_hand_eye, huber_delta=huber_delta)
    prob._pinned_keep = keep
    return prob, G.pack_bundle(intr0, g0, b0), G.pack_bundle(intr_gt, g_gt, b_gt)


def make_handeye_poses(seed=2024, n=60):
    """Pose lists for optimize_handeye (AX = XB): returns (b_se3_g list, c_se3_t list, X_gt)."""
    rng = np.random.default_rng(seed)
    X_gt = G.make_pose([0.02, -0.01, 0.09], rng.normal(size=3), np.deg2rad(10.0))
    b_t = G.make_pose([0.25, 0.05, 0.55], rng.normal(size=3), np.deg2rad(18.0))
    T = np.eye(4); bg, ct = [], []
    for _ in range(n):
        bg.append(T.copy()); ct.append(G.inv_pose(X_gt) @ G.inv_pose(T) @ b_t)
        d = G.make_pose(rng.uniform(-0.1, 0.1, size=3), rng.normal(size=3), np.deg2rad(rng.uniform(5.0, 25.0)))
        T = T @ d
    return bg, ct, X_gt


def synth_ransac(seed=17, n_problems=1000, n=500, outlier_fraction=0.3, noise=0.3):
    """C2: n_problems image pairs x n correspondences, arrays of shape (n_problems, n)."""
    rng = np.random.default_rng(seed)
    poses = random_poses(rng, n_problems, 35.0, (1.5, 3.0), 0.3)
    K = np.array([[1000.0, 0, 640.0], [0, 1005.0, 360.0], [0, 0, 1.0]])
    H = K[None] @ np.stack([poses[:, :3, 0], poses[:, :3, 1], poses[:, :3, 3]], axis=2)
    x = rng.uniform(-0.2, 0.2, size=(n_problems, n)); y = rng.uniform(-0.2, 0.2, size=(n_problems, n))
    q = np.einsum("pij,pnj->pni", H, np.stack([x, y, np.ones_like(x)], axis=2))
    u = q[..., 0] / q[..., 2] + rng.normal(scale=noise, size=x.shape)
    v = q[..., 1] / q[..., 2] + rng.normal(scale=noise, size=x.shape)
    out = rng.uniform(size=x.shape) < outlier_fraction
    u = np.where(out, rng.uniform(0, 1280, size=x.shape), u)
    v = np.where(out, rng.uniform(0, 720, size=x.shape), v)
    return x, y, u, v, H


def synth_plane_ransac(seed=23, n_problems=1000, n=500, outlier_fraction=0.3, noise=0.002):
    """Laser-plane style problems (fit_plane_ransac, linear/planefit.cpp:86-104): n points per problem near a
    random plane (unit normal within 40 deg of +z, offset 0.5..2 m), Gaussian distance noise, a fraction of
    gross outliers in the surrounding box.  Arrays of shape (n_problems, n) and the planes (n_problems, 4)."""
    rng = np.random.default_rng(seed)
    tilt = np.deg2rad(rng.uniform(0, 40, n_problems)); az = rng.uniform(0, 2 * np.pi, n_problems)
    nrm = np.stack([np.sin(tilt) * np.cos(az), np.sin(tilt) * np.sin(az), np.cos(tilt)], axis=1)
    d = -rng.uniform(0.5, 2.0, n_problems)
    x = rng.uniform(-0.5, 0.5, size=(n_problems, n)); y = rng.uniform(-0.5, 0.5, size=(n_problems, n))
    z = (-d[:, None] - nrm[:, :1] * x - nrm[:, 1:2] * y) / nrm[:, 2:3]
    e = rng.normal(scale=noise, size=x.shape)
    x = x + e * nrm[:, :1]; y = y + e * nrm[:, 1:2]; z = z + e * nrm[:, 2:3]
    out = rng.uniform(size=x.shape) < outlier_fraction
    z = np.where(out, rng.uniform(0.0, 3.0, size=x.shape), z)
    return x, y, z, np.concatenate([nrm, d[:, None]], axis=1)
