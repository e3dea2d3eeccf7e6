"""The reference's own synthetic-recovery tests, re-expressed (same seeds, sizes,
ground truths, perturbations and tolerances).  Each builder returns
(Problem, x0, info) where info carries the ground truth to assert against.

The RNG-dependent part (tests/unit/utils.h RNG + SimulatedHandEye::make_sequence)
is drawn by oracle/refdata.cpp with this toolchain's libstdc++, i.e. the same
stream the reference's tests see when built with GCC.
"""
import numpy as np

import oracle_lib as O
from calibration_b200 import abi
from calibration_b200 import geometry as G


def deg2rad(d):
    return d * np.pi / 180.0


def target_grid(rows, cols, spacing):
    """SimulatedHandEye::make_target_grid (utils.h:223-231)."""
    x0, y0 = -0.5 * (cols - 1) * spacing, -0.5 * (rows - 1) * spacing
    return np.array([[x0 + c * spacing, y0 + r * spacing] for r in range(rows) for c in range(cols)])


def render_views(intr, c_se3_t, obj):
    """render_pixels (utils.h:233-250), zero noise, with its z <= 1e-6 cull."""
    views = []
    for T in c_se3_t:
        P = obj @ T[:3, :2].T + T[:3, 3]
        keep = P[:, 2] > 1e-6
        uv = G.project(intr, P[keep])
        views.append((obj[keep], uv))
    return views


def sim_handeye(seed, n_frames, g_se3_c, b_se3_t, intr, rows, cols, spacing, n_pre=0, n_post=0):
    b_se3_g, pre, post = O.handeye_sequence(seed, n_frames, n_pre, n_post)
    c_se3_t = [G.inv_pose(g_se3_c) @ G.inv_pose(T) @ b_se3_t for T in b_se3_g]
    obj = target_grid(rows, cols, spacing)
    return b_se3_g, c_se3_t, render_views(intr, c_se3_t, obj), pre, post


def views_to_soa(views):
    xs = np.concatenate([v[0][:, 0] for v in views]); ys = np.concatenate([v[0][:, 1] for v in views])
    us = np.concatenate([v[1][:, 0] for v in views]); vs = np.concatenate([v[1][:, 1] for v in views])
    off = np.concatenate([[0], np.cumsum([len(v[0]) for v in views])])
    return xs, ys, us, vs, off


def intrinsics_scenario(skew):
    """OptimizeIntrinsics.RecoversIntrinsicsNoSkew / RecoversSkew (intrinsics_optimize_test.cpp:8-113)."""
    seed = 5 if skew else 7
    intr_gt = np.array([1000.0, 1005.0, 640.0, 360.0, 0.001 if skew else 0.0, 0, 0, 0, 0, 0])
    g = np.eye(4); b = G.make_pose([0.0, 0.0, 2.0])
    _, c_se3_t, views, _, _ = sim_handeye(seed, 15, g, b, intr_gt, 8, 11, 0.02)
    intr0 = intr_gt.copy()
    if skew:
        intr0[0] *= 0.95; intr0[1] *= 1.05; intr0[2] += 10.0; intr0[3] -= 6.0; intr0[4] = 0.0
    else:
        intr0[0] *= 0.97; intr0[1] *= 1.03; intr0[2] += 5.0; intr0[3] -= 4.0
    init = [O.estimate_planar_pose(v[0][:, 0], v[0][:, 1], v[1][:, 0], v[1][:, 1], intr0[:5]) for v in views]
    xs, ys, us, vs, off = views_to_soa(views)
    prob = abi.Problem(abi.KIND_INTRINSICS, abi.MODEL_PINHOLE_BC5, 1, len(views), xs, ys, us, vs, off,
                       np.zeros(len(views), dtype=np.int32), optimize_skew=skew, huber_delta=1.0)
    return prob, G.pack_intrinsics(intr0, init), dict(intr_gt=intr_gt, c_se3_t=c_se3_t)


def bundle_scenario(kind):
    """bundle_test.cpp:9-210.  kind in {'nodist', 'nodist_skew', 'distortion'}."""
    g_gt = G.make_pose([0.03, 0.00, 0.12], [0, 1, 0], deg2rad(8.0))
    if kind == "distortion":
        b_gt = G.make_pose([0.5, -0.1, 80], [1, 0, 0], deg2rad(14.0))
        intr_gt = np.array([900.0, 905.0, 640.0, 360.0, 0.0, -0.12, 0.02, 0.0005, -0.0007, 0.001])
        b_se3_g, _, views, _, _ = sim_handeye(137, 22, g_gt, b_gt, intr_gt, 7, 10, 0.022)
        intr0 = intr_gt.copy(); intr0[5:] = 0.0
        g0 = g_gt.copy(); g0[:3, 3] += [0.01, 0.006, -0.003]
        ax = np.array([0.1, 0.8, 0.1]); g0[:3, :3] = G.angle_axis_to_R(ax / np.linalg.norm(ax), deg2rad(2.0)) @ g0[:3, :3]
        huber, skew = 1.0, False
    else:
        sk = 0.001 if kind == "nodist_skew" else 0.0
        b_gt = G.make_pose([0.5, -0.1, 0.8], [1, 0, 0], deg2rad(14.0))
        intr_gt = np.array([1000.0, 1005.0, 640.0, 360.0, sk, 0, 0, 0, 0, 0])
        b_se3_g, _, views, _, _ = sim_handeye(7, 25, g_gt, b_gt, intr_gt, 8, 11, 0.02)
        intr0 = intr_gt.copy(); intr0[0] *= 0.97; intr0[1] *= 1.03; intr0[2] += 5.0; intr0[3] -= 4.0
        g0 = g_gt.copy(); g0[:3, 3] += [-0.01, 0.006, -0.004]
        ax = np.array([0.3, 0.7, -0.2]); g0[:3, :3] = G.angle_axis_to_R(ax / np.linalg.norm(ax), deg2rad(2.0)) @ g0[:3, :3]
        huber, skew = -1.0, False
    keep = [i for i, v in enumerate(views) if len(v[0]) > 0]
    views = [views[i] for i in keep]; b_se3_g = [b_se3_g[i] for i in keep]
    xs, ys, us, vs, off = views_to_soa(views)
    prob = abi.Problem(abi.KIND_BUNDLE, abi.MODEL_PINHOLE_BC5, 1, 0, xs, ys, us, vs, off,
                       np.zeros(len(views), dtype=np.int32), block_b_se3_g=np.stack([G.pose_to_vec12(T) for T in b_se3_g]),
                       optimize_intrinsics=True, optimize_skew=skew, optimize_target_pose=True, optimize_hand_eye=True,
                       huber_delta=huber)
    return prob, G.pack_bundle([intr0], [g0], b_gt), dict(intr_gt=intr_gt, g_gt=g_gt, b_gt=b_gt)


def circle_poses(n, radius, z0, z_step, rot_step, axis_z=1.0):
    """make_circle_poses (utils.h:81-96)."""
    out = []
    for i in range(n):
        ang = i * 2.0 * np.pi / n
        ax = np.array([np.cos(ang), np.sin(ang), axis_z])
        T = np.eye(4)
        T[:3, 3] = [radius * np.cos(ang), radius * np.sin(ang), z0 + z_step * i]
        T[:3, :3] = G.angle_axis_to_R(ax / np.linalg.norm(ax), rot_step * i)
        out.append(T)
    return out


def scheimpflug_scenario(which):
    """scheimpflug_bundle_test.cpp:13-94.  which in {'intrinsics', 'handeye'}."""
    taux, tauy = 0.02, -0.015
    intr_gt = np.array([100.0, 100.0, 64.0, 48.0, 0.0, 0, 0, 0, 0, 0, taux, tauy])
    g = np.eye(4); g[:3, :3] = G.angle_axis_to_R([0, 1, 0], 0.05); g[:3, 3] = [0.1, 0.0, 0.05]
    b = G.make_pose([0.2, 0.0, 0.0])
    obj = np.array([[-0.1, -0.1], [0.1, -0.1], [0.1, 0.1], [-0.1, 0.1], [0.05, 0.0], [-0.05, 0.0], [0.0, 0.05], [0.0, -0.05]])
    poses = circle_poses(8, 0.1, 0.3, 0.05, 0.1, 0.5)
    views = []
    for btg in poses:  # make_scheimpflug_observations (utils.h:117-137)
        T = G.inv_pose(g) @ G.inv_pose(btg) @ b
        P = obj @ T[:3, :2].T + T[:3, 3]
        views.append((obj, G.project(intr_gt, P)))
    xs, ys, us, vs, off = views_to_soa(views)
    intr0, g0 = intr_gt.copy(), g.copy()
    if which == "intrinsics":
        intr0[10] += 0.01; intr0[11] -= 0.01
        flags = dict(optimize_intrinsics=True, optimize_target_pose=False, optimize_hand_eye=False)
    else:
        g0[:3, 3] += [0.01, -0.01, 0.02]
        flags = dict(optimize_intrinsics=False, optimize_target_pose=False, optimize_hand_eye=True)
    prob = abi.Problem(abi.KIND_BUNDLE, abi.MODEL_SCHEIMPFLUG_BC5, 1, 0, xs, ys, us, vs, off,
                       np.zeros(len(views), dtype=np.int32), block_b_se3_g=np.stack([G.pose_to_vec12(T) for T in poses]),
                       huber_delta=1.0, **flags)
    return prob, G.pack_bundle([intr0], [g0], b), dict(intr_gt=intr_gt, g_gt=g, b_gt=b)


def _rot(axis, ang):
    return G.make_pose([0, 0, 0], axis, ang) if ang != 0 else np.eye(4)


def extrinsics_scenario(which):
    """extrinsics_test.cpp:9-199.  which in {'poses', 'all', 'first_fixed'}.

    'all' / 'first_fixed' start from estimate_extrinsic_dlt in the reference (a
    linear seed outside the hot path); here the seed is the ground truth
    perturbed by a fixed small motion, which the assertions do not depend on.
    """
    intr = np.array([100.0, 100.0, 0.0, 0.0, 0.0, 0, 0, 0, 0, 0])
    cam_gt = [np.eye(4), G.make_pose([1.0, 0.0, 0.0])]
    if which == "poses":
        target_gt = [G.make_pose([0.0, 0.0, 5.0]), G.make_pose([0.5, -0.2, 4.0], [0, 1, 0], 0.3),
                     G.make_pose([-0.3, 0.4, 6.0], [1, 0, 0], -0.2)]
        pts = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
        cam_init = [np.eye(4), G.make_pose([1.2, -0.1, 0.05], [0, 0, 1], 0.05)]
        target_init = [target_gt[0] @ G.make_pose([0.1, 0.0, 0.0]) @ _rot([0, 0, 1], 0.02),
                       target_gt[1] @ G.make_pose([-0.05, 0.1, 0.05]) @ _rot([0, 1, 0], -0.03),
                       target_gt[2] @ G.make_pose([0.02, -0.02, -0.1]) @ _rot([1, 0, 0], 0.01)]
        intr_init = [intr, intr]
        flags = dict(optimize_intrinsics=False)
    else:
        target_gt = [G.make_pose([0.0, 0.0, 5.0]), G.make_pose([0.5, -0.2, 4.0], [0, 1, 0], 0.3)]
        pts = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0], [0.5, 0.5], [-1.0, -1.0], [2.0, 2.0], [2.5, 0.5]])
        intr_init = [np.array([90.0, 95.0, 1.0, -1.0, 0, 0, 0, 0, 0, 0]), np.array([105.0, 98.0, -0.5, 0.5, 0, 0, 0, 0, 0, 0])]
        cam_init = [np.eye(4), cam_gt[1] @ G.make_pose([0.02, -0.01, 0.01], [0, 0, 1], 0.01)]
        target_init = [target_gt[0].copy(), target_gt[1] @ G.make_pose([0.01, 0.02, -0.02], [0, 1, 0], 0.01)]
        if which == "first_fixed":
            target_init[0] = G.make_pose([0.0, 0.0, 3.0])
        flags = dict(optimize_intrinsics=True)
    views, bcam, bview = [], [], []
    for v, Tt in enumerate(target_gt):
        for c, Tc in enumerate(cam_gt):
            T = Tc @ Tt
            P = pts @ T[:3, :2].T + T[:3, 3]
            views.append((pts, G.project(intr, P))); bcam.append(c); bview.append(v)
    xs, ys, us, vs, off = views_to_soa(views)
    prob = abi.Problem(abi.KIND_EXTRINSICS, abi.MODEL_PINHOLE_BC5, 2, len(target_gt), xs, ys, us, vs, off, bcam, bview,
                       optimize_extrinsics=True, huber_delta=1.0, **flags)
    return prob, G.pack_extrinsics(intr_init, cam_init, target_init), dict(cam_gt=cam_gt, target_gt=target_gt, target_init=target_init)
