"""Host-side pose / camera helpers (numpy) used to pack parameter blocks.

Parameter-block layouts follow the reference exactly (they are also the
covariance row/column order):
  populate_quat_tran / restore_pose   src/estimation/detail/observationutils.h:43-62
  IntrinsicBlocks::get_param_blocks   src/estimation/optim/intrinsics.cpp:34-50
  ExtrinsicBlocks::get_param_blocks   src/estimation/optim/extrinsics.cpp:50-69
  BundleBlocks::get_param_blocks      src/estimation/optim/bundle.cpp:48-68
Poses are 4x4 homogeneous matrices (Eigen::Isometry3d).
"""
import numpy as np


def angle_axis_to_R(axis, angle):
    """Eigen::AngleAxisd(angle, axis).toRotationMatrix() for a unit axis."""
    a = np.asarray(axis, dtype=np.float64)
    s, c = np.sin(angle), np.cos(angle)
    sa, ca = s * a, (1.0 - c) * a
    R = np.empty((3, 3))
    t = ca[0] * a[1]; R[0, 1] = t - sa[2]; R[1, 0] = t + sa[2]
    t = ca[0] * a[2]; R[0, 2] = t + sa[1]; R[2, 0] = t - sa[1]
    t = ca[1] * a[2]; R[1, 2] = t - sa[0]; R[2, 1] = t + sa[0]
    R[0, 0] = ca[0] * a[0] + c; R[1, 1] = ca[1] * a[1] + c; R[2, 2] = ca[2] * a[2] + c
    return R


def make_pose(t, axis=None, angle=0.0, R=None):
    """tests/unit/utils.h:51-57 make_pose (axis is normalised there)."""
    T = np.eye(4)
    if R is not None:
        T[:3, :3] = R
    elif axis is not None and angle >= 1e-16:
        ax = np.asarray(axis, dtype=np.float64)
        T[:3, :3] = angle_axis_to_R(ax / np.linalg.norm(ax), angle)
    T[:3, 3] = t
    return T


def inv_pose(T):
    Ti = np.eye(4)
    Ti[:3, :3] = T[:3, :3].T
    Ti[:3, 3] = -T[:3, :3].T @ T[:3, 3]
    return Ti


def pose_to_vec12(T):
    T = np.asarray(T)
    return np.concatenate([T[..., :3, :3].reshape(T.shape[:-2] + (9,)), T[..., :3, 3]], axis=-1)


def vec12_to_pose(v):
    T = np.eye(4)
    T[:3, :3] = np.asarray(v[:9]).reshape(3, 3)
    T[:3, 3] = v[9:12]
    return T


def rotmat_to_quat(R):
    """Eigen::Quaterniond(Matrix3d) -> (w, x, y, z)."""
    R = np.asarray(R, dtype=np.float64)
    q = np.empty(4)
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0.0:
        t = np.sqrt(t + 1.0); q[0] = 0.5 * t; t = 0.5 / t
        q[1] = (R[2, 1] - R[1, 2]) * t; q[2] = (R[0, 2] - R[2, 0]) * t; q[3] = (R[1, 0] - R[0, 1]) * t
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j = (i + 1) % 3; k = (j + 1) % 3
        t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[1 + i] = 0.5 * t; t = 0.5 / t
        q[0] = (R[k, j] - R[j, k]) * t
        q[1 + j] = (R[j, i] + R[i, j]) * t
        q[1 + k] = (R[k, i] + R[i, k]) * t
    return q


def quat_to_rotmat(q):
    """Eigen toRotationMatrix of a (not re-normalised) quaternion (w, x, y, z)."""
    w, x, y, z = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def pose_to_qt(T):
    return rotmat_to_quat(T[:3, :3]), np.array(T[:3, 3], dtype=np.float64)


def qt_to_pose(q, t):
    """restore_pose: normalises the quaternion first (observationutils.h:50-62)."""
    q = np.asarray(q, dtype=np.float64)
    T = np.eye(4)
    T[:3, :3] = quat_to_rotmat(q / np.linalg.norm(q))
    T[:3, 3] = t
    return T


def rotation_angle(R):
    c = max(-1.0, min(1.0, (np.trace(R) - 1.0) * 0.5))
    return float(np.arccos(c))


# ---------------------------------------------------------------------------
# camera models (vectorised restatement used only to SYNTHESISE pixels)
# ---------------------------------------------------------------------------
def project(intr, P):
    """Pinhole + Brown-Conrady (10 params) or Scheimpflug (12 params); P: (..., 3)."""
    intr = np.asarray(intr, dtype=np.float64)
    P = np.asarray(P, dtype=np.float64)
    fx, fy, cx, cy, sk, k1, k2, k3, p1, p2 = intr[:10]
    if len(intr) == 12:
        tx, ty = intr[10], intr[11]
        ctx, stx, cty, sty = np.cos(tx), np.sin(tx), np.cos(ty), np.sin(ty)
        a = np.array([cty, 0.0, -sty]); b = np.array([stx * sty, ctx, stx * cty]); n = np.array([ctx * sty, -stx, ctx * cty])
        sden = P @ n
        mx0, my0 = a[2] / n[2], b[2] / n[2]
        x = (P @ a) / sden - mx0
        y = (P @ b) / sden - my0
        shift_u, shift_v = fx * mx0 + sk * my0, fy * my0
    else:
        x = P[..., 0] / P[..., 2]
        y = P[..., 1] / P[..., 2]
        shift_u = shift_v = 0.0
    r2 = x * x + y * y
    radial = 1.0 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
    xd = x * radial + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
    yd = y * radial + p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y
    return np.stack([fx * xd + sk * yd + cx + shift_u, fy * yd + cy + shift_v], axis=-1)


# ---------------------------------------------------------------------------
# parameter packing
# ---------------------------------------------------------------------------
def pack_intrinsics(intr, c_se3_t):
    qs, ts = zip(*[pose_to_qt(T) for T in c_se3_t])
    return np.concatenate([np.asarray(intr, dtype=np.float64), np.concatenate(qs), np.concatenate(ts)])


def unpack_intrinsics(x, n_views, P=10):
    intr = np.array(x[:P])
    q = np.asarray(x[P:P + 4 * n_views]).reshape(n_views, 4)
    t = np.asarray(x[P + 4 * n_views:P + 7 * n_views]).reshape(n_views, 3)
    return intr, [qt_to_pose(q[i], t[i]) for i in range(n_views)]


def pack_extrinsics(intrs, c_se3_r, r_se3_t):
    cq, ct = zip(*[pose_to_qt(T) for T in c_se3_r])
    vq, vt = zip(*[pose_to_qt(T) for T in r_se3_t])
    return np.concatenate([np.concatenate([np.asarray(i, dtype=np.float64) for i in intrs]), np.concatenate(cq),
                           np.concatenate(ct), np.concatenate(vq), np.concatenate(vt)])


def unpack_extrinsics(x, n_cams, n_views, P=10):
    o = 0
    intrs = np.asarray(x[o:o + P * n_cams]).reshape(n_cams, P).copy(); o += P * n_cams
    cq = np.asarray(x[o:o + 4 * n_cams]).reshape(n_cams, 4); o += 4 * n_cams
    ct = np.asarray(x[o:o + 3 * n_cams]).reshape(n_cams, 3); o += 3 * n_cams
    vq = np.asarray(x[o:o + 4 * n_views]).reshape(n_views, 4); o += 4 * n_views
    vt = np.asarray(x[o:o + 3 * n_views]).reshape(n_views, 3)
    return (intrs, [qt_to_pose(cq[i], ct[i]) for i in range(n_cams)],
            [qt_to_pose(vq[i], vt[i]) for i in range(n_views)])


def pack_bundle(intrs, g_se3_c, b_se3_t):
    gq, gt = zip(*[pose_to_qt(T) for T in g_se3_c])
    bq, bt = pose_to_qt(b_se3_t)
    return np.concatenate([np.concatenate([np.asarray(i, dtype=np.float64) for i in intrs]), np.concatenate(gq),
                           np.concatenate(gt), bq, bt])


def unpack_bundle(x, n_cams, P=10):
    o = 0
    intrs = np.asarray(x[o:o + P * n_cams]).reshape(n_cams, P).copy(); o += P * n_cams
    gq = np.asarray(x[o:o + 4 * n_cams]).reshape(n_cams, 4); o += 4 * n_cams
    gt = np.asarray(x[o:o + 3 * n_cams]).reshape(n_cams, 3); o += 3 * n_cams
    return intrs, [qt_to_pose(gq[i], gt[i]) for i in range(n_cams)], qt_to_pose(x[o:o + 4], x[o + 4:o + 7])


def pack_handeye(g_se3_c):
    q, t = pose_to_qt(g_se3_c)
    return np.concatenate([q, t])
