"""ctypes loader of the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

from calibration_b200 import abi

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "liboracle.so")


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(_ROOT, "oracle")], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_ROOT, "oracle", f) for f in os.listdir(os.path.join(_ROOT, "oracle"))
                if f.endswith((".cpp", ".hpp", ".h")) or f == "Makefile"]
        if not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
            build()
        L = C.CDLL(_SO)
        dp, ip, i64 = abi.c_double_p, abi.c_int32_p, C.c_int64
        L.orc_param_count.restype = i64
        L.orc_tangent_count.restype = i64
        L.orc_param_count.argtypes = [C.POINTER(abi.ProblemDesc)]
        L.orc_tangent_count.argtypes = [C.POINTER(abi.ProblemDesc)]
        L.orc_refine_eval.argtypes = [C.POINTER(abi.ProblemDesc), dp, dp, dp, dp, C.c_int]
        L.orc_refine_solve.argtypes = [C.POINTER(abi.ProblemDesc), C.POINTER(abi.OptimOptions), dp,
                                       C.POINTER(abi.OptimResult), dp, C.c_int]
        L.orc_block_ssr.argtypes = [C.POINTER(abi.ProblemDesc), dp, dp, C.c_int]
        L.orc_analytic_bundle_eval.argtypes = [C.POINTER(abi.ProblemDesc), dp, dp, dp, dp, C.c_int]
        L.orc_project.argtypes = [C.c_int, dp, dp, dp]
        L.orc_axxb_eval.argtypes = [C.POINTER(abi.AxxbDesc), dp, dp, dp, dp, C.c_int]
        L.orc_axxb_solve.argtypes = [C.POINTER(abi.AxxbDesc), C.POINTER(abi.OptimOptions), dp,
                                     C.POINTER(abi.OptimResult), dp]
        L.orc_build_all_pairs.restype = i64
        L.orc_build_all_pairs.argtypes = [i64, dp, dp, C.c_double, dp, dp, dp, dp]
        L.orc_sample_stream.argtypes = [C.c_uint64, C.c_int32, C.c_int32, ip]
        L.orc_sample_stream_libstdcxx.argtypes = [C.c_uint64, C.c_int32, C.c_int32, ip]
        L.orc_ransac_homography.argtypes = [C.c_int32, dp, dp, dp, dp, C.POINTER(abi.RansacOptions), ip,
                                            C.POINTER(abi.RansacResult), abi.c_uint8_p]
        L.orc_ransac_homography_batch.argtypes = [i64, C.c_int32, dp, dp, dp, dp, C.POINTER(abi.RansacOptions),
                                                  C.c_int, C.POINTER(abi.RansacResult), abi.c_uint8_p, C.c_int]
        L.orc_homography_dlt.argtypes = [C.c_int32, dp, dp, dp, dp, dp]
        L.orc_ref_handeye_sequence.argtypes = [C.c_uint32, C.c_int, dp, C.c_int, dp, C.c_int, dp]
        L.orc_ref_gauss_stream.argtypes = [C.c_uint32, C.c_int, C.c_double, i64, dp]
        L.orc_ref_homography_data.argtypes = [C.c_int, C.c_double, C.c_int, C.c_uint32, dp, dp]
        L.orc_ref_estimate_planar_pose.argtypes = [C.c_int32, dp, dp, dp, dp, dp, dp]
        L.orc_zhang_rows.argtypes = [dp, dp]
        L.orc_zhang_intrinsics.argtypes = [i64, dp, dp]
        L.orc_sanitize_intrinsics.argtypes = [dp, dp]
        L.orc_pose_from_homography.argtypes = [dp, dp, dp, dp, dp]
        L.orc_estimate_intrinsics.argtypes = [i64, abi.c_int64_p, dp, dp, dp, dp, dp, dp, ip, dp, dp, dp]
        L.orc_estimate_intrinsics_ransac.argtypes = [i64, abi.c_int64_p, dp, dp, dp, dp, dp, C.POINTER(abi.RansacOptions), dp, ip, dp, dp, dp,
                                                     abi.c_uint8_p]
        L.orc_project_to_so3.argtypes = [dp, dp]
        L.orc_log_so3.argtypes = [dp, dp]
        L.orc_sample_stream_k.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_int32, ip]
        L.orc_sample_stream_k_libstdcxx.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_int32, ip]
        L.orc_ref_plane_data.argtypes = [dp, dp]
        L.orc_fit_plane_svd.argtypes = [C.c_int32, dp, dp, dp, dp]
        L.orc_ransac_plane.argtypes = [C.c_int32, dp, dp, dp, C.POINTER(abi.RansacOptions), C.POINTER(abi.PlaneResult), abi.c_uint8_p]
        L.orc_ransac_plane_batch.argtypes = [i64, C.c_int32, dp, dp, dp, C.POINTER(abi.RansacOptions), C.c_int,
                                             C.POINTER(abi.PlaneResult), abi.c_uint8_p, C.c_int]
        _lib = L
    return _lib


def refine_eval(prob, x, jac=True, threads=0):
    L = lib()
    x = abi.as_f64(x)
    n = int(L.orc_tangent_count(C.byref(prob.desc)))
    cost = C.c_double()
    g = np.zeros(n) if jac else None
    H = np.zeros((n, n)) if jac else None
    rc = L.orc_refine_eval(C.byref(prob.desc), abi.dptr(x), C.cast(C.byref(cost), abi.c_double_p), abi.dptr(g),
                           abi.dptr(H), threads)
    assert rc == 0
    return cost.value, g, H


def refine_solve(prob, x0, opts=None, force_dense=False, want_cov=True):
    L = lib()
    x = abi.as_f64(x0).copy()
    assert len(x) == int(L.orc_param_count(C.byref(prob.desc))), (len(x), L.orc_param_count(C.byref(prob.desc)))
    opts = opts or abi.OptimOptions.default()
    res = abi.OptimResult()
    cov = np.zeros((len(x), len(x))) if (want_cov and opts.compute_covariance) else None
    L.orc_refine_solve(C.byref(prob.desc), C.byref(opts), abi.dptr(x), C.byref(res), abi.dptr(cov), int(force_dense))
    return x, res, cov


def analytic_bundle_eval(prob, x, jac=True, threads=0):
    """The conservative CPU baseline (oracle/analytic_pass.cpp): hand-derived Jacobians, OpenMP over residual blocks;
    returns (cost, g, H) over the shared tangent blocks like refine_eval does for the bundle kind."""
    L = lib()
    n = int(L.orc_tangent_count(C.byref(prob.desc)))
    cost = C.c_double(); g = np.zeros(n) if jac else None; H = np.zeros((n, n)) if jac else None
    rc = L.orc_analytic_bundle_eval(C.byref(prob.desc), abi.dptr(abi.as_f64(x)), C.cast(C.byref(cost), abi.c_double_p), abi.dptr(g), abi.dptr(H), threads)
    assert rc == 0
    return cost.value, g, H


def block_ssr(prob, x, threads=0):
    L = lib()
    out = np.zeros(prob.desc.n_blocks)
    L.orc_block_ssr(C.byref(prob.desc), abi.dptr(abi.as_f64(x)), abi.dptr(out), threads)
    return out


def project(model, intr, P):
    L = lib()
    uv = np.zeros(2)
    L.orc_project(model, abi.dptr(abi.as_f64(intr)), abi.dptr(abi.as_f64(P)), abi.dptr(uv))
    return uv


def handeye_sequence(seed, n_frames, n_pre=0, n_post=0):
    L = lib()
    pre = np.zeros((max(n_pre, 1), 3)); post = np.zeros((max(n_post, 1), 3)); bg = np.zeros((n_frames, 12))
    L.orc_ref_handeye_sequence(seed, n_pre, abi.dptr(pre), n_frames, abi.dptr(bg), n_post, abi.dptr(post))
    poses = []
    for k in range(n_frames):
        T = np.eye(4); T[:3, :3] = bg[k, :9].reshape(3, 3); T[:3, 3] = bg[k, 9:]
        poses.append(T)
    return poses, pre[:n_pre], post[:n_post]


def estimate_planar_pose(x, y, u, v, K5):
    L = lib()
    out = np.zeros(12)
    x, y, u, v = (abi.as_f64(a) for a in (x, y, u, v))
    L.orc_ref_estimate_planar_pose(len(x), abi.dptr(x), abi.dptr(y), abi.dptr(u), abi.dptr(v),
                                   abi.dptr(abi.as_f64(K5)), abi.dptr(out))
    T = np.eye(4); T[:3, :3] = out[:9].reshape(3, 3); T[:3, 3] = out[9:]
    return T


def sample_stream(seed, n, iters, real=False):
    L = lib()
    out = np.zeros((iters, 4), dtype=np.int32)
    fn = L.orc_sample_stream_libstdcxx if real else L.orc_sample_stream
    fn(seed, n, iters, out.ctypes.data_as(abi.c_int32_p))
    return out


def ransac(x, y, u, v, opts=None, sample_idx=None):
    L = lib()
    x, y, u, v = (abi.as_f64(a) for a in (x, y, u, v))
    opts = opts or abi.RansacOptions.default()
    res = abi.RansacResult()
    mask = np.zeros(len(x), dtype=np.uint8)
    sp = None
    if sample_idx is not None:
        sample_idx = np.ascontiguousarray(sample_idx, dtype=np.int32)
        sp = sample_idx.ctypes.data_as(abi.c_int32_p)
    L.orc_ransac_homography(len(x), abi.dptr(x), abi.dptr(y), abi.dptr(u), abi.dptr(v), C.byref(opts), sp,
                            C.byref(res), mask.ctypes.data_as(abi.c_uint8_p))
    return res, mask


def ransac_batch(x, y, u, v, opts=None, seed_per_problem=True, threads=0):
    """x, y, u, v: (n_problems, n) arrays."""
    L = lib()
    x, y, u, v = (abi.as_f64(a) for a in (x, y, u, v))
    npb, n = x.shape
    opts = opts or abi.RansacOptions.default()
    res = (abi.RansacResult * npb)()
    mask = np.zeros((npb, n), dtype=np.uint8)
    L.orc_ransac_homography_batch(npb, n, abi.dptr(x), abi.dptr(y), abi.dptr(u), abi.dptr(v), C.byref(opts),
                                  int(seed_per_problem), res, mask.ctypes.data_as(abi.c_uint8_p), threads)
    return res, mask


_REF_SO = os.path.join(_ROOT, "oracle", "_ref", "libref_ransac.so")
_REF_TREE = "/root/reference"
_ref = None


def ref_lib():
    """oracle/_ref/libref_ransac.so: the reference's own calib::ransac<> template (common/ransac.h:121-194)
    compiled from where it lies in /root/reference, around the oracle's estimator hooks.  Built here when the
    reference tree is present; on the GPU box only the prebuilt file is used.  None when neither exists."""
    global _ref
    if _ref is None:
        if os.path.isdir(_REF_TREE):
            subprocess.run(["make", "-s", "-C", os.path.join(_ROOT, "oracle"), "ref"], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        if not os.path.exists(_REF_SO):
            return None
        L = C.CDLL(_REF_SO)
        dp = abi.c_double_p
        L.ref_ransac_homography.argtypes = [C.c_int32, dp, dp, dp, dp, C.POINTER(abi.RansacOptions),
                                            C.POINTER(abi.RansacResult), abi.c_uint8_p]
        L.ref_ransac_plane.argtypes = [C.c_int32, dp, dp, dp, C.POINTER(abi.RansacOptions), C.POINTER(abi.PlaneResult),
                                       abi.c_uint8_p]
        L.orc_ransac_plane.argtypes = L.ref_ransac_plane.argtypes
        # the harness compiles oracle/ransac.cpp into the same library, with the same flags
        L.orc_ransac_homography.argtypes = [C.c_int32, dp, dp, dp, dp, C.POINTER(abi.RansacOptions), abi.c_int32_p,
                                            C.POINTER(abi.RansacResult), abi.c_uint8_p]
        _ref = L
    return _ref


def ref_ransac(x, y, u, v, opts=None, oracle_twin=False):
    """The reference's RANSAC loop itself (see ref_lib): (result, inlier mask).  oracle_twin=True runs the
    oracle's ransac_one as compiled INTO the same library (same flags, no FMA contraction) instead."""
    L = ref_lib()
    x, y, u, v = (abi.as_f64(a) for a in (x, y, u, v))
    opts = opts or abi.RansacOptions.default()
    res = abi.RansacResult()
    mask = np.zeros(len(x), dtype=np.uint8)
    args = (len(x), abi.dptr(x), abi.dptr(y), abi.dptr(u), abi.dptr(v), C.byref(opts))
    if oracle_twin:
        L.orc_ransac_homography(*args, None, C.byref(res), mask.ctypes.data_as(abi.c_uint8_p))
    else:
        L.ref_ransac_homography(*args, C.byref(res), mask.ctypes.data_as(abi.c_uint8_p))
    return res, mask


def ref_ransac_plane(x, y, z, opts=None, oracle_twin=False):
    """fit_plane_ransac through the reference's loop (see ref_lib); oracle_twin as in ref_ransac."""
    L = ref_lib()
    x, y, z = (abi.as_f64(a) for a in (x, y, z))
    opts = opts or abi.RansacOptions.default()
    res = abi.PlaneResult()
    mask = np.zeros(len(x), dtype=np.uint8)
    fn = L.orc_ransac_plane if oracle_twin else L.ref_ransac_plane
    fn(len(x), abi.dptr(x), abi.dptr(y), abi.dptr(z), C.byref(opts), C.byref(res), mask.ctypes.data_as(abi.c_uint8_p))
    return res, mask


def sample_stream_k(seed, n, k, iters, real=False):
    out = np.zeros((iters, k), dtype=np.int32)
    fn = lib().orc_sample_stream_k_libstdcxx if real else lib().orc_sample_stream_k
    fn(seed, n, k, iters, out.ctypes.data_as(abi.c_int32_p))
    return out


def plane_testdata():
    """planefit_test.cpp:24-46: (ground-truth plane, 140 x 3 points)."""
    plane = np.zeros(4); xyz = np.zeros((140, 3))
    lib().orc_ref_plane_data(abi.dptr(plane), abi.dptr(xyz))
    return plane, xyz


def fit_plane_svd(x, y, z):
    x, y, z = (abi.as_f64(a) for a in (x, y, z))
    plane = np.zeros(4)
    rc = lib().orc_fit_plane_svd(len(x), abi.dptr(x), abi.dptr(y), abi.dptr(z), abi.dptr(plane))
    return rc, plane


def ransac_plane(x, y, z, opts=None):
    x, y, z = (abi.as_f64(a) for a in (x, y, z))
    opts = opts or abi.RansacOptions.default()
    res = abi.PlaneResult()
    mask = np.zeros(len(x), dtype=np.uint8)
    lib().orc_ransac_plane(len(x), abi.dptr(x), abi.dptr(y), abi.dptr(z), C.byref(opts), C.byref(res),
                           mask.ctypes.data_as(abi.c_uint8_p))
    return res, mask


def ransac_plane_batch(x, y, z, opts=None, seed_per_problem=True, threads=0):
    """x, y, z: (n_problems, n) arrays."""
    x, y, z = (abi.as_f64(a) for a in (x, y, z))
    npb, n = x.shape
    opts = opts or abi.RansacOptions.default()
    res = (abi.PlaneResult * npb)()
    mask = np.zeros((npb, n), dtype=np.uint8)
    lib().orc_ransac_plane_batch(npb, n, abi.dptr(x), abi.dptr(y), abi.dptr(z), C.byref(opts), int(seed_per_problem), res,
                                 mask.ctypes.data_as(abi.c_uint8_p), threads)
    return res, mask


def pose12_to_T(p):
    T = np.eye(4); T[:3, :3] = np.asarray(p[:9]).reshape(3, 3); T[:3, 3] = p[9:12]
    return T


def zhang_intrinsics(hmtx):
    """zhang_intrinsics_from_hs (zhang.cpp:183-208): (ok, K5 = fx, fy, cx, cy, skew)."""
    H = abi.as_f64(np.asarray(hmtx).reshape(-1, 9))
    k5 = np.zeros(5)
    ok = lib().orc_zhang_intrinsics(len(H), abi.dptr(H), abi.dptr(k5))
    return bool(ok), k5


def sanitize_intrinsics(k5, bounds10):
    k = abi.as_f64(k5).copy()
    mod = lib().orc_sanitize_intrinsics(abi.dptr(k), abi.dptr(abi.as_f64(bounds10)))
    return k, bool(mod)


def pose_from_homography(k5, H):
    """pose_from_homography (posefromhomography.cpp:12-67): (ok, T 4x4, scale, cond_check)."""
    out = np.zeros(12); sc = np.zeros(1); cd = np.zeros(1)
    ok = lib().orc_pose_from_homography(abi.dptr(abi.as_f64(k5)), abi.dptr(abi.as_f64(np.asarray(H).ravel())), abi.dptr(out),
                                        abi.dptr(sc), abi.dptr(cd))
    return bool(ok), pose12_to_T(out), float(sc[0]), float(cd[0])


def estimate_intrinsics_ransac(x, y, u, v, view_offset, ransac_opts, bounds10=None):
    """estimate_intrinsics with IntrinsicsEstimOptions::homography_ransac, one camera."""
    x, y, u, v = (abi.as_f64(a) for a in (x, y, u, v))
    off = np.ascontiguousarray(view_offset, dtype=np.int64)
    nv = len(off) - 1
    k5 = np.zeros(5); succ = np.zeros(nv, dtype=np.int32); H = np.zeros((nv, 9)); rms = np.zeros(nv); poses = np.zeros((nv, 12))
    mask = np.zeros(len(x), dtype=np.uint8)
    b = None if bounds10 is None else abi.as_f64(bounds10)
    ok = lib().orc_estimate_intrinsics_ransac(nv, abi.i64ptr(off), abi.dptr(x), abi.dptr(y), abi.dptr(u), abi.dptr(v), abi.dptr(b),
                                              C.byref(ransac_opts), abi.dptr(k5), abi.i32ptr(succ), abi.dptr(H), abi.dptr(rms), abi.dptr(poses),
                                              mask.ctypes.data_as(abi.c_uint8_p))
    return dict(success=bool(ok), kmtx=k5, view_success=succ, hmtx=H.reshape(nv, 3, 3), sym_rms=rms, poses=poses, inlier_mask=mask)


def estimate_intrinsics(x, y, u, v, view_offset, bounds10=None):
    """estimate_intrinsics (intrinsicsdlt.cpp:101-145) of one camera, no RANSAC."""
    x, y, u, v = (abi.as_f64(a) for a in (x, y, u, v))
    off = np.ascontiguousarray(view_offset, dtype=np.int64)
    nv = len(off) - 1
    k5 = np.zeros(5); succ = np.zeros(nv, dtype=np.int32); H = np.zeros((nv, 9)); rms = np.zeros(nv); poses = np.zeros((nv, 12))
    b = None if bounds10 is None else abi.as_f64(bounds10)
    ok = lib().orc_estimate_intrinsics(nv, abi.i64ptr(off), abi.dptr(x), abi.dptr(y), abi.dptr(u), abi.dptr(v), abi.dptr(b),
                                       abi.dptr(k5), abi.i32ptr(succ), abi.dptr(H), abi.dptr(rms), abi.dptr(poses))
    return dict(success=bool(ok), kmtx=k5, view_success=succ, hmtx=H.reshape(nv, 3, 3), sym_rms=rms, poses=poses)


def homography_dlt(x, y, u, v):
    L = lib()
    x, y, u, v = (abi.as_f64(a) for a in (x, y, u, v))
    H = np.zeros(9)
    rc = L.orc_homography_dlt(len(x), abi.dptr(x), abi.dptr(y), abi.dptr(u), abi.dptr(v), abi.dptr(H))
    return rc, H.reshape(3, 3)


def homography_testdata(n_points, noise, n_out, seed_out):
    L = lib()
    H = np.zeros(9); xyuv = np.zeros((n_points + n_out, 4))
    L.orc_ref_homography_data(n_points, noise, n_out, seed_out, abi.dptr(H), abi.dptr(xyuv))
    return H.reshape(3, 3), xyuv


def axxb_desc(rot_a, rot_b, tra_a, tra_b, huber_delta=1.0):
    d = abi.AxxbDesc()
    keep = [abi.as_f64(a) for a in (rot_a, rot_b, tra_a, tra_b)]
    d.n_pairs = len(keep[2].reshape(-1, 3))
    d.rot_a, d.rot_b, d.tra_a, d.tra_b = (abi.dptr(a) for a in keep)
    d.huber_delta = huber_delta
    d._keep = keep
    return d


def project_to_so3(R):
    out = np.zeros(9); lib().orc_project_to_so3(abi.dptr(abi.as_f64(np.asarray(R).reshape(9))), abi.dptr(out))
    return out.reshape(3, 3)


def log_so3(R):
    out = np.zeros(3); lib().orc_log_so3(abi.dptr(abi.as_f64(np.asarray(R).reshape(9))), abi.dptr(out))
    return out


def build_all_pairs(base_se3_gripper, cam_se3_target, min_angle_deg):
    from calibration_b200 import geometry as G
    L = lib()
    n = len(base_se3_gripper)
    bg = abi.as_f64(np.stack([G.pose_to_vec12(T) for T in base_se3_gripper]))
    ct = abi.as_f64(np.stack([G.pose_to_vec12(T) for T in cam_se3_target]))
    m = n * (n - 1) // 2
    ra, rb, ta, tb = np.zeros((m, 9)), np.zeros((m, 9)), np.zeros((m, 3)), np.zeros((m, 3))
    k = int(L.orc_build_all_pairs(n, abi.dptr(bg), abi.dptr(ct), min_angle_deg, abi.dptr(ra), abi.dptr(rb),
                                  abi.dptr(ta), abi.dptr(tb)))
    return ra[:k].copy(), rb[:k].copy(), ta[:k].copy(), tb[:k].copy()


def axxb_eval(d, x7, jac=True, threads=0):
    L = lib()
    cost = C.c_double(); g = np.zeros(6); H = np.zeros((6, 6))
    L.orc_axxb_eval(C.byref(d), abi.dptr(abi.as_f64(x7)), C.cast(C.byref(cost), abi.c_double_p),
                    abi.dptr(g) if jac else None, abi.dptr(H) if jac else None, threads)
    return cost.value, g, H


def axxb_solve(d, x7, opts=None):
    L = lib()
    x = abi.as_f64(x7).copy()
    opts = opts or abi.OptimOptions.default()
    res = abi.OptimResult(); cov = np.zeros((7, 7))
    L.orc_axxb_solve(C.byref(d), C.byref(opts), abi.dptr(x), C.byref(res), abi.dptr(cov))
    return x, res, cov
