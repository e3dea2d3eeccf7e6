"""ctypes binding of libcalib_b200.so — the thin host layer over the C ABI.

Everything that computes lives in the shared library (CUDA kernels for sm_100a);
this module only marshals numpy buffers.  There is no CPU fallback: without the
library or without a CUDA device the calls raise.
"""
import ctypes as C
import os

import numpy as np

from . import abi, build as _build

CAL_OK, CAL_ERR_INVALID_ARGUMENT, CAL_ERR_RUNTIME, CAL_ERR_CUDA, CAL_ERR_COMM = range(5)


class CalibCudaError(RuntimeError):
    """CUDA device missing or a CUDA call failed (the product has no CPU path)."""


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        path = _build.build()
    L = C.CDLL(path)
    dp, ip, u8p, i64 = abi.c_double_p, abi.c_int32_p, abi.c_uint8_p, C.c_int64
    hp = C.c_void_p
    L.cal_last_error.restype = C.c_char_p
    L.cal_device_count.restype = C.c_int
    L.cal_refine_create.argtypes = [C.POINTER(abi.ProblemDesc), C.c_int, C.POINTER(hp)]
    L.cal_refine_destroy.argtypes = [hp]
    L.cal_refine_param_count.restype = i64
    L.cal_refine_param_count.argtypes = [hp]
    L.cal_refine_tangent_count.restype = i64
    L.cal_refine_tangent_count.argtypes = [hp]
    L.cal_refine_eval.argtypes = [hp, dp, dp, dp, dp]
    L.cal_refine_cost.argtypes = [hp, dp, dp, dp]
    L.cal_refine_view_errors.argtypes = [hp, dp, dp, dp]
    L.cal_refine_bench_pass.argtypes = [hp, dp, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), dp]
    L.cal_refine_launch_count.restype = i64
    L.cal_refine_launch_count.argtypes = [hp]
    L.cal_refine_layout_info.argtypes = [hp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), ip, ip]
    L.cal_fp64_peak_tflops.argtypes = [C.c_int, dp]
    L.cal_refine_solve.argtypes = [hp, C.POINTER(abi.OptimOptions), dp, C.POINTER(abi.OptimResult), dp]
    L.cal_comm_unique_id.argtypes = [u8p]
    L.cal_comm_create.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.POINTER(hp)]
    L.cal_comm_destroy.argtypes = [hp]
    L.cal_refine_attach_comm.argtypes = [hp, hp]
    L.cal_comm_peer_export.argtypes = [hp, u8p]
    L.cal_comm_peer_enable.argtypes = [hp, u8p]
    L.cal_comm_peer_disable.argtypes = [hp]
    L.cal_comm_allreduce_test.argtypes = [hp, dp, C.c_int32, C.c_int]
    for name, argt in (
        ("cal_axxb_create", [C.POINTER(abi.AxxbDesc), C.c_int, C.POINTER(hp)]),
        ("cal_axxb_create_from_poses", [i64, dp, dp, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(hp),
                                        C.POINTER(i64)]),
        ("cal_axxb_destroy", [hp]),
        ("cal_axxb_eval", [hp, dp, dp, dp, dp]),
        ("cal_axxb_solve", [hp, C.POINTER(abi.OptimOptions), dp, C.POINTER(abi.OptimResult), dp]),
        ("cal_axxb_bench_pass", [hp, dp, C.c_int, C.POINTER(C.c_float)]),
        ("cal_axxb_attach_comm", [hp, hp]),
        ("cal_ransac_homography_batch", [i64, C.c_int32, dp, dp, dp, dp, C.POINTER(abi.RansacOptions), C.c_int, C.c_int,
                                         C.POINTER(abi.RansacResult), u8p]),
        ("cal_ransac_homography_batch_multi", [i64, C.c_int32, dp, dp, dp, dp, C.POINTER(abi.RansacOptions), C.c_int, C.c_int32,
                                               C.POINTER(C.c_int32), C.POINTER(abi.RansacResult), u8p]),
        ("cal_seed_intrinsics", [i64, abi.c_int64_p, ip, dp, dp, dp, dp, C.c_int32, C.POINTER(abi.SeedOptions), C.c_int, dp, ip,
                                 ip, dp, dp, dp]),
        ("cal_seed_intrinsics_ransac", [i64, abi.c_int64_p, ip, dp, dp, dp, dp, C.c_int32, C.POINTER(abi.SeedOptions),
                                        C.POINTER(abi.RansacOptions), C.c_int, dp, ip, ip, dp, dp, dp, u8p]),
        ("cal_seed_planar_poses", [i64, abi.c_int64_p, ip, dp, dp, dp, dp, C.c_int32, dp, C.c_int, dp, ip]),
        ("cal_dataset_write", [C.c_char_p, i64, C.c_int32, abi.c_int64_p, ip, dp, dp, dp, dp]),
        ("cal_dataset_open", [C.c_char_p, C.c_int, C.POINTER(abi.Dataset)]),
        ("cal_dataset_close", [C.POINTER(abi.Dataset)]),
        ("cal_dataset_from_planar_json", [C.POINTER(C.c_char_p), C.c_int32, C.c_int32, C.c_char_p, C.POINTER(i64), C.POINTER(i64)]),
        ("cal_ransac_plane_batch", [i64, C.c_int32, dp, dp, dp, C.POINTER(abi.RansacOptions), C.c_int, C.c_int,
                                    C.POINTER(abi.PlaneResult), u8p]),
        ("cal_ransac_plane_batch_dev", [i64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(abi.RansacOptions), C.c_int,
                                        C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
        ("cal_ransac_homography_batch_dev", [i64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.POINTER(abi.RansacOptions), C.c_int, C.c_void_p, C.c_void_p,
                                             C.POINTER(C.c_float)]),
    ):
        if hasattr(L, name):
            getattr(L, name).argtypes = argt
    _lib = L
    return L


def _check(rc):
    if rc == CAL_OK:
        return
    msg = lib().cal_last_error().decode(errors="replace")
    if rc == CAL_ERR_INVALID_ARGUMENT:
        raise ValueError(msg)           # the reference throws std::invalid_argument
    if rc == CAL_ERR_CUDA:
        raise CalibCudaError(msg)
    raise RuntimeError(msg)             # std::runtime_error


def device_count():
    return int(lib().cal_device_count())


def fp64_peak_tflops(device=0):
    v = C.c_double()
    _check(lib().cal_fp64_peak_tflops(device, C.cast(C.byref(v), abi.c_double_p)))
    return v.value


class RefineHandle:
    """cal_refine_handle: observations resident on one GPU, laid out for the kernels."""

    def __init__(self, problem, device=0):
        self.problem = problem  # keeps the host buffers alive during create
        self._h = C.c_void_p()
        _check(lib().cal_refine_create(C.byref(problem.desc), device, C.byref(self._h)))
        self.n_amb = int(lib().cal_refine_param_count(self._h))
        self.n_tan = int(lib().cal_refine_tangent_count(self._h))

    def close(self):
        if self._h:
            lib().cal_refine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def eval(self, x, jac=True):
        x = abi.as_f64(x)
        assert len(x) == self.n_amb
        cost = C.c_double()
        g = np.zeros(self.n_tan) if jac else None
        H = np.zeros((self.n_tan, self.n_tan)) if jac else None
        _check(lib().cal_refine_eval(self._h, abi.dptr(x), C.cast(C.byref(cost), abi.c_double_p), abi.dptr(g), abi.dptr(H)))
        return cost.value, g, H

    def cost(self, x, want_block_ssr=False):
        x = abi.as_f64(x)
        cost = C.c_double()
        ssr = np.zeros(self.problem.desc.n_blocks) if want_block_ssr else None
        _check(lib().cal_refine_cost(self._h, abi.dptr(x), C.cast(C.byref(cost), abi.c_double_p), abi.dptr(ssr)))
        return (cost.value, ssr) if want_block_ssr else cost.value

    def bench_pass(self, x, reps=1, jacobian=True):
        x = abi.as_f64(x)
        ms = C.c_float(); k1 = C.c_float(); cost = C.c_double()
        _check(lib().cal_refine_bench_pass(self._h, abi.dptr(x), reps, int(jacobian), C.byref(ms), C.byref(k1),
                                           C.cast(C.byref(cost), abi.c_double_p)))
        return ms.value, k1.value, cost.value

    def bench_breakdown(self):
        """(ms of the first set-up kernel, ms of reduction (+ all-reduce) + next set-up per pass) of the last bench_pass"""
        a = C.c_float(); b = C.c_float()
        lib().cal_refine_bench_breakdown.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        _check(lib().cal_refine_bench_breakdown(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def launch_count(self):
        return int(lib().cal_refine_launch_count(self._h))

    def view_errors(self, x):
        """(per-block RMS reprojection error in px, global RMS) at x."""
        rms = np.zeros(int(self.problem.desc.n_blocks)); g = C.c_double()
        _check(lib().cal_refine_view_errors(self._h, abi.dptr(abi.as_f64(x)), abi.dptr(rms), C.cast(C.byref(g), abi.c_double_p)))
        return rms, g.value

    def layout_info(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64(); d, e = C.c_int32(), C.c_int32()
        _check(lib().cal_refine_layout_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e)))
        return dict(n_segments=a.value, n_tiles=b.value, obs_bytes=c.value, local_entries=d.value, k1_passes=e.value)

    def solve(self, x0, opts=None, want_cov=True):
        x = abi.as_f64(x0).copy()
        assert len(x) == self.n_amb
        opts = opts or abi.OptimOptions.default()
        res = abi.OptimResult()
        cov = np.zeros((self.n_amb, self.n_amb)) if (want_cov and opts.compute_covariance) else None
        _check(lib().cal_refine_solve(self._h, C.byref(opts), abi.dptr(x), C.byref(res), abi.dptr(cov)))
        return x, res, cov

    def attach_comm(self, comm):
        self._comm = comm  # keep it alive
        _check(lib().cal_refine_attach_comm(self._h, comm._c if comm is not None else None))


def comm_unique_id():
    buf = (C.c_uint8 * 128)()
    _check(lib().cal_comm_unique_id(buf))
    return bytes(buf)


class Comm:
    """cal_comm: process-lifetime NCCL communicator (one process per GPU)."""

    def __init__(self, unique_id, rank, world, device):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._c = C.c_void_p()
        _check(lib().cal_comm_create(buf, rank, world, device, C.byref(self._c)))
        self.rank, self.world = rank, world

    def enable_peer(self, all_gather):
        """Switch the small all-reduces to the NVLink peer-memory kernel.  `all_gather(bytes) -> list of bytes in
        rank order` is supplied by the caller (e.g. torch.distributed.all_gather_object).  Every rank takes part in
        every exchange whatever happens locally, and the path is verified against NCCL on a rank-dependent vector
        before it is trusted; returns False (NCCL stays in use on ALL ranks) if any rank fails any phase."""
        L = lib()
        self.peer, self.peer_error = False, ""

        def agree(ok):
            return all(o == b"1" for o in all_gather(b"1" if ok else b"0"))

        buf = (C.c_uint8 * 64)()
        ok = L.cal_comm_peer_export(self._c, buf) == CAL_OK
        handles = all_gather(bytes(buf) if ok else b"")
        ok = ok and all(len(h) == 64 for h in handles)
        if ok:
            blob = (C.c_uint8 * (64 * self.world)).from_buffer_copy(b"".join(handles))
            ok = L.cal_comm_peer_enable(self._c, blob) == CAL_OK
        if not ok:
            self.peer_error = L.cal_last_error().decode(errors="replace")
        if not agree(ok):
            L.cal_comm_peer_disable(self._c)
            return False
        n = 1000
        ref = np.arange(n) * 0.25 + 1.0 / (self.rank + 3.0)
        a = ref.copy()
        for _ in range(3):   # several epochs: both parities of the double buffer
            a = ref.copy()
            ok = ok and L.cal_comm_allreduce_test(self._c, abi.dptr(a), n, 1) == CAL_OK
        b = ref.copy()
        ok = L.cal_comm_allreduce_test(self._c, abi.dptr(b), n, 0) == CAL_OK and ok
        expect = sum(np.arange(n) * 0.25 + 1.0 / (r + 3.0) for r in range(self.world))
        ok = ok and bool(np.allclose(a, expect, rtol=1e-14, atol=0) and np.array_equal(a, a) and np.allclose(a, b, rtol=1e-14, atol=0))
        if not agree(ok):
            self.peer_error = self.peer_error or "peer all-reduce self-test failed"
            L.cal_comm_peer_disable(self._c)
            return False
        self.peer = True
        return True

    def close(self):
        if self._c:
            lib().cal_comm_destroy(self._c)
            self._c = C.c_void_p()


class AxxbHandle:
    def __init__(self, rot_a, rot_b, tra_a, tra_b, huber_delta=1.0, device=0):
        self._keep = [abi.as_f64(a) for a in (rot_a, rot_b, tra_a, tra_b)]
        d = abi.AxxbDesc()
        d.n_pairs = len(self._keep[2].reshape(-1, 3))
        d.rot_a, d.rot_b, d.tra_a, d.tra_b = (abi.dptr(a) for a in self._keep)
        d.huber_delta = huber_delta
        self.desc = d
        self._h = C.c_void_p()
        _check(lib().cal_axxb_create(C.byref(d), device, C.byref(self._h)))

    @classmethod
    def from_poses(cls, base_se3_gripper, cam_se3_target, huber_delta=1.0, min_angle_deg=0.5, reject_axis_parallel=True,
                   axis_parallel_eps=1e-3, device=0):
        """optimize_handeye's inputs as they are (lists of 4x4 poses): build_all_pairs runs on the GPU and the
        motion pairs are formed on the fly in every pass."""
        def pack(poses):
            P = np.asarray(poses, dtype=np.float64)
            return abi.as_f64(np.concatenate([P[:, :3, :3].reshape(len(P), 9), P[:, :3, 3]], axis=1))
        self = cls.__new__(cls)
        g, c = pack(base_se3_gripper), pack(cam_se3_target)
        if len(g) != len(c):
            raise RuntimeError("Inconsistent hand-eye input sizes")   # handeyedlt.cpp:56-58
        self._keep = [g, c]
        self._h = C.c_void_p()
        kept = C.c_int64()
        _check(lib().cal_axxb_create_from_poses(len(g), abi.dptr(g), abi.dptr(c), min_angle_deg, int(reject_axis_parallel),
                                                axis_parallel_eps, huber_delta, device, C.byref(self._h), C.byref(kept)))
        self.n_pairs = kept.value
        return self

    def close(self):
        if self._h:
            lib().cal_axxb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def eval(self, x7):
        cost = C.c_double(); g = np.zeros(6); H = np.zeros((6, 6))
        _check(lib().cal_axxb_eval(self._h, abi.dptr(abi.as_f64(x7)), C.cast(C.byref(cost), abi.c_double_p), abi.dptr(g), abi.dptr(H)))
        return cost.value, g, H

    def attach_comm(self, comm):
        """shard the pair tiles over the ranks of `comm` (a capi.Comm), 28 sums all-reduced per pass; None detaches"""
        _check(lib().cal_axxb_attach_comm(self._h, comm._c if comm is not None else None))
        self._comm = comm

    def bench_pass(self, x7, reps=1):
        """device time (ms, CUDA events on the handle's stream) of `reps` residual + Jacobian passes"""
        ms = C.c_float()
        _check(lib().cal_axxb_bench_pass(self._h, abi.dptr(abi.as_f64(x7)), reps, C.byref(ms)))
        return ms.value

    def solve(self, x7, opts=None):
        x = abi.as_f64(x7).copy()
        opts = opts or abi.OptimOptions.default()
        res = abi.OptimResult(); cov = np.zeros((7, 7))
        _check(lib().cal_axxb_solve(self._h, C.byref(opts), abi.dptr(x), C.byref(res), abi.dptr(cov)))
        return x, res, cov


def ransac_homography_batch(x, y, u, v, opts=None, seed_per_problem=True, device=0, want_mask=True, devices=None):
    """x, y, u, v: (n_problems, n) float64.  Returns (results array, inlier mask).  devices = [ids]: the batch is split by
    problem over these devices (cal_ransac_homography_batch_multi), no communication, same results."""
    x, y, u, v = (abi.as_f64(a) for a in (x, y, u, v))
    npb, n = x.shape
    opts = opts or abi.RansacOptions.default()
    res = (abi.RansacResult * npb)()
    mask = np.zeros((npb, n), dtype=np.uint8) if want_mask else None
    if devices is not None:
        ids = (C.c_int32 * len(devices))(*devices)
        _check(lib().cal_ransac_homography_batch_multi(npb, n, abi.dptr(x), abi.dptr(y), abi.dptr(u), abi.dptr(v), C.byref(opts),
                                                       int(seed_per_problem), len(devices), ids, res,
                                                       mask.ctypes.data_as(C.POINTER(C.c_uint8)) if want_mask else None))
        return res, mask
    _check(lib().cal_ransac_homography_batch(npb, n, abi.dptr(x), abi.dptr(y), abi.dptr(u), abi.dptr(v), C.byref(opts),
                                             int(seed_per_problem), device, res,
                                             mask.ctypes.data_as(abi.c_uint8_p) if want_mask else None))
    return res, mask


def ransac_plane_batch(x, y, z, opts=None, seed_per_problem=True, device=0, want_mask=True):
    """fit_plane_ransac (linear/planefit.h:23-24) batched: x, y, z (n_problems, n) float64.
    Returns (PlaneResult array, inlier mask)."""
    x, y, z = (abi.as_f64(a) for a in (x, y, z))
    npb, n = x.shape
    opts = opts or abi.RansacOptions.default()
    res = (abi.PlaneResult * npb)()
    mask = np.zeros((npb, n), dtype=np.uint8) if want_mask else None
    _check(lib().cal_ransac_plane_batch(npb, n, abi.dptr(x), abi.dptr(y), abi.dptr(z), C.byref(opts), int(seed_per_problem),
                                        device, res, mask.ctypes.data_as(abi.c_uint8_p) if want_mask else None))
    return res, mask


def _views(x, y, u, v, view_offset, view_cam):
    x, y, u, v = (abi.as_f64(a) for a in (x, y, u, v))
    off = np.ascontiguousarray(view_offset, dtype=np.int64)
    nv = len(off) - 1
    cam = np.zeros(nv, dtype=np.int32) if view_cam is None else np.ascontiguousarray(view_cam, dtype=np.int32)
    return x, y, u, v, off, cam, nv


def seed_intrinsics(x, y, u, v, view_offset, view_cam=None, n_cams=1, bounds=None, device=0, ransac=None):
    """estimate_intrinsics (linear/intrinsics.h:58-59) for every camera, batched on the GPU: Zhang's K,
    per-view homographies, symmetric rms and poses from the homographies.  ransac = RansacOptions switches the
    per-view homographies to ransac<HomographyEstimator> (IntrinsicsEstimOptions::homography_ransac)."""
    x, y, u, v, off, cam, nv = _views(x, y, u, v, view_offset, view_cam)
    opts = abi.SeedOptions.from_bounds(bounds)
    kmtx = np.zeros((n_cams, 5)); cam_ok = np.zeros(n_cams, dtype=np.int32); ok = np.zeros(nv, dtype=np.int32)
    H = np.zeros((nv, 9)); rms = np.zeros(nv); poses = np.zeros((nv, 12))
    if ransac is not None:
        mask = np.zeros(len(x), dtype=np.uint8)
        _check(lib().cal_seed_intrinsics_ransac(nv, abi.i64ptr(off), abi.i32ptr(cam), abi.dptr(x), abi.dptr(y), abi.dptr(u), abi.dptr(v),
                                                n_cams, C.byref(opts), C.byref(ransac), device, abi.dptr(kmtx), abi.i32ptr(cam_ok), abi.i32ptr(ok),
                                                abi.dptr(H), abi.dptr(rms), abi.dptr(poses), mask.ctypes.data_as(abi.c_uint8_p)))
        return dict(kmtx=kmtx, cam_success=cam_ok, view_success=ok, hmtx=H.reshape(nv, 3, 3), sym_rms=rms, poses=poses, inlier_mask=mask)
    _check(lib().cal_seed_intrinsics(nv, abi.i64ptr(off), abi.i32ptr(cam), abi.dptr(x), abi.dptr(y), abi.dptr(u), abi.dptr(v),
                                     n_cams, C.byref(opts), device, abi.dptr(kmtx), abi.i32ptr(cam_ok), abi.i32ptr(ok), abi.dptr(H),
                                     abi.dptr(rms), abi.dptr(poses)))
    return dict(kmtx=kmtx, cam_success=cam_ok, view_success=ok, hmtx=H.reshape(nv, 3, 3), sym_rms=rms, poses=poses)


def seed_planar_poses(x, y, u, v, view_offset, kmtx, view_cam=None, device=0):
    """estimate_planar_pose (linear/planarpose.h:34) for every view, batched on the GPU.  kmtx: (n_cams, 5)."""
    x, y, u, v, off, cam, nv = _views(x, y, u, v, view_offset, view_cam)
    k = abi.as_f64(np.asarray(kmtx).reshape(-1, 5))
    poses = np.zeros((nv, 12)); ok = np.zeros(nv, dtype=np.int32)
    _check(lib().cal_seed_planar_poses(nv, abi.i64ptr(off), abi.i32ptr(cam), abi.dptr(x), abi.dptr(y), abi.dptr(u), abi.dptr(v),
                                       len(k), abi.dptr(k), device, abi.dptr(poses), abi.i32ptr(ok)))
    return poses, ok


class Dataset:
    """Columnar observation store (cal_dataset_*): numpy views straight onto the read-only mapping."""

    def __init__(self, path, pin=False):
        self._d = abi.Dataset()
        _check(lib().cal_dataset_open(os.fsencode(path), int(pin), C.byref(self._d)))
        d = self._d
        self.n_views, self.n_obs, self.n_cams, self.pinned = int(d.n_views), int(d.n_obs), int(d.n_cams), bool(d.pinned)
        as_np = lambda p, n, dt: np.ctypeslib.as_array(p, shape=(n,)) if n else np.zeros(0, dtype=dt)
        self.view_offset = as_np(d.view_offset, self.n_views + 1, np.int64)
        self.view_cam = as_np(d.view_cam, self.n_views, np.int32)
        self.x, self.y = as_np(d.obj_x, self.n_obs, np.float64), as_np(d.obj_y, self.n_obs, np.float64)
        self.u, self.v = as_np(d.img_u, self.n_obs, np.float64), as_np(d.img_v, self.n_obs, np.float64)

    def close(self):
        if self._d.impl:
            self.view_offset = self.view_cam = self.x = self.y = self.u = self.v = None
            lib().cal_dataset_close(C.byref(self._d))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @staticmethod
    def write(path, view_offset, view_cam, x, y, u, v, n_cams=None):
        off = np.ascontiguousarray(view_offset, dtype=np.int64); cam = np.ascontiguousarray(view_cam, dtype=np.int32)
        x, y, u, v = (abi.as_f64(a) for a in (x, y, u, v))
        n_cams = int(cam.max()) + 1 if n_cams is None else n_cams
        _check(lib().cal_dataset_write(os.fsencode(path), len(off) - 1, n_cams, abi.i64ptr(off), abi.i32ptr(cam), abi.dptr(x), abi.dptr(y),
                                       abi.dptr(u), abi.dptr(v)))

    @staticmethod
    def from_planar_json(json_paths, out_path, min_corners_per_view=0):
        """One PlanarDetections JSON per camera (the reference's dataset schema) -> columnar file; returns (n_views, n_obs)."""
        arr = (C.c_char_p * len(json_paths))(*[os.fsencode(p) for p in json_paths])
        nv, no = C.c_int64(), C.c_int64()
        _check(lib().cal_dataset_from_planar_json(arr, len(json_paths), int(min_corners_per_view), os.fsencode(out_path), C.byref(nv), C.byref(no)))
        return nv.value, no.value
