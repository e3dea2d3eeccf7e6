"""ctypes mirror of include/calib_b200.h (the C ABI of libcalib_b200.so).

Plain-old-data structs only; nothing here computes.  The CPU oracle
(oracle/oracle_api.h, test infrastructure) uses byte-identical layouts, so the
tests drive both sides with the same objects.
"""
import ctypes as C

import numpy as np

KIND_INTRINSICS, KIND_EXTRINSICS, KIND_BUNDLE = 0, 1, 2
MODEL_PINHOLE_BC5, MODEL_SCHEIMPFLUG_BC5 = 0, 1

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint8_p = C.POINTER(C.c_uint8)


class ProblemDesc(C.Structure):
    """cal_problem_desc — one refinement problem in SoA/CSR form."""

    _fields_ = [
        ("kind", C.c_int32),
        ("model", C.c_int32),
        ("n_cams", C.c_int32),
        ("n_views", C.c_int32),
        ("n_blocks", C.c_int64),
        ("n_obs", C.c_int64),
        ("obj_x", c_double_p),
        ("obj_y", c_double_p),
        ("img_u", c_double_p),
        ("img_v", c_double_p),
        ("block_offset", c_int64_p),
        ("block_cam", c_int32_p),
        ("block_view", c_int32_p),
        ("block_b_se3_g", c_double_p),
        ("optimize_intrinsics", C.c_int32),
        ("optimize_skew", C.c_int32),
        ("optimize_extrinsics", C.c_int32),
        ("optimize_target_pose", C.c_int32),
        ("optimize_hand_eye", C.c_int32),
        ("view_base", C.c_int32),
        ("huber_delta", C.c_double),
        ("board_x", c_double_p),
        ("board_y", c_double_p),
        ("board_n", C.c_int32),
        ("n_views_total", C.c_int32),
    ]


class OptimOptions(C.Structure):
    """cal_optim_options — mirrors calib::OptimOptions (optim/optimize.h:24-33)."""

    _fields_ = [
        ("optimizer", C.c_int32),
        ("max_iterations", C.c_int32),
        ("epsilon", C.c_double),
        ("compute_covariance", C.c_int32),
        ("verbose", C.c_int32),
        ("num_threads", C.c_int32),
        ("reserved", C.c_int32),
    ]

    @classmethod
    def default(cls, **kw):
        o = cls(0, 1000, 1e-9, 1, 0, 0, 0)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class OptimResult(C.Structure):
    _fields_ = [
        ("success", C.c_int32),
        ("iterations", C.c_int32),
        ("num_jac_evals", C.c_int32),
        ("num_cost_evals", C.c_int32),
        ("termination", C.c_int32),
        ("covariance_ok", C.c_int32),
        ("initial_cost", C.c_double),
        ("final_cost", C.c_double),
        ("report", C.c_char * 256),
    ]


class AxxbDesc(C.Structure):
    _fields_ = [
        ("n_pairs", C.c_int64),
        ("rot_a", c_double_p),
        ("rot_b", c_double_p),
        ("tra_a", c_double_p),
        ("tra_b", c_double_p),
        ("huber_delta", C.c_double),
    ]


class RansacOptions(C.Structure):
    """cal_ransac_options — mirrors calib::RansacOptions (common/ransac.h:22-29)."""

    _fields_ = [
        ("max_iters", C.c_int32),
        ("min_inliers", C.c_int32),
        ("thresh", C.c_double),
        ("confidence", C.c_double),
        ("seed", C.c_uint64),
        ("refit_on_inliers", C.c_int32),
        ("reserved", C.c_int32),
    ]

    @classmethod
    def default(cls, **kw):
        o = cls(1000, 12, 2.0, 0.99, 1234567, 1, 0)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class RansacResult(C.Structure):
    _fields_ = [
        ("success", C.c_int32),
        ("iters", C.c_int32),
        ("n_inliers", C.c_int32),
        ("iters_run", C.c_int32),
        ("hmtx", C.c_double * 9),
        ("inlier_rms", C.c_double),
        ("symmetric_rms_px", C.c_double),
        ("min_margin", C.c_double),
    ]


class PlaneResult(C.Structure):
    """cal_plane_ransac_result — PlaneRansacResult (linear/planefit.h:14-19) plus the loop counters."""
    _fields_ = [
        ("success", C.c_int32),
        ("iters", C.c_int32),
        ("n_inliers", C.c_int32),
        ("iters_run", C.c_int32),
        ("plane", C.c_double * 4),
        ("inlier_rms", C.c_double),
        ("min_margin", C.c_double),
    ]


class SeedOptions(C.Structure):
    """cal_seed_options: IntrinsicsEstimOptions.bounds (optional CalibrationBounds)."""
    _fields_ = [("use_bounds", C.c_int32), ("reserved", C.c_int32)] + [(n, C.c_double) for n in (
        "fx_min", "fx_max", "fy_min", "fy_max", "cx_min", "cx_max", "cy_min", "cy_max", "skew_min", "skew_max")]

    @classmethod
    def from_bounds(cls, b):
        """b = None (std::nullopt) or 10 values; the reference's defaults (camera_matrix.h:51-60) are
        (0, 2000, 0, 2000, 0, 1280, 0, 720, -0.01, 0.01)."""
        o = cls()
        if b is not None:
            o.use_bounds = 1
            (o.fx_min, o.fx_max, o.fy_min, o.fy_max, o.cx_min, o.cx_max, o.cy_min, o.cy_max, o.skew_min, o.skew_max) = [float(t) for t in b]
        return o


class Dataset(C.Structure):
    """cal_dataset: columnar observation store (mmap-ed)."""
    _fields_ = [("n_views", C.c_int64), ("n_obs", C.c_int64), ("n_cams", C.c_int32), ("pinned", C.c_int32),
                ("view_offset", c_int64_p), ("view_cam", c_int32_p), ("obj_x", c_double_p), ("obj_y", c_double_p),
                ("img_u", c_double_p), ("img_v", c_double_p), ("impl", C.c_void_p)]


def dptr(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i64ptr(a):
    return a.ctypes.data_as(c_int64_p) if a is not None else None


def i32ptr(a):
    return a.ctypes.data_as(c_int32_p) if a is not None else None


class Problem:
    """Owns the numpy buffers a ProblemDesc points into."""

    def __init__(self, kind, model, n_cams, n_views, x, y, u, v, block_offset, block_cam, block_view=None,
                 block_b_se3_g=None, optimize_intrinsics=True, optimize_skew=False, optimize_extrinsics=True,
                 optimize_target_pose=True, optimize_hand_eye=True, huber_delta=1.0):
        self.x, self.y, self.u, self.v = (as_f64(a) for a in (x, y, u, v))
        self.block_offset = np.ascontiguousarray(block_offset, dtype=np.int64)
        self.block_cam = np.ascontiguousarray(block_cam, dtype=np.int32)
        nb = len(self.block_cam)
        self.block_view = np.ascontiguousarray(
            block_view if block_view is not None else np.arange(nb), dtype=np.int32)
        self.block_b_se3_g = as_f64(block_b_se3_g).reshape(nb, 12) if block_b_se3_g is not None else None
        assert len(self.block_offset) == nb + 1 and self.block_offset[-1] == len(self.x)
        d = ProblemDesc()
        d.kind, d.model, d.n_cams, d.n_views = kind, model, n_cams, n_views
        d.n_blocks, d.n_obs = nb, len(self.x)
        d.obj_x, d.obj_y, d.img_u, d.img_v = dptr(self.x), dptr(self.y), dptr(self.u), dptr(self.v)
        d.block_offset = self.block_offset.ctypes.data_as(c_int64_p)
        d.block_cam = self.block_cam.ctypes.data_as(c_int32_p)
        d.block_view = self.block_view.ctypes.data_as(c_int32_p)
        d.block_b_se3_g = dptr(self.block_b_se3_g)
        d.optimize_intrinsics = int(optimize_intrinsics)
        d.optimize_skew = int(optimize_skew)
        d.optimize_extrinsics = int(optimize_extrinsics)
        d.optimize_target_pose = int(optimize_target_pose)
        d.optimize_hand_eye = int(optimize_hand_eye)
        d.huber_delta = float(huber_delta)
        self.desc = d

    def with_shared_board(self):
        """The same problem in the shared-board form of cal_problem_desc (board_n > 0): every residual block must
        observe the same object points in the same order; obj_x / obj_y are then NULL in the descriptor and only
        one board crosses PCIe.  Raises ValueError when the views do not share one board."""
        import copy
        off = self.block_offset
        n = int(off[1] - off[0])
        if not np.all(np.diff(off) == n):
            raise ValueError("views of different sizes cannot share a board")
        bx, by = self.x[:n].copy(), self.y[:n].copy()
        if not (np.array_equal(self.x.reshape(-1, n), np.broadcast_to(bx, (len(off) - 1, n)))
                and np.array_equal(self.y.reshape(-1, n), np.broadcast_to(by, (len(off) - 1, n)))):
            raise ValueError("object points differ between views")
        p = copy.copy(self)
        p.board_x, p.board_y = bx, by
        d = ProblemDesc.from_buffer_copy(self.desc)
        d.obj_x, d.obj_y = None, None
        d.board_x, d.board_y, d.board_n = dptr(bx), dptr(by), n
        p.desc = d
        return p

    @property
    def intr_size(self):
        return 12 if self.desc.model == MODEL_SCHEIMPFLUG_BC5 else 10

    @property
    def n_amb(self):
        d, P = self.desc, self.intr_size
        if d.kind == KIND_INTRINSICS:
            return P + 7 * d.n_views
        if d.kind == KIND_EXTRINSICS:
            return d.n_cams * (P + 7) + 7 * d.n_views
        return d.n_cams * (P + 7) + 7
