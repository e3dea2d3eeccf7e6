"""Oracle AX = XB path against the reference's own tests: se3 utilities (tests/unit/se3_utils_test.cpp), pair
construction and the Ceres refinement (tests/unit/handeye_test.cpp:54-60,101-152).  CPU only."""
import numpy as np

import oracle_lib as O
import ref_scenarios as RS
from calibration_b200 import abi
from calibration_b200 import geometry as G


def test_project_to_so3_returns_rotation():
    # Se3Utils.ProjectToSO3ReturnsRotation (:10-22)
    R = G.angle_axis_to_R([1.0, 0.0, 0.0], 0.3)
    P = R.copy(); P[0, 1] += 0.05; P[1, 0] -= 0.02
    Q = O.project_to_so3(P)
    assert abs(np.linalg.det(Q) - 1.0) < 1e-12 and np.linalg.norm(Q @ Q.T - np.eye(3)) < 1e-10
    # the closest rotation in the Frobenius norm: U V^T of numpy's SVD
    U, _, Vt = np.linalg.svd(P)
    assert np.abs(Q - U @ Vt).max() < 1e-12
    # a reflection is mapped to a proper rotation (sigma(2,2) = -1 branch, se3_utils.h:15-17)
    M = np.diag([1.0, 1.0, -1.0]) @ R
    Q = O.project_to_so3(M)
    assert abs(np.linalg.det(Q) - 1.0) < 1e-12 and np.linalg.norm(Q @ Q.T - np.eye(3)) < 1e-10


def test_log_exp_round_trip():
    # Se3Utils.LogExpRoundTrip (:24-29)
    w = np.array([0.1, -0.2, 0.3])
    R = G.angle_axis_to_R(w / np.linalg.norm(w), np.linalg.norm(w))
    assert np.linalg.norm(O.log_so3(R) - w) < 1e-10
    assert not O.log_so3(np.eye(3)).any()                      # theta < 1e-12 -> Zero (se3_utils.h:33-35)


def test_identical_poses_leave_no_valid_pair():
    # TsaiLenzAllPairsWeighted.ThrowsOnDegenerateSmallMotions (:54-60): build_all_pairs finds nothing to keep
    # (the reference then throws std::runtime_error, handeyedlt.cpp:76-79; the C ABI returns CAL_ERR_RUNTIME)
    I = [np.eye(4)] * 5
    ra, rb, ta, tb = O.build_all_pairs(I, I, 2.0)
    assert len(ta) == 0


def test_refinement_improves_over_initializer():
    # CeresAXXBRefine.ImprovesOverInitializer (:101-152), same RNG stream (RNG(2024): two axes, the sequence, one axis)
    b_se3_g, pre, post = O.handeye_sequence(2024, 18, n_pre=2, n_post=1)
    X_gt = G.make_pose([0.02, -0.01, 0.09], pre[0], RS.deg2rad(10.0))
    b_t = G.make_pose([0.25, 0.05, 0.55], pre[1], RS.deg2rad(18.0))
    c_se3_t = [G.inv_pose(X_gt) @ G.inv_pose(T) @ b_t for T in b_se3_g]
    X0 = X_gt.copy()
    X0[:3, :3] = G.angle_axis_to_R(post[0] / np.linalg.norm(post[0]), RS.deg2rad(2.0)) @ X0[:3, :3]
    X0[:3, 3] += [0.01, -0.005, 0.004]
    ra, rb, ta, tb = O.build_all_pairs(b_se3_g, c_se3_t, 0.5)  # optimize_handeye: min_angle 0.5 deg (optim/handeye.cpp:64)
    x, res, cov = O.axxb_solve(O.axxb_desc(ra, rb, ta, tb, 1.0), G.pack_handeye(X0), abi.OptimOptions.default(max_iterations=60))
    Xr = G.qt_to_pose(x[:4], x[4:7])
    err0_rot = np.rad2deg(G.rotation_angle(X0[:3, :3].T @ X_gt[:3, :3])); err0_tr = np.linalg.norm(X0[:3, 3] - X_gt[:3, 3])
    err1_rot = np.rad2deg(G.rotation_angle(Xr[:3, :3].T @ X_gt[:3, :3])); err1_tr = np.linalg.norm(Xr[:3, 3] - X_gt[:3, 3])
    assert err1_rot < err0_rot and err1_tr < err0_tr
    assert err1_rot < 0.05 and err1_tr < 0.002
    assert res.success and abs(np.linalg.norm(x[:4]) - 1.0) < 1e-12      # QuaternionManifold keeps the norm
