"""The seeding-stage oracle (oracle/seed.cpp) against the reference's own tests for these functions,
re-expressed with the same seeds, sizes and tolerances (runs on CPU):
  tests/unit/posefromhomography_test.cpp, planarpose_test.cpp (HomographyDecomposition, DLTEstimation),
  intrinsics_estimate_test.cpp (RecoversCameraMatrix, FailsWithTooFewViews, SanitizeIntrinsics)."""
import numpy as np

import oracle_lib as O
import ref_scenarios as RS
from calibration_b200 import geometry as G


def rot(axis, a):
    axis = np.asarray(axis, float)
    return G.angle_axis_to_R(axis / np.linalg.norm(axis), a)


def kmat(k5):
    return np.array([[k5[0], k5[4], k5[2]], [0, k5[1], k5[3]], [0, 0, 1.0]])


def test_pose_from_homography_recovers_pose():  # posefromhomography_test.cpp:9-37
    k5 = np.array([800.0, 820.0, 320.0, 240.0, 0.0])
    R = rot([1, 0, 0], 0.2) @ rot([0, 1, 0], -0.1) @ rot([0, 0, 1], 0.15); t = np.array([0.1, -0.2, 3.0])
    ok, T, scale, cond = O.pose_from_homography(k5, kmat(k5) @ np.column_stack([R[:, 0], R[:, 1], t]))
    assert ok and np.allclose(T[:3, :3], R, atol=1e-9) and np.allclose(T[:3, 3], t, atol=1e-9)
    assert abs(scale - 1.0) < 1e-12 and abs(cond - 1.0) < 1e-12


def test_pose_from_homography_negative_z_flips():  # posefromhomography_test.cpp:39-64
    k5 = np.array([500.0, 510.0, 320.0, 240.0, 0.0])
    R = rot([1, 0, 0], 0.05) @ rot([0, 1, 0], 0.1); t = np.array([0.2, 0.1, -2.0])
    ok, T, _, _ = O.pose_from_homography(k5, kmat(k5) @ np.column_stack([R[:, 0], R[:, 1], t]))
    assert ok and T[2, 3] > 0 and np.allclose(T[:3, :3], -R, atol=1e-9) and np.allclose(T[:3, 3], -t, atol=1e-9)


def test_pose_from_degenerate_homography_fails():  # posefromhomography_test.cpp:66-77
    ok, _, _, _ = O.pose_from_homography(np.array([400.0, 400.0, 320.0, 240.0, 0.0]), np.zeros((3, 3)))
    assert not ok


def test_planar_pose_dlt():  # planarpose_test.cpp:60-93
    k5 = np.array([1000.0, 1000.0, 500.0, 500.0, 0.0])
    T = np.eye(4); T[:3, :3] = rot([1, 1, 1], 0.1); T[:3, 3] = [0.1, 0.2, 2.0]
    obj = np.array([[i * 0.1, j * 0.1] for i in range(-5, 6, 2) for j in range(-5, 6, 2)])
    P = obj @ T[:3, :2].T + T[:3, 3]
    uv = np.column_stack([k5[0] * P[:, 0] / P[:, 2] + k5[2], k5[1] * P[:, 1] / P[:, 2] + k5[3]])
    est = O.estimate_planar_pose(obj[:, 0], obj[:, 1], uv[:, 0], uv[:, 1], k5)
    assert np.allclose(est[:3, :3], T[:3, :3], atol=1e-9) and np.allclose(est[:3, 3], T[:3, 3], atol=1e-9)


def intrinsics_estimate_scenario(seed=10, n_frames=8, rows=6, cols=9, spacing=0.03, k=(900.0, 920.0, 640.0, 360.0)):
    intr = np.array([k[0], k[1], k[2], k[3], 0, 0, 0, 0, 0, 0])
    _, c_se3_t, views, _, _ = RS.sim_handeye(seed, n_frames, np.eye(4), G.make_pose([0.0, 0.0, 2.0]), intr, rows, cols, spacing)
    return intr, c_se3_t, RS.views_to_soa(views)


def test_estimate_intrinsics_recovers_camera_matrix():  # intrinsics_estimate_test.cpp:11-55
    intr, c_se3_t, (xs, ys, us, vs, off) = intrinsics_estimate_scenario()
    r = O.estimate_intrinsics(xs, ys, us, vs, off)
    assert r["success"] and np.abs(r["kmtx"][:4] - intr[:4]).max() < 1e-6 and abs(r["kmtx"][4]) < 1e-9
    for i, T in enumerate(c_se3_t):
        est = O.pose12_to_T(r["poses"][i])
        assert min(np.abs(est[:3, :3] - T[:3, :3]).max(), np.abs(est[:3, :3] + T[:3, :3]).max()) < 1e-6
        assert abs(np.dot(est[:3, 3], T[:3, 3])) / (np.linalg.norm(est[:3, 3]) * np.linalg.norm(T[:3, 3])) > 0.999


def test_estimate_intrinsics_fails_with_too_few_views():  # intrinsics_estimate_test.cpp:57-81
    _, _, (xs, ys, us, vs, off) = intrinsics_estimate_scenario(seed=5, n_frames=3, rows=5, cols=7, spacing=0.04, k=(800.0, 805.0, 320.0, 240.0))
    assert not O.estimate_intrinsics(xs, ys, us, vs, off)["success"]


def test_sanitize_intrinsics_clamps():  # intrinsics_estimate_test.cpp:83-106
    k, mod = O.sanitize_intrinsics([-50.0, np.inf, -100.0, 2000.0, np.nan], [200.0, 2000.0, 150.0, 2000.0, 100.0, 200.0, 50.0, 75.0, -1.0, 1.0])
    assert mod and k.tolist() == [200.0, 150.0, 150.0, 62.5, 0.0]
