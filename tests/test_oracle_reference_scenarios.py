"""The CPU oracle against the reference's own synthetic-recovery tests (same seeds, sizes,
perturbations and tolerances).  This is what pins the oracle."""
import numpy as np
import pytest

import oracle_lib as O
import ref_scenarios as RS
from calibration_b200 import abi
from calibration_b200 import geometry as G


@pytest.mark.parametrize("skew", [False, True])
def test_optimize_intrinsics_recovers(skew):
    # intrinsics_optimize_test.cpp:8-113
    prob, x0, info = RS.intrinsics_scenario(skew)
    x, res, cov = O.refine_solve(prob, x0)
    intr, _ = G.unpack_intrinsics(x, prob.desc.n_views)
    gt = info["intr_gt"]
    assert res.success
    assert np.abs(intr[:4] - gt[:4]).max() < 1e-6
    assert abs(intr[4] - gt[4]) < (1e-8 if skew else 1e-9)
    if not skew:
        assert res.final_cost < 1e-6
    assert res.covariance_ok and cov.shape == (10 + 7 * 15, 10 + 7 * 15)


@pytest.mark.parametrize("kind", ["nodist", "nodist_skew"])
def test_optimize_bundle_recovers_x_and_intrinsics(kind):
    # bundle_test.cpp:9-154
    prob, x0, info = RS.bundle_scenario(kind)
    x, res, _ = O.refine_solve(prob, x0)
    intr, g, b = G.unpack_bundle(x, 1)
    assert np.degrees(G.rotation_angle(g[0][:3, :3].T @ info["g_gt"][:3, :3])) < 1e-6
    assert np.linalg.norm(g[0][:3, 3] - info["g_gt"][:3, 3]) < 1e-6
    assert np.abs(intr[0][:4] - info["intr_gt"][:4]).max() < 1e-6
    assert abs(intr[0][4] - info["intr_gt"][4]) < 1e-9
    assert np.degrees(G.rotation_angle(b[:3, :3].T @ info["b_gt"][:3, :3])) < 1e-6
    assert np.linalg.norm(b[:3, 3] - info["b_gt"][:3, 3]) < 1e-6


def test_optimize_bundle_distortion_recovery():
    # bundle_test.cpp:156-210 (Huber default)
    prob, x0, info = RS.bundle_scenario("distortion")
    x, res, _ = O.refine_solve(prob, x0)
    intr, g, _ = G.unpack_bundle(x, 1)
    assert np.degrees(G.rotation_angle(g[0][:3, :3].T @ info["g_gt"][:3, :3])) < 0.1
    assert np.linalg.norm(g[0][:3, 3] - info["g_gt"][:3, 3]) < 0.02
    assert np.abs(intr[0][5:10] - info["intr_gt"][5:10]).max() < 1e-5


@pytest.mark.parametrize("which", ["intrinsics", "handeye"])
def test_scheimpflug_bundle(which):
    # scheimpflug_bundle_test.cpp:13-94
    prob, x0, info = RS.scheimpflug_scenario(which)
    x, res, _ = O.refine_solve(prob, x0)
    intr, g, _ = G.unpack_bundle(x, 1, 12)
    assert np.linalg.norm(g[0][:3, 3] - info["g_gt"][:3, 3]) < 1e-6
    assert G.rotation_angle(g[0][:3, :3] @ info["g_gt"][:3, :3].T) < 1e-6
    assert np.abs(intr[0][10:12] - info["intr_gt"][10:12]).max() < 1e-6


def test_scheimpflug_model_known_answers():
    # scheimpflug_test.cpp:11-51: zero tilt == pinhole; principal ray maps to project((-tan ty / cos tx, tan tx))
    rng = np.random.default_rng(3)
    pin = np.array([800.0, 820.0, 320.0, 240.0, 0.1, -0.1, 0.01, 0.001, 0.0005, -0.0004])
    for _ in range(20):
        P = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(0.5, 2.0)])
        a = O.project(abi.MODEL_PINHOLE_BC5, pin, P)
        b = O.project(abi.MODEL_SCHEIMPFLUG_BC5, np.concatenate([pin, [0.0, 0.0]]), P)
        assert np.abs(a - b).max() < 1e-9
    taux, tauy = 0.05, -0.03
    sc = np.concatenate([pin, [taux, tauy]])
    uv = O.project(abi.MODEL_SCHEIMPFLUG_BC5, sc, np.array([0.0, 0.0, 1.0]))
    mx0, my0 = -np.tan(tauy) / np.cos(taux), np.tan(taux)
    exp = np.array([pin[0] * mx0 + pin[4] * my0 + pin[2], pin[1] * my0 + pin[3]])
    assert np.abs(uv - exp).max() < 1e-9
    # numpy synthesis helper agrees with the oracle's templated model
    for _ in range(10):
        P = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(0.5, 2.0)])
        assert np.abs(G.project(sc, P) - O.project(abi.MODEL_SCHEIMPFLUG_BC5, sc, P)).max() < 1e-10


def test_extrinsics_recover_camera_and_target_poses():
    # extrinsics_test.cpp:9-73
    prob, x0, info = RS.extrinsics_scenario("poses")
    x, res, _ = O.refine_solve(prob, x0)
    _, cams, tg = G.unpack_extrinsics(x, 2, prob.desc.n_views)
    assert res.final_cost < 1e-6
    assert np.allclose(cams[1][:3, 3], info["cam_gt"][1][:3, 3], rtol=1e-3, atol=1e-3)
    assert np.allclose(cams[1][:3, :3], info["cam_gt"][1][:3, :3], atol=1e-3)
    for v, T in enumerate(info["target_gt"]):
        assert np.allclose(tg[v][:3, 3], T[:3, 3], rtol=1e-3, atol=1e-3)
        assert np.allclose(tg[v][:3, :3], T[:3, :3], atol=1e-3)


def test_extrinsics_recover_all_parameters_and_covariance():
    # extrinsics_test.cpp:75-140 (the only covariance assertion on the reprojection path)
    prob, x0, info = RS.extrinsics_scenario("all")
    x, res, cov = O.refine_solve(prob, x0)
    intrs, cams, tg = G.unpack_extrinsics(x, 2, prob.desc.n_views)
    assert res.final_cost < 1e-6
    assert abs(intrs[0][0] - 100.0) < 1e-3 and abs(intrs[0][1] - 100.0) < 1e-3
    assert np.allclose(cams[1][:3, 3], info["cam_gt"][1][:3, 3], atol=1e-3)
    assert np.allclose(tg[0][:3, 3], info["target_gt"][0][:3, 3], atol=1e-3)
    assert res.covariance_ok and np.trace(cov) > 0.0


def test_extrinsics_first_target_pose_fixed():
    # extrinsics_test.cpp:142-199
    prob, x0, info = RS.extrinsics_scenario("first_fixed")
    x, res, _ = O.refine_solve(prob, x0)
    _, _, tg = G.unpack_extrinsics(x, 2, prob.desc.n_views)
    assert np.abs(tg[0][:3, 3] - info["target_init"][0][:3, 3]).max() < 1e-12
    assert res.final_cost > 0.1


def test_schur_equals_dense_normal_equations():
    prob, x0, _ = RS.intrinsics_scenario(False)
    xs, rs, _ = O.refine_solve(prob, x0, want_cov=False)
    xd, rd, _ = O.refine_solve(prob, x0, force_dense=True, want_cov=False)
    assert rs.iterations == rd.iterations
    assert np.abs(xs - xd).max() < 1e-9


def test_oracle_gradient_matches_finite_differences():
    """The dual-number Jacobian: g = J^T r against central differences of the cost along tangent moves."""
    from calibration_b200 import synth
    prob, x0, _ = synth.make_bundle(n_cams=2, n_poses=6, huber_delta=-1.0)
    c0, g, H = O.refine_eval(prob, x0)
    P = 10
    # Euclidean blocks only (intrinsics without skew, translations): tangent == ambient there
    amb = [0, 1, 2, 3, 5, 6, 7, 8, 9]               # camera 0 intrinsics minus skew
    tan = list(range(9))
    off_gt = 2 * P + 4 * 2                           # g_t_c of camera 0
    n_tan_intr = 18
    amb += [off_gt, off_gt + 1, off_gt + 2]
    tan += [n_tan_intr + 6, n_tan_intr + 7, n_tan_intr + 8]  # after 2 quats (3 each)
    for a, t in zip(amb, tan):
        hstep = 1e-6 * max(1.0, abs(x0[a]))
        xp, xm = x0.copy(), x0.copy(); xp[a] += hstep; xm[a] -= hstep
        cp, _, _ = O.refine_eval(prob, xp, jac=False); cm, _, _ = O.refine_eval(prob, xm, jac=False)
        fd = (cp - cm) / (2 * hstep)
        assert abs(fd - g[t]) <= 1e-5 * max(1.0, abs(g[t])), (a, t, fd, g[t])
    assert np.allclose(H, H.T)
    assert np.linalg.eigvalsh(H).min() > -1e-6 * np.abs(H).max()
