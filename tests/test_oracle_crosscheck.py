"""Independent cross-checks of the oracle ("sanity cross-checks available in the image").

The reference's Ceres cannot run here, so its converged parameters on NOISY data are not recorded anywhere; what
can be checked is that the oracle's LM lands on the minimiser of the reference's objective, found by solvers that
share no code with it:
  * OpenCV's calibrateCamera — the same pinhole + Brown-Conrady model (k1 k2 p1 p2 k3, no skew) and the same
    sum of squared reprojection errors as optimize_intrinsics without a loss (optim/intrinsics.cpp:63-90);
  * scipy.optimize.least_squares on a numpy restatement of the residuals (geometry.project, written to
    SYNTHESISE the pixels) with the per-block Huber loss folded into the residuals: r_b * sqrt(rho(s_b) / s_b)
    has exactly the robustified cost as its sum of squares, whatever Gauss-Newton model a solver builds on it.
Neither is parity with the reference; both pin the objective (models, pose chains, loss per residual block) and
the minimiser the restated LM converges to."""
import numpy as np
import pytest
from scipy.optimize import least_squares
from scipy.spatial.transform import Rotation

import oracle_lib as O
from calibration_b200 import abi, synth
from calibration_b200 import geometry as G


def huber_scale(s, delta):
    """sqrt(rho(s) / s) for ceres::HuberLoss(delta): rho(s) = s for s <= delta^2, else 2 delta sqrt(s) - delta^2."""
    if delta <= 0 or s <= delta * delta:
        return 1.0
    return np.sqrt((2.0 * delta * np.sqrt(s) - delta * delta) / s)


def to_rt(T):
    return np.concatenate([Rotation.from_matrix(T[:3, :3]).as_rotvec(), T[:3, 3]])


def from_rt(p):
    T = np.eye(4); T[:3, :3] = Rotation.from_rotvec(p[:3]).as_matrix(); T[:3, 3] = p[3:6]
    return T


def test_intrinsics_minimiser_equals_opencv_calibrate_camera():
    cv2 = pytest.importorskip("cv2")
    prob, x0, _ = synth.make_intrinsics(seed=7, noise=0.2, huber_delta=-1.0)
    for name in ("x", "y", "u", "v"):   # calibrateCamera takes float32 points: give both solvers the same numbers
        a = getattr(prob, name); a[:] = a.astype(np.float32).astype(np.float64)
    off = np.asarray(prob.block_offset); nv = len(off) - 1
    x_o, res, _ = O.refine_solve(prob, x0, abi.OptimOptions.default(compute_covariance=0))
    assert res.success
    objp = [np.stack([prob.x[off[k]:off[k + 1]], prob.y[off[k]:off[k + 1]], np.zeros(off[k + 1] - off[k])], 1).astype(np.float32) for k in range(nv)]
    imgp = [np.stack([prob.u[off[k]:off[k + 1]], prob.v[off[k]:off[k + 1]]], 1).astype(np.float32).reshape(-1, 1, 2) for k in range(nv)]
    K0 = np.array([[x0[0], 0, x0[2]], [0, x0[1], x0[3]], [0, 0, 1.0]])
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 2000, 1e-16)
    rms, K, dist, rv, tv = cv2.calibrateCamera(objp, imgp, (1280, 720), K0.copy(), np.zeros(5), flags=cv2.CALIB_USE_INTRINSIC_GUESS, criteria=crit)
    rms_o = np.sqrt(2.0 * res.final_cost / len(prob.u))          # cost = 1/2 sum r^2 over 2 n_obs residuals; OpenCV: per point
    assert abs(rms - rms_o) <= 1e-9 * rms                         # same minimum of the same objective
    assert abs(K[0, 0] - x_o[0]) <= 1e-6 * x_o[0] and abs(K[1, 1] - x_o[1]) <= 1e-6 * x_o[1]
    assert abs(K[0, 2] - x_o[2]) <= 1e-3 and abs(K[1, 2] - x_o[3]) <= 1e-3 and x_o[4] == 0.0
    d = dist.ravel()                                              # OpenCV order k1 k2 p1 p2 k3; ours k1 k2 k3 p1 p2
    assert abs(d[0] - x_o[5]) <= 1e-4 * abs(x_o[5]) and abs(d[1] - x_o[6]) <= 1e-3 * abs(x_o[6]) and abs(d[4] - x_o[7]) <= 1e-3 * abs(x_o[7])
    assert abs(d[2] - x_o[8]) <= 1e-6 and abs(d[3] - x_o[9]) <= 1e-6
    _, poses = G.unpack_intrinsics(x_o, nv)
    for k in range(nv):
        R, _ = cv2.Rodrigues(rv[k])
        assert np.abs(R - poses[k][:3, :3]).max() <= 1e-5 and np.abs(tv[k].ravel() - poses[k][:3, 3]).max() <= 1e-5


def intrinsics_residuals(p, prob, off, delta):
    nv = len(off) - 1
    intr = np.concatenate([p[:4], [0.0], p[4:9]])
    out = []
    for k in range(nv):
        T = from_rt(p[9 + 6 * k:15 + 6 * k])
        s = slice(off[k], off[k + 1])
        P = np.stack([prob.x[s], prob.y[s], np.zeros(off[k + 1] - off[k])], 1) @ T[:3, :3].T + T[:3, 3]
        r = (G.project(intr, P) - np.stack([prob.u[s], prob.v[s]], 1)).ravel()
        out.append(r * huber_scale(r @ r, delta))
    return np.concatenate(out)


@pytest.mark.parametrize("delta", [1.0, -1.0])
def test_intrinsics_minimiser_equals_scipy_with_per_block_huber(delta):
    prob, x0, _ = synth.make_intrinsics(seed=11, n_views=12, noise=0.3, huber_delta=delta)
    off = np.asarray(prob.block_offset); nv = len(off) - 1
    x_o, res, _ = O.refine_solve(prob, x0, abi.OptimOptions.default(compute_covariance=0, epsilon=1e-13))
    assert res.success
    intr0, poses0 = G.unpack_intrinsics(x0, nv)
    p0 = np.concatenate([intr0[:4], intr0[5:10]] + [to_rt(T) for T in poses0])
    sol = least_squares(intrinsics_residuals, p0, args=(prob, off, delta), method="trf", jac="3-point", x_scale="jac", xtol=1e-15, ftol=1e-15,
                        gtol=1e-15, max_nfev=400)
    assert abs(sol.cost - res.final_cost) <= 1e-8 * res.final_cost       # scipy's cost is 1/2 sum f^2 as well
    intr_o, poses_o = G.unpack_intrinsics(x_o, nv)
    assert np.allclose(sol.x[:4], intr_o[:4], rtol=2e-5, atol=2e-3)
    assert np.allclose(sol.x[4:9], intr_o[5:10], rtol=5e-3, atol=2e-5)
    for k in range(nv):
        assert np.abs(from_rt(sol.x[9 + 6 * k:15 + 6 * k]) - poses_o[k])[:3].max() <= 2e-5
    if delta > 0:   # every block of 54 corners with 0.3 px noise is beyond delta = 1: the loss really acts
        r = intrinsics_residuals(sol.x, prob, off, -1.0)
        assert all((r[2 * off[k]:2 * off[k + 1]] ** 2).sum() > 1.0 for k in range(nv))


def bundle_residuals(p, prob, off, cam, bTg, n_cams, delta):
    out = []
    b_T_t = from_rt(p[9 * n_cams + 6 * n_cams:9 * n_cams + 6 * n_cams + 6])
    g_T_c = [from_rt(p[9 * n_cams + 6 * c:9 * n_cams + 6 * c + 6]) for c in range(n_cams)]
    for b in range(len(off) - 1):
        c = cam[b]
        q = p[9 * c:9 * c + 9]
        intr = np.concatenate([q[:4], [0.0], q[4:9]])
        T = np.linalg.inv(g_T_c[c]) @ np.linalg.inv(bTg[b]) @ b_T_t        # bundleresidual.h:15-27
        s = slice(off[b], off[b + 1])
        P = np.stack([prob.x[s], prob.y[s], np.zeros(off[b + 1] - off[b])], 1) @ T[:3, :3].T + T[:3, 3]
        r = (G.project(intr, P) - np.stack([prob.u[s], prob.v[s]], 1)).ravel()
        out.append(r * huber_scale(r @ r, delta))
    return np.concatenate(out)


@pytest.mark.parametrize("delta", [1.0, -1.0])
def test_bundle_minimiser_equals_scipy(delta):
    n_cams = 2
    prob, x0, _ = synth.make_bundle(seed=3, n_cams=n_cams, n_poses=14, noise=0.25, huber_delta=delta)
    off = np.asarray(prob.block_offset); cam = np.asarray(prob.block_cam)
    bTg = [O.pose12_to_T(p) for p in np.asarray(prob.block_b_se3_g).reshape(-1, 12)]
    x_o, res, _ = O.refine_solve(prob, x0, abi.OptimOptions.default(compute_covariance=0, epsilon=1e-13))
    assert res.success
    intr0, g0, b0 = G.unpack_bundle(x0, n_cams)
    p0 = np.concatenate([np.concatenate([k[:4], k[5:10]]) for k in intr0] + [to_rt(T) for T in g0] + [to_rt(b0)])
    sol = least_squares(bundle_residuals, p0, args=(prob, off, cam, bTg, n_cams, delta), method="trf", jac="3-point", x_scale="jac",
                        xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=400)
    assert abs(sol.cost - res.final_cost) <= 1e-8 * res.final_cost
    intr_o, g_o, b_o = G.unpack_bundle(x_o, n_cams)
    for c in range(n_cams):
        assert np.allclose(sol.x[9 * c:9 * c + 4], intr_o[c][:4], rtol=2e-5, atol=2e-3)
        assert np.abs(from_rt(sol.x[9 * n_cams + 6 * c:9 * n_cams + 6 * c + 6]) - g_o[c])[:3].max() <= 2e-5
    assert np.abs(from_rt(sol.x[15 * n_cams:15 * n_cams + 6]) - b_o)[:3].max() <= 2e-5


def test_covariance_blocks_equal_independent_inverse_normal_matrix():
    """compute_covariance (ceresutils.h:69-126) without a loss is (J^T J)^-1 lifted to the ambient blocks.  Its
    intrinsics block and its translation blocks do not depend on how rotations are parametrised, so they must equal
    the same blocks of inv(J^T J) for the numpy residuals above (rotation-vector poses, numeric Jacobian)."""
    from scipy.optimize._numdiff import approx_derivative
    prob, x0, _ = synth.make_intrinsics(seed=11, n_views=12, noise=0.3, huber_delta=-1.0)
    off = np.asarray(prob.block_offset); nv = len(off) - 1
    x_o, res, cov = O.refine_solve(prob, x0, abi.OptimOptions.default(compute_covariance=1, epsilon=1e-13))
    assert res.success and res.covariance_ok and cov.shape == (10 + 7 * nv, 10 + 7 * nv)
    intr_o, poses_o = G.unpack_intrinsics(x_o, nv)
    p = np.concatenate([intr_o[:4], intr_o[5:10]] + [to_rt(T) for T in poses_o])
    J = approx_derivative(intrinsics_residuals, p, method="3-point", args=(prob, off, -1.0))
    Cn = np.linalg.inv(J.T @ J)
    idx = [0, 1, 2, 3, 5, 6, 7, 8, 9]
    Co = cov[np.ix_(idx, idx)]
    assert np.abs(Cn[:9, :9] - Co).max() <= 1e-7 * np.abs(Co).max()
    assert not cov[4].any() and not cov[:, 4].any()                  # the constant skew has zero rows and columns
    t0 = 10 + 4 * nv                                                 # block order [intr, all quats, all trans] (intrinsics.cpp:14-61)
    for k in range(nv):
        Ct = Cn[9 + 6 * k + 3:9 + 6 * k + 6, 9 + 6 * k + 3:9 + 6 * k + 6]
        assert np.abs(Ct - cov[t0 + 3 * k:t0 + 3 * k + 3, t0 + 3 * k:t0 + 3 * k + 3]).max() <= 1e-7 * np.abs(Ct).max()
        # intrinsics x translation cross block
        Cx = Cn[:9, 9 + 6 * k + 3:9 + 6 * k + 6]
        assert np.abs(Cx - cov[np.ix_(idx, range(t0 + 3 * k, t0 + 3 * k + 3))]).max() <= 1e-7 * np.abs(Cx).max()


def extrinsics_residuals(p, prob, off, cam, view, n_cams, n_views, cam0, tgt0, delta):
    """Free parameters only: cameras 1.., views 1.. (camera 0 and view 0 are held at their start values,
    optim/extrinsics.cpp:122-139)."""
    o = 9 * n_cams
    c_T_r = [cam0] + [from_rt(p[o + 6 * (c - 1):o + 6 * c]) for c in range(1, n_cams)]
    o += 6 * (n_cams - 1)
    r_T_t = [tgt0] + [from_rt(p[o + 6 * (v - 1):o + 6 * v]) for v in range(1, n_views)]
    out = []
    for b in range(len(off) - 1):
        c = cam[b]
        q = p[9 * c:9 * c + 9]
        T = c_T_r[c] @ r_T_t[view[b]]                                   # extrinsicsresidual.h:14-20
        s = slice(off[b], off[b + 1])
        P = np.stack([prob.x[s], prob.y[s], np.zeros(off[b + 1] - off[b])], 1) @ T[:3, :3].T + T[:3, 3]
        r = (G.project(np.concatenate([q[:4], [0.0], q[4:9]]), P) - np.stack([prob.u[s], prob.v[s]], 1)).ravel()
        out.append(r * huber_scale(r @ r, delta))
    return np.concatenate(out)


@pytest.mark.parametrize("delta", [1.0, -1.0])
def test_extrinsics_minimiser_equals_scipy(delta):
    n_cams, n_views = 2, 10
    prob, x0, _ = synth.make_extrinsics(seed=5, n_cams=n_cams, n_views=n_views, noise=0.25, huber_delta=delta)
    off = np.asarray(prob.block_offset); cam = np.asarray(prob.block_cam); view = np.asarray(prob.block_view)
    x_o, res, _ = O.refine_solve(prob, x0, abi.OptimOptions.default(compute_covariance=0, epsilon=1e-13))
    assert res.success
    intr0, c0, t0 = G.unpack_extrinsics(x0, n_cams, n_views)
    p0 = np.concatenate([np.concatenate([k[:4], k[5:10]]) for k in intr0] + [to_rt(T) for T in c0[1:]] + [to_rt(T) for T in t0[1:]])
    sol = least_squares(extrinsics_residuals, p0, args=(prob, off, cam, view, n_cams, n_views, c0[0], t0[0], delta), method="trf",
                        jac="3-point", x_scale="jac", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=400)
    assert abs(sol.cost - res.final_cost) <= 1e-8 * res.final_cost
    intr_o, c_o, t_o = G.unpack_extrinsics(x_o, n_cams, n_views)
    assert np.abs(c_o[0] - c0[0]).max() == 0.0 and np.abs(t_o[0] - t0[0]).max() == 0.0    # the held blocks did not move
    for c in range(n_cams):
        assert np.allclose(sol.x[9 * c:9 * c + 4], intr_o[c][:4], rtol=5e-5, atol=5e-3)
    assert np.abs(from_rt(sol.x[9 * n_cams:9 * n_cams + 6]) - c_o[1])[:3].max() <= 5e-5


def axxb_residuals(p, ra, rb, ta, tb, delta):
    """AxXbResidual (residuals/handeyeresidual.h:25-49): [Log(R_A R_X R_B^T R_X^T); (R_A - I) t_X - (R_X t_B - t_A)]."""
    RX = Rotation.from_rotvec(p[:3]).as_matrix(); tX = p[3:6]
    RA = ra.reshape(-1, 3, 3); RB = rb.reshape(-1, 3, 3)
    S = RA @ RX @ np.transpose(RB, (0, 2, 1)) @ RX.T
    rot = Rotation.from_matrix(S).as_rotvec()
    tra = np.einsum("nij,j->ni", RA - np.eye(3), tX) - (tb @ RX.T - ta)
    r = np.concatenate([rot, tra], axis=1)
    s = (r * r).sum(axis=1)
    return (r * np.array([huber_scale(v, delta) for v in s])[:, None]).ravel()


@pytest.mark.parametrize("delta", [1.0, 0.002])
def test_axxb_minimiser_equals_scipy(delta):
    bg, ct, X_gt = synth.make_handeye_poses(seed=2024, n=16)
    rng = np.random.default_rng(1)
    ct = [synth.perturb_pose(rng, T, 0.3, 0.002) for T in ct]            # noisy camera poses: the minimum is not X_gt
    ra, rb, ta, tb = O.build_all_pairs(bg, ct, 1.0)
    assert len(ta) > 60
    d = O.axxb_desc(ra, rb, ta, tb, huber_delta=delta)
    X0 = synth.perturb_pose(rng, X_gt, 3.0, 0.01)
    q0, t0 = G.pose_to_qt(X0)
    x_o, res, _ = O.axxb_solve(d, np.concatenate([q0, t0]), abi.OptimOptions.default(epsilon=1e-14))
    assert res.success
    sol = least_squares(axxb_residuals, to_rt(X0), args=(ra, rb, ta, tb, delta), method="trf", jac="3-point", x_scale="jac", xtol=1e-15,
                        ftol=1e-15, gtol=1e-15, max_nfev=300)
    assert abs(sol.cost - res.final_cost) <= 1e-8 * res.final_cost
    assert np.abs(from_rt(sol.x) - G.qt_to_pose(x_o[:4], x_o[4:7]))[:3].max() <= 1e-6
    if delta < 1.0:   # the small delta puts pairs beyond it: the per-pair loss acts
        r = axxb_residuals(sol.x, ra, rb, ta, tb, -1.0).reshape(-1, 6)
        assert ((r * r).sum(axis=1) > delta * delta).mean() > 0.2
