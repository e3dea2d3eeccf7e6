"""The conservative CPU baseline of bench.py (oracle/analytic_pass.cpp: hand-derived Jacobians, OpenMP over residual
blocks) against the dual-number restatement of the reference's autodiff cost functions (oracle/refine.cpp)."""
import numpy as np
import pytest

import oracle_lib as O
from calibration_b200 import abi, synth


@pytest.mark.parametrize("kw", [dict(), dict(optimize_intrinsics=False), dict(optimize_skew=True), dict(model=abi.MODEL_SCHEIMPFLUG_BC5),
                                dict(optimize_target_pose=False), dict(huber_delta=-1.0)])
def test_analytic_pass_matches_the_dual_number_restatement(kw):
    prob, x0, _ = synth.make_bundle(seed=11, n_cams=3, n_poses=40, **kw)
    c_d, g_d, H_d = O.refine_eval(prob, x0)
    for threads in (1, 3):
        c_a, g_a, H_a = O.analytic_bundle_eval(prob, x0, threads=threads)
        assert abs(c_a - c_d) <= 1e-12 * abs(c_d)
        assert np.abs(g_a - g_d).max() <= 1e-10 * np.abs(g_d).max()
        assert np.abs(H_a - H_d).max() <= 1e-10 * np.abs(H_d).max()


def test_analytic_pass_reads_the_shared_board_form():
    prob, x0, _ = synth.make_bundle(seed=12, n_cams=2, n_poses=30)
    c0, g0, H0 = O.analytic_bundle_eval(prob, x0)
    c1, g1, H1 = O.analytic_bundle_eval(prob.with_shared_board(), x0)
    assert c0 == c1 and np.array_equal(g0, g1) and np.array_equal(H0, H1)
