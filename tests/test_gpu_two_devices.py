"""Two handles on two devices in ONE process (needs a box with at least two GPUs; skipped otherwise).  The kernels that
use more than 48 KB of dynamic shared memory opt in per (function, device): a handle on the second device must get the
attribute too (the round-1 build set it once per process)."""
import numpy as np
import pytest

from calibration_b200 import abi, capi, synth

pytestmark = pytest.mark.gpu


def test_two_handles_on_two_devices(monkeypatch):
    if capi.device_count() < 2:
        pytest.skip("one GPU on this box")
    monkeypatch.setenv("CALIB_B200_FUSED", "1")   # the fused layout; the Scheimpflug + skew instance runs three role warps and 60 KB of shared memory
    for kw in (dict(model=abi.MODEL_SCHEIMPFLUG_BC5, optimize_skew=True), dict()):
        prob, x0, _ = synth.make_bundle(n_cams=2, n_poses=80, **kw)
        h0 = capi.RefineHandle(prob, device=0)
        h1 = capi.RefineHandle(prob, device=1)
        try:
            r0, r1 = h0.eval(x0), h1.eval(x0)
            assert r0[0] == r1[0] and np.array_equal(r0[1], r1[1]) and np.array_equal(r0[2], r1[2])
            c0, c1 = h0.cost(x0), h1.cost(x0)
            assert c0 == c1 and abs(c0 - r0[0]) <= 1e-12 * c0
            x_a, res_a, _ = h0.solve(x0)
            x_b, res_b, _ = h1.solve(x0)
            assert res_a.success and res_b.success and np.array_equal(x_a, x_b)
        finally:
            h0.close(); h1.close()
