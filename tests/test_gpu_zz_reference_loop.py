"""The CUDA RANSAC kernels against outputs of the REFERENCE's own calib::ransac<> loop (tests/golden/ransac_ref.npz,
generated in the build container from /root/reference through oracle/_ref; see tests/golden/make_golden.py):
same inlier set, same best.iters, model to 1e-7 relative (the tolerance of the GPU-vs-oracle RANSAC tests: the DLT null
vector comes from a different factorisation).  Cases whose oracle run reports a residual within 1e-10 of the
threshold are skipped (rounding could move a point across it)."""
import numpy as np
import pytest

import oracle_lib as O
from calibration_b200 import capi
from test_golden import check_against_reference_loop, ransac_ref_cases

pytestmark = pytest.mark.gpu


def test_cuda_ransac_reproduces_reference_loop_outputs():
    n_ok = 0
    for key, data, opts, ref in ransac_ref_cases():
        if key[0] == "h":
            margin = O.ransac(*data, opts)[0].min_margin
            res, mask = capi.ransac_homography_batch(*[d[None] for d in data], opts, seed_per_problem=False)
            model = res[0].hmtx
        else:
            margin = O.ransac_plane(*data, opts)[0].min_margin
            res, mask = capi.ransac_plane_batch(*[d[None] for d in data], opts, seed_per_problem=False)
            model = np.array(res[0].plane)
            if res[0].success and np.dot(model[:3], ref["model"][:3]) < 0:
                model = -model                                           # refit sign: see ransac_plane.cu header
        n_ok += check_against_reference_loop(key, res[0], mask[0], model, margin > 1e-10, ref, model_rtol=1e-7, rms_atol=1e-9)
    assert n_ok >= 11
