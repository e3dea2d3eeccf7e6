"""Generates the committed golden fixtures of tests/golden/ (run from the repository root:
`python tests/golden/make_golden.py`).

The reference itself cannot be compiled or imported in this image (Ceres / Eigen / Boost.PFR absent, see
DESIGN.md §2), so the fixtures come from the two things that CAN be run here:
  * std_sample.json    — the REAL libstdc++ std::sample / std::mt19937_64 of this toolchain (the reference's
                         hypothesis stream, include/calib/estimation/common/ransac.h:135,144-145), drawn by
                         oracle/ransac.cpp::orc_sample_stream_libstdcxx, plus the known answers recorded in
                         SURVEY Appendix C;
  * ransac_ref.npz     — outputs of the REFERENCE's own calib::ransac<> template (common/ransac.h:121-194), compiled from
                         /root/reference into oracle/_ref/libref_ransac.so around the oracle's estimator hooks
                         (oracle/ref_ransac_harness.cpp): inputs, options and the loop's results (success, best.iters,
                         inlier mask, model, inlier_rms) for homography and plane problems, so the pin on the loop
                         survives where neither /root/reference nor the prebuilt library exists;
  * passes.npz         — the oracle's (cost, J^T r, J^T J) at the start point of the reference's own test
                         scenarios (tests/ref_scenarios.py <- tests/unit/*_test.cpp) and its converged
                         parameters: a frozen copy of the checker, so a later change of the oracle or of the
                         CUDA path shows up against numbers that no longer move.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle_lib as O  # noqa: E402
import ref_scenarios as RS  # noqa: E402
from calibration_b200 import abi  # noqa: E402

SCENARIOS = {
    "intrinsics_noskew": lambda: RS.intrinsics_scenario(False),
    "intrinsics_skew": lambda: RS.intrinsics_scenario(True),
    "bundle_nodist": lambda: RS.bundle_scenario("nodist"),
    "bundle_distortion": lambda: RS.bundle_scenario("distortion"),
    "scheimpflug_intrinsics": lambda: RS.scheimpflug_scenario("intrinsics"),
    "scheimpflug_handeye": lambda: RS.scheimpflug_scenario("handeye"),
    "extrinsics_poses": lambda: RS.extrinsics_scenario("poses"),
    "extrinsics_all": lambda: RS.extrinsics_scenario("all"),
}


def ransac_ref_fixture():
    """Needs oracle/_ref (i.e. /root/reference): run in the build container only."""
    from calibration_b200 import synth
    assert O.ref_lib() is not None, "oracle/_ref/libref_ransac.so is needed to regenerate ransac_ref.npz"
    out = {}
    cases = []
    x, y, u, v, _ = synth.synth_ransac(seed=101, n_problems=6, n=120)
    variants = [dict(), dict(refit_on_inliers=0), dict(confidence=0.0, max_iters=60), dict(confidence=0.999999, max_iters=300),
                dict(min_inliers=110, max_iters=80), dict(thresh=0.8)]
    for p, kw in enumerate(variants):
        cases.append(("h", p, np.stack([x[p], y[p], u[p], v[p]]), dict(seed=500 + p, **kw)))
    _, d = O.homography_testdata(100, 0.0, 30, 7)       # homography_test.cpp:104-134
    cases.append(("h", 6, d.T.copy(), dict(thresh=1.0, min_inliers=90, seed=123)))
    px, py, pz, _ = synth.synth_plane_ransac(seed=102, n_problems=4, n=150)
    for p, kw in enumerate([dict(), dict(refit_on_inliers=0), dict(confidence=0.0, max_iters=50), dict(min_inliers=140, max_iters=90)]):
        cases.append(("p", p, np.stack([px[p], py[p], pz[p]]), dict(seed=900 + p, thresh=0.006, **kw)))
    _, xyz = O.plane_testdata()                           # planefit_test.cpp:22-75
    cases.append(("p", 4, xyz.T.copy(), dict(max_iters=2000, thresh=0.01, min_inliers=80, confidence=0.999)))
    for kind, p, data, kw in cases:
        opts = abi.RansacOptions.default(**kw)
        res, mask = (O.ref_ransac if kind == "h" else O.ref_ransac_plane)(*data, opts)
        key = f"{kind}{p}"
        out[key + "/data"] = data
        out[key + "/opts"] = np.array([opts.max_iters, opts.min_inliers, opts.thresh, opts.confidence, opts.seed, opts.refit_on_inliers], dtype=np.float64)
        out[key + "/success"] = np.array(res.success); out[key + "/iters"] = np.array(res.iters)
        out[key + "/mask"] = mask; out[key + "/rms"] = np.array(res.inlier_rms)
        out[key + "/model"] = np.array(res.hmtx if kind == "h" else res.plane)
    np.savez_compressed(os.path.join(HERE, "ransac_ref.npz"), **out)
    print("wrote", len(cases), "reference RANSAC cases")


def main():
    streams = {}
    for seed, n in [(1234567, 10), (1234567, 500), (123, 500), (123, 130), (42, 54), (7, 4), (2 ** 63 + 11, 1000), (0, 88)]:
        streams[f"{seed}:{n}"] = O.sample_stream(seed, n, 8, real=True).tolist()
    with open(os.path.join(HERE, "std_sample.json"), "w") as f:
        json.dump({"generator": "std::sample(iota(n), 4, std::mt19937_64(seed)) of libstdc++ (g++ 13), 8 successive draws", "streams": streams},
                  f, indent=1)
    out = {}
    for name, mk in SCENARIOS.items():
        prob, x0, _ = mk()
        c, g, H = O.refine_eval(prob, x0)
        x, res, _ = O.refine_solve(prob, x0, abi.OptimOptions.default(compute_covariance=0))
        out[name + "/x0"] = x0; out[name + "/cost"] = np.array(c); out[name + "/g"] = g; out[name + "/H"] = H
        out[name + "/x"] = x; out[name + "/final_cost"] = np.array(res.final_cost); out[name + "/iterations"] = np.array(res.iterations)
    np.savez_compressed(os.path.join(HERE, "passes.npz"), **out)
    print("wrote", len(streams), "sample streams and", len(SCENARIOS), "scenarios")
    if O.ref_lib() is not None:
        ransac_ref_fixture()


if __name__ == "__main__":
    main()
