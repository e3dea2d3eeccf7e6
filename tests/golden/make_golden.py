"""Generates the committed golden fixtures of tests/golden/ (run from the repository root:
`python tests/golden/make_golden.py`).

The reference itself cannot be compiled or imported in this image (Ceres / Eigen / Boost.PFR absent, see
DESIGN.md §2), so the fixtures come from the two things that CAN be run here:
  * std_sample.json    — the REAL libstdc++ std::sample / std::mt19937_64 of this toolchain (the reference's
                         hypothesis stream, include/calib/estimation/common/ransac.h:135,144-145), drawn by
                         oracle/ransac.cpp::orc_sample_stream_libstdcxx, plus the known answers recorded in
                         SURVEY Appendix C;
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


if __name__ == "__main__":
    main()
