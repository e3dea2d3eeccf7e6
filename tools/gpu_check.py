"""Quick GPU-vs-oracle comparison used while developing (not a test)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
import ref_scenarios as RS
from calibration_b200 import abi, capi, synth, geometry as G


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def check_eval(name, prob, x):
    h = capi.RefineHandle(prob)
    c_o, g_o, H_o = O.refine_eval(prob, x)
    c_g, g_g, H_g = h.eval(x)
    cc = h.cost(x)
    print(f"[eval] {name}: n_tan={len(g_o)} cost {c_o:.12e} vs {c_g:.12e} (cost-only {cc:.12e}) "
          f"rel cost {abs(c_o-c_g)/abs(c_o):.2e} g {relerr(g_g, g_o):.2e} H {relerr(H_g, H_o):.2e}")
    if relerr(H_g, H_o) > 1e-9:
        n = len(g_o)
        bad = np.argwhere(np.abs(H_g - H_o) > 1e-9 * np.abs(H_o).max())
        print("   first mismatches:", bad[:12].tolist())
        print("   g diff idx:", np.argwhere(np.abs(g_g - g_o) > 1e-9 * np.abs(g_o).max()).ravel()[:20].tolist())
    h.close()


def check_solve(name, prob, x0, **kw):
    h = capi.RefineHandle(prob)
    opts = abi.OptimOptions.default(**kw)
    t = time.time(); xo, ro, covo = O.refine_solve(prob, x0, opts); to = time.time() - t
    t = time.time(); xg, rg, covg = h.solve(x0, opts); tg = time.time() - t
    print(f"[solve] {name}: oracle {ro.report.decode()} ({to:.3f}s)\n          gpu    {rg.report.decode()} ({tg:.3f}s)")
    print(f"          max |dx| {np.abs(xo-xg).max():.3e} rel {relerr(xg, xo):.3e} cov ok {ro.covariance_ok}/{rg.covariance_ok}"
          + (f" cov rel {relerr(covg, covo):.2e}" if (covo is not None and covg is not None and ro.covariance_ok and rg.covariance_ok) else ""))
    h.close()


if __name__ == "__main__":
    print("devices:", capi.device_count())
    for skew in (False, True):
        prob, x0, _ = RS.intrinsics_scenario(skew)
        check_eval(f"ref intrinsics skew={skew}", prob, x0)
    for k in ("nodist", "distortion"):
        prob, x0, _ = RS.bundle_scenario(k)
        check_eval(f"ref bundle {k}", prob, x0)
    for w in ("intrinsics", "handeye"):
        prob, x0, _ = RS.scheimpflug_scenario(w)
        check_eval(f"ref scheimpflug {w}", prob, x0)
    for w in ("poses", "all"):
        prob, x0, _ = RS.extrinsics_scenario(w)
        check_eval(f"ref extrinsics {w}", prob, x0)
    prob, x0, _ = synth.make_intrinsics()
    check_eval("C1 intrinsics", prob, x0)
    prob, x0, _ = synth.make_intrinsics(model=abi.MODEL_SCHEIMPFLUG_BC5, optimize_skew=True)
    check_eval("C1 scheimpflug skew", prob, x0)
    prob, x0, _ = synth.make_extrinsics(n_views=50, drop_fraction=0.2)
    check_eval("extrinsics 50 views", prob, x0)
    prob, x0, _ = synth.make_extrinsics(n_cams=3, n_views=40, optimize_intrinsics=False)
    check_eval("extrinsics fixed intr", prob, x0)
    prob, x0, _ = synth.make_bundle(n_cams=4, n_poses=60)
    check_eval("bundle 4x60", prob, x0)
    prob, x0, _ = synth.make_bundle(n_cams=2, n_poses=60, optimize_intrinsics=False)
    check_eval("bundle fixed intr", prob, x0)
    prob, x0, _ = synth.make_bundle(n_cams=2, n_poses=60, model=abi.MODEL_SCHEIMPFLUG_BC5)
    check_eval("bundle scheimpflug", prob, x0)
    # solves
    for skew in (False, True):
        prob, x0, _ = RS.intrinsics_scenario(skew)
        check_solve(f"ref intrinsics skew={skew}", prob, x0)
    prob, x0, _ = RS.bundle_scenario("distortion")
    check_solve("ref bundle distortion", prob, x0)
    prob, x0, _ = RS.extrinsics_scenario("all")
    check_solve("ref extrinsics all", prob, x0)
    prob, x0, _ = synth.make_intrinsics()
    check_solve("C1 intrinsics noisy", prob, x0)
    prob, x0, _ = synth.make_extrinsics(n_views=100)
    check_solve("extrinsics 100 views noisy", prob, x0)
    prob, x0, _ = synth.make_bundle(n_cams=4, n_poses=200)
    check_solve("bundle 4x200 noisy", prob, x0)
