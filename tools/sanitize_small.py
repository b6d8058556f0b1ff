"""Small Opt / BnB / Alt fits for compute-sanitizer (memcheck, racecheck, synccheck):
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
o, oc = g.load_oracle()
X, y, P = o.make_synthetic(600, 24, 5, seed=7, mixed_sign=True, rho=0.3)
ctx = pkg.Context(0)
which = sys.argv[1:] or ["opt", "optv5", "bnb", "alt"]
if "opt" in which:
    r = ctx.opt_fit(X, y, P, eta=1e-3, return_all=True)
    ref = oc.opt_fit(X, y, P, 1e-3)
    assert r["b_best"] == ref["b_best"] and abs(r["opt"] - ref["obj_best"]) <= 1e-9 * ref["obj_best"]
    print("opt ok", r["b_best"], r["opt"])
if "opth" in which:
    os.environ.update(PLS_K2_IMPL="v3", PLS_K3_QS="3", PLS_K3_T="256")
    r = ctx.opt_fit(X, y, P, eta=1e-3)
    for k in ("PLS_K2_IMPL", "PLS_K3_QS", "PLS_K3_T"):
        os.environ.pop(k)
    print("opt v3 hybrid ok", r["b_best"], r["opt"])
if "optv5" in which:
    # the two-tableau kernel (nnls5.cu) on a few long walks: per-orthant outputs forced onto it, then the winner-only
    # path; 128-thread walks and the 512-thread wide variant
    X5, y5, P5 = o.make_synthetic(700, 40, 7, seed=9, mixed_sign=True, rho=0.3)
    ref5 = oc.opt_fit(X5, y5, P5, 1e-3)
    for e in (dict(PLS_K5_GRID="3", PLS_K5_L="2", PLS_K5_VERIFY="5"), dict(PLS_K5_GRID="2", PLS_K5_L="3", PLS_K5_T="512", PLS_K5_NR="160")):
        os.environ.update(PLS_K2_IMPL="v5", **e)
        r = ctx.opt_fit(X5, y5, P5, eta=1e-3, return_all=True)
        w = ctx.opt_fit(X5, y5, P5, eta=1e-3)
        for k in ("PLS_K2_IMPL",) + tuple(e):
            os.environ.pop(k)
        assert r["stats"]["k2_variant"] == 5 and w["stats"]["k2_variant"] == 5
        assert r["b_best"] == ref5["b_best"] == w["b_best"] and abs(r["opt"] - ref5["obj_best"]) <= 1e-9 * ref5["obj_best"]
        assert np.allclose(r["objs"], ref5["objs"], rtol=1e-9, atol=1e-6 * np.linalg.norm(y5))
        print("opt v5 ok", e, r["b_best"], r["opt"], r["stats"]["k2_threads"])
if "bnb" in which:
    r = ctx.bnb_fit(X, y, P, eta=1e-3)
    print("bnb ok", r["opt"], r["nopen"])
if "alt" in which:
    b0 = pkg.draw_alt_starts(1, 25, 6, restarts=4)
    r = ctx.alt_fit(X, y, P, b0, eta=1e-3)
    print("alt ok", r["opt"], r["best_restart"], r["iters"])
