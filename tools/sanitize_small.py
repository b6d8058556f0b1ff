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
which = sys.argv[1:] or ["opt", "optv2", "bnb", "alt"]
if "opt" in which:
    r = ctx.opt_fit(X, y, P, eta=1e-3, return_all=True)
    ref = oc.opt_fit(X, y, P, 1e-3)
    assert r["b_best"] == ref["b_best"] and abs(r["opt"] - ref["obj_best"]) <= 1e-9 * ref["obj_best"]
    print("opt ok", r["b_best"], r["opt"])
if "optv2" in which:
    os.environ["PLS_K2_IMPL"] = "v2"
    r = ctx.opt_fit(X, y, P, eta=1e-3)
    os.environ.pop("PLS_K2_IMPL")
    print("opt v2 ok", r["b_best"], r["opt"])
if "opth" in which:
    os.environ.update(PLS_K2_IMPL="v3", PLS_K3_QS="3", PLS_K3_T="256")
    r = ctx.opt_fit(X, y, P, eta=1e-3)
    for k in ("PLS_K2_IMPL", "PLS_K3_QS", "PLS_K3_T"):
        os.environ.pop(k)
    print("opt v3 hybrid ok", r["b_best"], r["opt"])
if "bnb" in which:
    r = ctx.bnb_fit(X, y, P, eta=1e-3)
    print("bnb ok", r["opt"], r["nopen"])
if "alt" in which:
    b0 = pkg.draw_alt_starts(1, 25, 6, restarts=4)
    r = ctx.alt_fit(X, y, P, b0, eta=1e-3)
    print("alt ok", r["opt"], r["best_restart"], r["iters"])
