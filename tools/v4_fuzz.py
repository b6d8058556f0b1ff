"""Differential check of the two-level K2 kernel (v4) against the one-level kernel (v3) on random problems:
every orthant's objective and alpha (returnAllSolutions path), different shapes, correlations, group layouts,
eta, CTA counts and fast-group counts.   python tools/v4_fuzz.py [n_cases] [seed]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
from importlib import import_module
synth = import_module(g.PKG_NAME + ".synth")
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
KEYS = ("PLS_K2_IMPL", "PLS_K4_GRID", "PLS_K4_L", "PLS_K4_VERIFY", "PLS_K4_T", "PLS_K4_QS")
ctx = pkg.Context(0)
worst = dict(obj=0.0, alpha=0.0)
for case in range(n_cases):
    K = int(rng.integers(6, 13)); M = int(rng.integers(K + 2, 140)); N = int(rng.integers(M + 20, 4000))
    rho = float(rng.choice([0.0, 0.3, 0.7, 0.9])); eta = float(rng.choice([0.0, 1e-3, 1e-1]))
    mixed = bool(rng.integers(0, 2))
    X, y, P = synth.make_synthetic(N, M, K, int(rng.integers(1, 1 << 30)), mixed_sign=mixed, rho=rho)
    if rng.random() < 0.3:                      # overlapping groups / a group-less feature
        P = P.copy(); P[int(rng.integers(0, M)), int(rng.integers(0, K))] = 1; P[int(rng.integers(0, M)), :] = 0
        P = np.asfortranarray(P)
    for k in KEYS: os.environ.pop(k, None)
    os.environ["PLS_K2_IMPL"] = "v3"
    a = ctx.opt_fit(X, y, P, eta=eta, return_all=True)
    os.environ["PLS_K2_IMPL"] = "v4"
    os.environ["PLS_K4_GRID"] = str(int(rng.choice([1, 2, 5, 37, 600])))
    os.environ["PLS_K4_L"] = str(int(rng.integers(1, 7)))
    os.environ["PLS_K4_VERIFY"] = str(int(rng.choice([1, 7, 128])))
    if rng.random() < 0.3: os.environ["PLS_K4_T"] = str(int(rng.choice([64, 256])))
    if rng.random() < 0.2: os.environ["PLS_K4_QS"] = str(int(rng.integers(1, 30)))
    b = ctx.opt_fit(X, y, P, eta=eta, return_all=True)
    # winner-only fit under the same forced two-level settings: paired orthants (intercept free, 2^K problems) and,
    # where the KKT checks saw the tableau lose digits, the winner polished by the one-level kernel -- against
    # the one-level kernel's literal enumeration `a` at the parity tolerance
    c = ctx.opt_fit(X, y, P, eta=eta)
    # (an exact tie -- a group whose weights are all zero ties both of its signs -- may be broken either way by rounding:
    #  the winner is then compared with the one-level kernel's solution of the orthant it chose)
    yn0 = float(np.linalg.norm(y))
    cb = c["b_best"]
    pair_ok = ((cb == a["b_best"] or abs(a["objs"][cb] - a["objs"][a["b_best"]]) <= 1e-10 * yn0)
               and abs(c["opt"] - a["opt"]) <= 1e-9 * max(a["opt"], 1e-300) + 1e-10 * yn0
               and np.all(np.abs(c["alpha_raw"] - a["alphas"][cb]) <= 1e-9 * max(np.abs(a["alphas"][cb]).max(), 1e-300))
               and c["stats"]["nnls_problems"] * 2 == b["stats"]["nnls_problems"])
    yn = float(np.linalg.norm(y))
    eo = float(np.abs(a["objs"] - b["objs"]).max() / yn)
    sc = np.maximum(np.abs(a["alphas"]).max(axis=1, keepdims=True), 1e-300)
    ea = float((np.abs(a["alphas"] - b["alphas"]) / sc).max())
    near_tie = abs(a["objs"][a["b_best"]] - a["objs"][b["b_best"]]) <= 1e-10 * yn     # zero-weight groups: both signs tie, rounding picks
    ok = (a["b_best"] == b["b_best"] or near_tie) and eo <= 1e-6 and ea <= 1e-6 and pair_ok     # per-orthant outputs of a FORCED v4 run: diagnostic bound
    worst["obj"] = max(worst["obj"], eo); worst["alpha"] = max(worst["alpha"], ea)
    print(json.dumps(dict(case=case, N=N, M=M, K=K, rho=rho, eta=eta, mixed=mixed, env={k: os.environ.get(k) for k in KEYS if os.environ.get(k)},
                          obj_err=eo, alpha_err=ea, same_b=bool(a["b_best"] == b["b_best"]), pairs_equal=bool(pair_ok), rebuilds=b["stats"]["rebuilds"], drift_restarts=b["stats"]["spills"], v3_rebuilds=a["stats"]["rebuilds"], polished=c["stats"]["rebuilds"], ok=bool(ok))), flush=True)
    if not ok:
        sys.exit(1)
print(json.dumps(dict(cases=n_cases, worst=worst, result="all equal")))
