"""Consistency of the single-process multi-GPU context (all visible GPUs) with the one-GPU context on one problem:
Opt (paired and literal), Alt, BnB, predict.   python tools/multi_check.py [n_gpus]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
from importlib import import_module
synth = import_module(g.PKG_NAME + ".synth")
ng = int(sys.argv[1]) if len(sys.argv) > 1 else pkg._abi.lib.pls_device_count()
X, y, P = synth.make_synthetic(30011, 96, 11, 4242, mixed_sign=True, rho=0.3)
eta = 1e-3
one = pkg.Context(0); many = pkg.Context(list(range(ng)))
rec = dict(gpus=ng)
a = one.opt_fit(X, y, P, eta=eta); b = many.opt_fit(X, y, P, eta=eta)
rec["opt_pairs"] = dict(same_b=bool(a["b_best"] == b["b_best"]), opt_rel=abs(a["opt"] - b["opt"]) / a["opt"],
                        alpha_rel=float(np.abs(a["alpha_raw"] - b["alpha_raw"]).max() / np.abs(a["alpha_raw"]).max()), problems=b["stats"]["nnls_problems"])
c = many.opt_fit(X, y, P, eta=eta, flags=pkg._abi.PLS_FLAG_ENUMERATE_INTERCEPT)
rec["opt_literal"] = dict(same_b=bool(a["b_best"] == c["b_best"]), opt_rel=abs(a["opt"] - c["opt"]) / a["opt"], problems=c["stats"]["nnls_problems"])
w = a["alpha_raw"] * np.array([pkg._abi.lib.pls_version() * 0 + 1.0] * len(a["alpha_raw"]))
beta0 = pkg.draw_alt_starts(7, X.shape[1] + 1, P.shape[1] + 1, restarts=37)
ra = one.alt_fit(X, y, P, beta0, eta=eta); rb = many.alt_fit(X, y, P, beta0, eta=eta)
rec["alt"] = dict(same_restart=bool(ra["best_restart"] == rb["best_restart"]), opt_rel=abs(ra["opt"] - rb["opt"]) / ra["opt"])
ba = one.bnb_fit(X, y, P, eta=eta); bb = many.bnb_fit(X, y, P, eta=eta)
rec["bnb"] = dict(opt_rel=abs(ba["opt"] - bb["opt"]) / ba["opt"], alpha_rel=float(np.abs(ba["alpha_signed"] - bb["alpha_signed"]).max() / np.abs(ba["alpha_signed"]).max()),
                  nopen_one=ba["nopen"], nopen_many=bb["nopen"], equals_opt=abs(ba["opt"] - a["opt"]) / a["opt"])
wv = np.random.default_rng(0).standard_normal(X.shape[1] + 1)
many.load(X, y, P, eta=eta)
yh = many.predict_resident(wv, len(y)); ref = X @ wv[:-1] + wv[-1]
rec["predict_rel"] = float(np.abs(yh - ref).max() / np.abs(ref).max())
ok = (rec["opt_pairs"]["same_b"] and rec["opt_pairs"]["opt_rel"] < 1e-9 and rec["opt_pairs"]["alpha_rel"] < 1e-9 and rec["opt_literal"]["same_b"]
      and rec["alt"]["same_restart"] and rec["alt"]["opt_rel"] < 1e-8 and rec["bnb"]["opt_rel"] < 1e-9 and rec["bnb"]["alpha_rel"] < 1e-9 and rec["predict_rel"] < 1e-12)
rec["ok"] = bool(ok)
print(json.dumps(rec))
sys.exit(0 if ok else 1)
