"""fit(BnB) on synthetic mixed-sign workloads: wall time, nodes visited, waves; compared with Opt where feasible.
   python tools/bnb_bench.py N M K [eta] [--opt]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
from importlib import import_module
synth = import_module(g.PKG_NAME + ".synth")
N, M, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
eta = float(sys.argv[4]) if len(sys.argv) > 4 and not sys.argv[4].startswith("--") else 0.0
flip = float(os.environ.get("BNB_FLIP", "1.0"))   # 1.0: fully mixed-sign w ~ N(0,1); f < 1: sign-consistent groups with a
if flip >= 1.0:                                    # fraction f of the weights flipped and shrunk (the tree branches, but decisively)
    X, y, P = synth.make_synthetic(N, M, K, 20240419, mixed_sign=True)
else:
    rng = np.random.default_rng(20240419)
    X = np.asfortranarray(rng.standard_normal((N, M)))
    gidx = (np.arange(M) * K) // M
    sgn = rng.choice([-1.0, 1.0], size=K)
    w = np.abs(rng.standard_normal(M)) * sgn[gidx]
    fl = rng.random(M) < flip
    w[fl] *= -0.3
    y = X @ w + 0.5 + rng.standard_normal(N)
    P = np.zeros((M, K), dtype=np.int64); P[np.arange(M), gidx] = 1
    P = np.asfortranarray(P)
_ng = int(os.environ.get('NGPU', '1'))
ctx = pkg.Context(list(range(_ng)) if _ng > 1 else 0)
ctx.load(X, y, P, eta=eta)
for rep in range(int(os.environ.get("BNB_REPS", "2"))):
    t0 = time.perf_counter(); r = ctx.bnb_fit_resident(); dt = time.perf_counter() - t0
st = r["stats"]
rec = dict(N=N, M=M, K=K, eta=eta, flip=flip, s=dt, opt=r["opt"], nopen=r["nopen"], waves=st["waves"], max_open=st["max_open"],
           ms_gram=st["ms_gram"], ms_bnb=st["ms_nnls"], nodes_per_s=r["nopen"] / (st["ms_nnls"] * 1e-3), pivots=st["pivots"],
           grad_evals=st["grad_evals"], full_tree=2 ** (K + 2) - 1)
if "--opt" in sys.argv:
    t0 = time.perf_counter(); ro = ctx.opt_fit_resident(); rec["opt_s"] = time.perf_counter() - t0
    rec["opt_obj"] = ro["opt"]; rec["same_opt"] = bool(abs(ro["opt"] - r["opt"]) <= 1e-9 * ro["opt"])
print(json.dumps(rec), flush=True)
