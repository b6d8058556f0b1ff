"""fit(Alt) with R batched restarts on synthetic workloads: wall time, restarts/s, iterations.
   python tools/alt_bench.py N M K R [eta]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
from importlib import import_module
synth = import_module(g.PKG_NAME + ".synth")
N, M, K, R = (int(a) for a in sys.argv[1:5])
eta = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
X, y, P = synth.make_synthetic(N, M, K, 20240418, mixed_sign=False)
_ng = int(os.environ.get('NGPU', '1'))
ctx = pkg.Context(list(range(_ng)) if _ng > 1 else 0)
ctx.load(X, y, P, eta=eta)
beta0 = pkg.draw_alt_starts(20240418, M + 1, K + 1, R)
for rep in range(2):
    t0 = time.perf_counter(); r = ctx.alt_fit(None, None, None, beta0, eps=1e-6, T=100, resident=True); dt = time.perf_counter() - t0
st = r["stats"]
fin = r["all_obj"][np.isfinite(r["all_obj"])]
rec = dict(N=N, M=M, K=K, R=R, eta=eta, s=dt, opt=r["opt"], best_restart=r["best_restart"], iters_best=r["iters"],
           max_iters=st["waves"], ms_gram=st["ms_gram"], ms_alt=st["ms_nnls"], restarts_per_s=R / (st["ms_nnls"] * 1e-3),
           pivots=st["pivots"], grad_evals=st["grad_evals"], failed=int(R - len(fin)),
           obj_min=float(fin.min()), obj_median=float(np.median(fin)), obj_max=float(fin.max()),
           n_at_best=int(np.sum(fin <= fin.min() * (1 + 1e-9))))
print(json.dumps(rec), flush=True)
