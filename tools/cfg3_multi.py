"""BASELINE configs[2]: synthetic N=1M, M=512, K=24, fit(Opt) -- 2^25 = 33 554 432 orthant NNLS -- driven by ONE
process over all visible GPUs (pls_create with a device list).  Prints wall times and checks the KKT
conditions of the winner against a numpy Gram matrix.   python tools/cfg3_multi.py [N] [M] [K] [ngpus]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
from importlib import import_module
synth = import_module(g.PKG_NAME + ".synth")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 512
K = int(sys.argv[3]) if len(sys.argv) > 3 else 24
ng = int(sys.argv[4]) if len(sys.argv) > 4 else pkg._abi.lib.pls_device_count()
t0 = time.perf_counter()
X, y, P = synth.make_synthetic(N, M, K, 20240417)
t_gen = time.perf_counter() - t0
ctx = pkg.Context(list(range(ng)) if ng > 1 else 0)
t0 = time.perf_counter(); ctx.load(X, y, P, eta=0.0); t_load = time.perf_counter() - t0
t0 = time.perf_counter(); r = ctx.opt_fit_resident(); t_fit = time.perf_counter() - t0
st = r["stats"]
total = 1 << (K + 1)
rec = dict(N=N, M=M, K=K, gpus=ng, orthants=total, s_generate=t_gen, s_load=t_load, s_fit_resident=t_fit,
           solves_per_s=total / t_fit, ms_gram=st["ms_gram"], ms_nnls_max_over_gpus=st["ms_nnls"], ms_recompute=st["ms_recompute"],
           b_best=r["b_best"], opt=r["opt"], pivots=st["pivots"], grad_evals=st["grad_evals"], rebuilds=st["rebuilds"],
           gram_tflops=st["gram_flops"] / (st["ms_gram"] * 1e-3) / 1e12)
# KKT check of the winner on a host Gram matrix
Xo = np.hstack([X, np.ones((N, 1))])
G = Xo.T @ Xo; c = Xo.T @ y
Po = np.zeros((M + 1, K + 1)); Po[:M, :K] = P; Po[M, K] = 1
beta = np.array([2 * ((r["b_best"] >> k) & 1) - 1 for k in range(K + 1)], dtype=float)
d = Po @ beta; w = d * r["alpha_raw"]
grad = c - G @ w
passive = r["alpha_raw"] > 0
rec["kkt_max_passive_grad_rel"] = float(np.abs(grad[passive]).max() / np.abs(c).max())
rec["kkt_max_active_violation_rel"] = float(max(0.0, (d * grad)[~passive].max()) / np.abs(c).max()) if (~passive).any() else 0.0
rec["obj_check_rel"] = float(abs(np.sqrt(np.sum((Xo @ w - y) ** 2)) - r["opt"]) / r["opt"])
print(json.dumps(rec), flush=True)
