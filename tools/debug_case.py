"""Runs one saved problem (npz: X, y, P, eta) through K2 variants with per-orthant outputs and reports the orthants
that failed (NaN objective), the solver counters, and the worst disagreement with the C oracle."""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
A = pkg._abi
o, oc = g.load_oracle(); oc.build()
d = np.load(sys.argv[1]); X, y, P, eta = np.asfortranarray(d["X"]), d["y"], np.asfortranarray(d["P"]), float(d["eta"])
ctx = pkg.Context(0)
ctx.load(X, y, P, eta=eta); ctx.gram_build(); ctx.gram_finalize()
N, M = X.shape; K = P.shape[1]; total = 1 << (K + 1); Mp = M + 1
ref = oc.opt_fit(X, y, P, eta)
for impl in sys.argv[2:]:
    for k in ("PLS_K2_IMPL", "PLS_K4_GRID", "PLS_K4_L", "PLS_K3_CHAIN"): os.environ.pop(k, None)
    for kv in impl.split(","):
        if kv:
            k, v = kv.split("="); os.environ[k] = v
    alpha = np.zeros(Mp); b = C.c_int64(); obj = C.c_double()
    all_obj = np.zeros(total); all_alpha = np.zeros((total, Mp))
    rc = A.lib.pls_opt_solve_range(ctx._h, 0, total, A._d(alpha), C.byref(b), C.byref(obj), A._d(all_obj), A._d(all_alpha))
    st = ctx.stats()
    bad = np.flatnonzero(~np.isfinite(all_obj))
    good = np.isfinite(all_obj)
    sc = np.maximum(np.abs(ref["alphas"]).max(axis=1), 1e-300)
    ea = (np.abs(all_alpha - ref["alphas"]).max(axis=1) / sc)[good].max() if good.any() else None
    eo = (np.abs(all_obj - ref["objs"])[good] / np.linalg.norm(y)).max() if good.any() else None
    print(json.dumps(dict(impl=impl, rc=rc, failed_orthants=bad.tolist()[:16], n_failed=int(len(bad)), max_alpha_err_vs_oracle=ea, max_obj_err_vs_oracle=eo,
                          passive_at_failed=[int(np.count_nonzero(ref["alphas"][q])) for q in bad[:8]],
                          counters={k: st[k] for k in ("pivots", "grad_evals", "bpp_iters", "rebuilds", "blocked", "spills")})), flush=True)
