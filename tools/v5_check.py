"""Quick differential check + timing of the v5 K2 kernel (nnls5.cu) against the one-level kernel (v3) and v4.
   python tools/v5_check.py [small|cfg2|k20|all]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
from importlib import import_module
synth = import_module(g.PKG_NAME + ".synth")
KEYS = ("PLS_K2_IMPL", "PLS_K2_NO_V5", "PLS_K5_SEED")
SMALL_KEYS = ("PLS_K5_GRID", "PLS_K5_L", "PLS_K5_VERIFY")


def env(**kw):
    for k in KEYS + (SMALL_KEYS if any(q in kw for q in SMALL_KEYS) or os.environ.get("_V5_SMALL") else ()):
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ[k] = str(v)


def small(ctx):
    bad = 0
    os.environ["_V5_SMALL"] = "1"
    cases = [(900, 40, 6, 1e-3, True, 0.3, dict(PLS_K5_GRID=3, PLS_K5_L=2, PLS_K5_VERIFY=5)),
             (900, 40, 6, 1e-3, True, 0.3, dict(PLS_K5_GRID=1, PLS_K5_L=1, PLS_K5_VERIFY=4)),
             (2000, 64, 8, 1e-3, True, 0.0, dict(PLS_K5_GRID=5, PLS_K5_L=3)),
             (1500, 96, 10, 0.0, False, 0.5, dict(PLS_K5_GRID=16, PLS_K5_L=3)),
             (3000, 150, 11, 1e-3, True, 0.2, dict(PLS_K5_GRID=64)),
             (600, 230, 9, 1e-3, False, 0.0, dict(PLS_K5_GRID=32))]
    for (N, M, K, eta, mixed, rho, e) in cases:
        X, y, P = synth.make_synthetic(N, M, K, 1000 + M, mixed_sign=mixed, rho=rho)
        env(PLS_K2_IMPL="v3")
        a = ctx.opt_fit(X, y, P, eta=eta, return_all=True)
        env(PLS_K2_IMPL="v5", **e)
        t0 = time.perf_counter()
        b = ctx.opt_fit(X, y, P, eta=eta, return_all=True)
        dt = time.perf_counter() - t0
        sc = np.abs(a["alphas"]).max(axis=1, keepdims=True) + 1e-300
        da = float((np.abs(a["alphas"] - b["alphas"]) / sc).max())
        do = float(np.nanmax(np.abs(a["objs"] - b["objs"])) / np.linalg.norm(y))
        st = b["stats"]
        ok = b["b_best"] == a["b_best"] and da < 1e-9 and do < 1e-9 and st["k2_variant"] == 5
        bad += not ok
        print(json.dumps(dict(case=[N, M, K, eta, mixed, rho], env=e, ok=bool(ok), variant=st["k2_variant"], same_b=b["b_best"] == a["b_best"], max_dalpha=da, max_dobj=do,
                              sweeps=st["pivots"], streams=st["grad_evals"], iters=st["bpp_iters"], rebuilds=st["rebuilds"], blocked=st["blocked"], spills=st["spills"],
                              drift=st["k2_max_drift"], ms=st["ms_nnls"], s=dt)), flush=True)
        env()
        c = ctx.opt_fit(X, y, P, eta=eta)           # default dispatch, winner only (pairs)
        ok2 = c["b_best"] == a["b_best"] and abs(c["opt"] - a["opt"]) <= 1e-9 * a["opt"]
        bad += not ok2
        print(json.dumps(dict(default_winner_only=True, ok=bool(ok2), variant=c["stats"]["k2_variant"], b=c["b_best"], b_ref=a["b_best"])), flush=True)
    return bad


def big(ctx, name, reps=3):
    N, M, K, eta, seed, mixed = synth.CONFIGS[name]
    X, y, P = synth.make_synthetic(N, M, K, seed, mixed_sign=mixed)
    ctx.load(X, y, P, eta=eta)
    out = {}
    for label, e in (("v5", {}), ("v4", dict(PLS_K2_NO_V5=1))):
        env(**e)
        for _ in range(reps):
            r = ctx.opt_fit_resident()
        st = r["stats"]
        out[label] = dict(variant=st["k2_variant"], ms_nnls=st["ms_nnls"], b=r["b_best"], opt=r["opt"], sweeps=st["pivots"], streams=st["grad_evals"], sum_s=st["sum_p"],
                          iters=st["bpp_iters"], rebuilds=st["rebuilds"], blocked=st["blocked"], spills=st["spills"], drift=st["k2_max_drift"], grid=st["k2_grid"], occ=st["k2_ctas_per_sm"])
        print(name, label, json.dumps(out[label]), flush=True)
    same = out["v5"]["b"] == out["v4"]["b"] and abs(out["v5"]["opt"] - out["v4"]["opt"]) <= 1e-9 * out["v4"]["opt"]
    print(name, "same winner:", same, flush=True)
    return 0 if same else 1


def shard(ctx, name, nshard, reps=3):
    """one rank's share of a strong-scaling run: the first 1/nshard of the sign patterns (pairs)"""
    N, M, K, eta, seed, mixed = synth.CONFIGS[name]
    X, y, P = synth.make_synthetic(N, M, K, seed, mixed_sign=mixed)
    ctx.load(X, y, P, eta=eta); ctx.gram_build(); ctx.gram_finalize()
    for label, e in (("v5", {}), ("v5_noseed", dict(PLS_K5_SEED=0)), ("v4", dict(PLS_K2_NO_V5=1))):
        env(**e)
        for _ in range(reps):
            r = ctx.opt_solve_pairs(0, (1 << K) // nshard)
        st = ctx.stats()
        print(name, f"1/{nshard}", label, json.dumps(dict(variant=st["k2_variant"], ms_nnls=st["ms_nnls"], b=r["b_best"], sweeps=st["pivots"], grid=st["k2_grid"])), flush=True)
    return 0


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    ctx = pkg.Context(0)
    bad = 0
    if what in ("small", "all"):
        bad += small(ctx)
    if what in ("cfg2", "all"):
        bad += big(ctx, "cfg2")
    if what in ("k20", "all"):
        bad += big(ctx, "k20_m200")
    if what in synth.CONFIGS and what not in ("cfg2",):
        bad += big(ctx, what, reps=2)
    if what == "cold":
        N, M, K, eta, seed, mixed = synth.CONFIGS["k20_m200"]
        X, y, P = synth.make_synthetic(N, M, K, seed, mixed_sign=mixed)
        ctx.load(X, y, P, eta=eta); ctx.gram_build(); ctx.gram_finalize()
        env(PLS_K2_IMPL="v5", PLS_K5_SEED=0)
        for cnt in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
            for _ in range(3):
                r = ctx.opt_solve_pairs(0, cnt)
            st = ctx.stats()
            print("cold", cnt, json.dumps(dict(ms=st["ms_nnls"], grid=st["k2_grid"], sweeps=st["pivots"], streams=st["grad_evals"], iters=st["bpp_iters"])), flush=True)
    if what.startswith("shard"):
        bad += shard(ctx, "k20_m200", int(what[5:]))
    print("RESULT", "ok" if bad == 0 else f"{bad} FAILED")
    sys.exit(1 if bad else 0)
