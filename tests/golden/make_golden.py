"""Generates tests/golden/*.npz from the numpy/scipy oracle (oracle/pls_oracle.py, Lawson-Hanson via
scipy.optimize.nnls).  The reference is Julia and cannot run in this image, so these are
restatement values; the only reference-held fixture is the toy problem's opt ~ 0 and
predict(X) == y (test/runtests.jl:35-36).  Run:  python tests/golden/make_golden.py"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import pls_oracle as o

CASES = {
    # name: (N, M, K, seed, mixed_sign, rho, eta, tweak)
    "rand_a": (400, 12, 3, 11, False, 0.0, 0.0, None),
    "rand_b": (600, 18, 4, 12, True, 0.0, 1e-3, None),
    "rand_c_corr": (500, 16, 4, 13, True, 0.7, 0.0, None),
    "rand_d_overlap": (300, 10, 3, 14, True, 0.3, 1e-2, "overlap"),
}

def build(name):
    N, M, K, seed, mixed, rho, eta, tweak = CASES[name]
    X, y, P = o.make_synthetic(N, M, K, seed, mixed_sign=mixed, rho=rho)
    if tweak == "overlap":
        P[0, 1] = 1      # feature 0 in two groups
        P[4, :] = 0      # feature 4 in no group
    return X, y, P, eta

if __name__ == "__main__":
    r = o.fit_opt(o.TOY_X, o.TOY_Y, o.TOY_P, 0.0)
    r3 = o.fit_opt(o.TOY_X, o.TOY_Y, o.TOY_P, 1e-3)
    np.savez(os.path.join(HERE, "toy.npz"), X=o.TOY_X, y=o.TOY_Y, P=o.TOY_P, objs=r["objs"], b_best=r["b_best"],
             alpha_raw=r["alpha_raw"], alpha=r["alpha"], beta=r["beta"], t=r["t"],
             eta3_objs=r3["objs"], eta3_b_best=r3["b_best"], eta3_alpha=r3["alpha"], eta3_beta=r3["beta"],
             eta3_t=r3["t"], eta3_opt=r3["opt"])
    for name in CASES:
        X, y, P, eta = build(name)
        r = o.fit_opt(X, y, P, eta, return_all=True)
        Xo, Po = o.homogeneous_coords(X, P); Xa, ya = o.regularize_problem(Xo, y, Po, eta)
        alphas = np.array([o.opt_orthant(Xa, ya, Po, b)[1] for b in range(len(r["objs"]))])
        np.savez(os.path.join(HERE, name + ".npz"), X=X, y=y, P=P, eta=eta, objs=r["objs"], alphas=alphas,
                 b_best=r["b_best"], alpha=r["alpha"], beta=r["beta"], t=r["t"], opt=r["opt"])
        print(name, r["b_best"], r["opt"])
