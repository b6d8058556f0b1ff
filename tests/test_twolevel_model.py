"""CPU test of the algebra behind the two-level K2 kernel (partitionedls.jl_b200/csrc/nnls4.cu):
the swept tableau, forward / reverse sweeps in blocks of <= 8 variables, the implied weights of
committed variables and the reduced objective -- checked, orthant by orthant along a Gray walk,
against the oracle's per-orthant Lawson-Hanson solutions (src/PartitionedLSOpt.jl:85-94).
The model is tools/proto_twolevel.py (numpy, design-time tool); the CUDA kernel follows it step
for step and is compared with the oracle itself in tests/test_gpu_parity.py (variants v4, v4q)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("shape", [(400, 24, 5, 1e-3, 0.0, 2), (300, 18, 4, 0.0, 0.6, 1), (500, 30, 6, 1e-2, 0.3, 3)])
def test_tableau_walk_matches_oracle(oracle, shape):
    o, oc = oracle
    from proto_twolevel import TwoLevel
    N, M, K, eta, rho, l_inner = shape
    X, y, P = o.make_synthetic(N, M, K, seed=7 + M, mixed_sign=True, rho=rho)
    ref = oc.opt_fit(X, y, P, eta)                       # all 2^(K+1) orthants, data-space Lawson-Hanson
    Xo, Po = o.homogeneous_coords(X, P)
    G = Xo.T @ Xo + eta * (Po @ Po.T)
    c = Xo.T @ y
    yy = float(y @ y)
    Mp, Kp = Po.shape
    gmask = np.array([sum(1 << k for k in range(Kp) if Po[m, k]) for m in range(Mp)], dtype=np.int64)
    tl = TwoLevel(G, c, yy, gmask, l_inner)
    n_orth = 1 << Kp
    for i in range(n_orth):
        b = i ^ (i >> 1)
        beta = np.array([2 * ((b >> k) & 1) - 1 for k in range(Kp)], float)
        d = Po @ beta
        tl.last_unswept = []
        w, obj2 = tl.solve(np.sign(d))
        # the tableau stays symmetric and its swept block is -inv(G_OO)
        assert np.abs(tl.T - tl.T.T).max() <= 1e-9 * np.abs(tl.T).max()
        O = np.flatnonzero(tl.swp)
        if len(O):
            HO = -tl.T[np.ix_(O, O)]
            assert np.abs(HO @ G[np.ix_(O, O)] - np.eye(len(O))).max() < 1e-8
        alpha = np.where(d != 0, w / np.where(d != 0, d, 1.0), 0.0)
        sc = max(np.abs(ref["alphas"][b]).max(), 1e-300)
        assert np.abs(alpha - ref["alphas"][b]).max() <= 1e-9 * sc
        assert abs(np.sqrt(max(obj2, 0.0)) - ref["objs"][b]) <= 1e-9 * ref["objs"][b] + 1e-6 * np.sqrt(yy)
        tl.commit()
    assert tl.n_sweep_blocks > 0 and tl.n_unsweep_blocks > 0   # both tableau operations were exercised


@pytest.mark.parametrize("shape", [(300, 14, 4, 1e-3, 0.4, 0.8), (200, 9, 3, 0.0, 0.0, -1.5), (250, 12, 5, 1e-2, 0.6, 0.0)])
def test_free_intercept_resolves_both_orthants_of_a_pair(oracle, shape):
    """The claim behind the paired-orthant mode: for every sign pattern p of the K user groups, the
    NNLS problem with the intercept sign left FREE has the optimum of the better of the two reference orthants p and
    p + 2^K (Opt.jl:85-96 enumerates both), and the sign of its intercept weight tells which one.  Checked with the
    oracle's Lawson-Hanson: free intercept = the intercept column entered twice, as +1 and -1 (both >= 0)."""
    o, oc = oracle
    N, M, K, eta, rho, shift = shape
    X, y, P = o.make_synthetic(N, M, K, seed=11 * M, mixed_sign=True, rho=rho)
    y = y - 0.5 + shift
    ref = oc.opt_fit(X, y, P, eta)
    Xo, Po = o.homogeneous_coords(X, P)
    Xa, ya = o.regularize_problem(Xo, y, Po, eta)
    for p in range(1 << K):
        beta = o.index_to_beta(p, K + 1).astype(float)          # top bit 0: intercept sign -1 (ignored below)
        d = Po @ beta
        A = np.hstack([Xa[:, :M] * d[None, :M], Xa[:, M:M + 1], -Xa[:, M:M + 1]])
        a = o.nonneg_lsq(A, ya)
        t = a[M] - a[M + 1]
        obj = float(np.linalg.norm(A @ a - ya))
        lo, hi = ref["objs"][p], ref["objs"][p + (1 << K)]
        assert abs(obj - min(lo, hi)) <= 1e-9 * max(min(lo, hi), 1e-12) + 1e-12 * np.linalg.norm(y)
        if abs(lo - hi) > 1e-9 * max(lo, hi):
            assert (t > 0) == (hi < lo)                         # the intercept's sign names the better orthant
        b_full = p | ((1 << K) if t > 0 else 0)
        alpha = np.concatenate([a[:M], [abs(t)]])
        sc = max(np.abs(ref["alphas"][b_full]).max(), 1e-300)
        assert np.abs(alpha - ref["alphas"][b_full]).max() <= 1e-8 * sc
