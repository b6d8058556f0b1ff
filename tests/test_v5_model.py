"""CPU test of the algebra behind the default K2 kernel (partitionedls.jl_b200/csrc/nnls5.cu): two swept
tableaus per Gray walk -- T1 = sweep([G c; c' yy], O) changed only by folds, T2 = sweep(T1[Rb, Rb], S) over a window --
with block principal pivoting on T2, the streaming check, joins, folds, and the round-2 additions: window COMPACTION
(untoggled variables leave the window without recomputation) and the FUSED fold of the cold solve (every toggled
variable into T1 in one pass, T1 += Z' T1[S, :]).  The model is tools/proto_v5.py (numpy, design-time tool); every
orthant of a Gray walk is checked against the oracle's per-orthant Lawson-Hanson solutions
(src/PartitionedLSOpt.jl:85-94).  The CUDA kernel is compared with the oracle itself in tests/test_gpu_parity.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _setup(o, oc, N, M, K, eta, rho, seed):
    X, y, P = o.make_synthetic(N, M, K, seed=seed, mixed_sign=True, rho=rho)
    ref = oc.opt_fit(X, y, P, eta)                       # all 2^(K+1) orthants, data-space Lawson-Hanson
    Xo, Po = o.homogeneous_coords(X, P)
    G = Xo.T @ Xo + eta * (Po @ Po.T)
    c = Xo.T @ y
    yy = float(y @ y)
    Mp, Kp = Po.shape
    gmask = np.array([sum(1 << k for k in range(Kp) if Po[m, k]) for m in range(Mp)], dtype=np.int64)
    return ref, G, c, yy, Po, gmask


@pytest.mark.parametrize("shape", [(400, 24, 5, 1e-3, 0.0, 2, 14), (300, 18, 4, 0.0, 0.6, 1, 10), (500, 30, 6, 1e-2, 0.3, 2, 16)])
def test_two_tableau_walk_matches_oracle(oracle, shape):
    o, oc = oracle
    from proto_v5 import V5
    N, M, K, eta, rho, l, capR = shape
    ref, G, c, yy, Po, gmask = _setup(o, oc, N, M, K, eta, rho, seed=17 + M)
    Mp, Kp = Po.shape
    s5 = V5(G, c, yy, gmask, 0, capR)                    # cold solve: no fast groups
    for i in range(1 << Kp):
        b = i ^ (i >> 1)
        beta = np.array([2 * ((b >> k) & 1) - 1 for k in range(Kp)], float)
        d = Po @ beta
        v = s5.solve(np.sign(d), cold=(i == 0))
        w = s5.full_weights(v)
        alpha = np.where(d != 0, w / np.where(d != 0, d, 1.0), 0.0)
        sc = max(np.abs(ref["alphas"][b]).max(), 1e-300)
        assert np.abs(alpha - ref["alphas"][b]).max() <= 1e-9 * sc
        assert abs(np.sqrt(max(v[Mp], 0.0)) - ref["objs"][b]) <= 1e-9 * ref["objs"][b] + 1e-6 * np.sqrt(yy)
        # T1 stays symmetric and its swept block is -inv(G_OO)
        assert np.abs(s5.T1 - s5.T1.T).max() <= 1e-9 * np.abs(s5.T1).max()
        O = np.flatnonzero(s5.inO)
        if len(O):
            assert np.abs(-s5.T1[np.ix_(O, O)] @ G[np.ix_(O, O)] - np.eye(len(O))).max() < 1e-8
        if i == 0:                                       # after the cold solve: everything toggled goes into T1 in one
            s5.compact()                                 # pass, then the fast groups come into the window
            s5.fold_fused()
            s5.set_fast(l)
        else:
            fb = (i & -i).bit_length() - 1
            nslow = sum(1 for m in s5.tog if not s5.fast[m])
            if fb >= l or len(s5.R) > capR - 4 or nslow >= 8:
                s5.fold()
    st = s5.stat
    assert st["joins"] > 0 and st["folds"] > 0 and st["fused_folds"] >= 1      # every operation was exercised


def test_fused_fold_and_compaction_are_exact(oracle):
    """fold_fused5's one-pass formula equals the variables swept into T1 one at a time, and dropping untoggled variables
    from the window leaves T2 = sweep(T1[R', R'], S) -- checked on a state with entered and swept-back variables."""
    o, oc = oracle
    from proto_v5 import V5, sweep
    _, G, c, yy, Po, gmask = _setup(o, oc, 300, 20, 4, 1e-3, 0.4, seed=5)
    Mp = len(c)
    rng = np.random.default_rng(1)
    s5 = V5(G, c, yy, gmask, 0, Mp)
    for m in (1, 4, 7, 9, 12):                           # commit a few variables to T1 first
        sweep(s5.T1, m, True); s5.inO[m] = True
    s5.rebuild_T2()
    for m in (0, 3, 4, 9, 15, 17, 18):                   # a window with entered (0, 3, 15, 17), swept-back (4, 9) and
        s5.join(m)                                       # later untoggled (18) variables
    for m in (0, 3, 4, 9, 15, 17, 18):
        s5.toggle(m)
    s5.toggle(18)                                        # 18 goes back to its T1 state
    ref1 = s5.T1.copy()
    for m, e in s5.tog.items():
        sweep(ref1, m, e > 0)
    # compaction: T2 on the remaining slots is the sweep of T1 restricted to them
    s5.compact()
    assert 18 not in s5.R and len(s5.R) == 6
    idx = s5.R + [Mp]
    T2ref = s5.T1[np.ix_(idx, idx)].copy()
    for m, e in s5.tog.items():
        sweep(T2ref, s5.R.index(m), e > 0)
    assert np.abs(T2ref - s5.T2).max() <= 1e-10 * np.abs(T2ref).max()
    # fused fold
    s5.fold_fused()
    assert np.abs(s5.T1 - ref1).max() <= 1e-9 * np.abs(ref1).max()
    assert s5.inO[[0, 1, 3, 7, 12, 15, 17]].all() and not s5.inO[[4, 9, 18]].any() and not s5.tog
