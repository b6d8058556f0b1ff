"""CPU oracle for PartitionedLS.jl -- numpy/scipy restatement of the reference algorithms.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product path: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it, and
only as the checker / the timed CPU baseline.  The product (``libpls_cuda.so``) never calls it.

Every function cites the reference lines it restates (paths relative to /root/reference).
The NNLS arithmetic of the reference lives in the un-vendored package NonNegLeastSquares.jl
(compat "0.4", Project.toml:20; Manifest not committed => patch version unpinned); its ``:nnls``
algorithm is Lawson & Hanson's 1974 active-set NNLS in data space.  Here that call is
``scipy.optimize.nnls`` (also Lawson-Hanson) and, in ``pls_oracle.c``, an own restatement of the
published algorithm.

Parity pinning: the only reference-held known answers reproducible without Julia are the toy
problem's ``opt ~ 0`` / ``predict(X) == y`` (test/runtests.jl:8-39, 41-69, 123-146).  Everything
else (eta > 0, K > 2, BnB/Alt on random data) is "parity unpinned by the reference's own tests":
it rests on this restatement plus uniqueness of each orthant's NNLS optimum.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import nnls as _scipy_nnls


# ----------------------------------------------------------------------------------------
# Problem rewriting  (src/PartitionedLS.jl)
# ----------------------------------------------------------------------------------------
def homogeneous_coords(X, P):
    """src/PartitionedLS.jl:76-81 -- Xo = [X 1], Po = [[P 0]; [0 .. 0 1]]."""
    X = np.asarray(X, dtype=np.float64)
    P = np.asarray(P, dtype=np.int64)
    N, M = X.shape
    K = P.shape[1]
    Xo = np.hstack([X, np.ones((N, 1))])
    Po = np.zeros((M + 1, K + 1), dtype=np.int64)
    Po[:M, :K] = P
    Po[M, K] = 1
    return Xo, Po


def regularize_problem(Xo, y, Po, eta):
    """src/PartitionedLS.jl:108-123 -- eta == 0: identity; else one extra row per column of Po,
    sqrt(eta) * 1[Po[:,k] == 1], and one extra 0 in y per row (intended semantics; the reference's
    own eta>0 path is untested upstream)."""
    y = np.asarray(y, dtype=np.float64)
    if eta == 0:
        return Xo, y
    rows = np.sqrt(eta) * (Po.T == 1).astype(np.float64)
    return np.vstack([Xo, rows]), np.concatenate([y, np.zeros(Po.shape[1])])


# ----------------------------------------------------------------------------------------
# Opt  (src/PartitionedLSOpt.jl)
# ----------------------------------------------------------------------------------------
def index_to_beta(b, K):
    """src/PartitionedLSOpt.jl:4-20 -- beta[k] = 2*bit_k(b) - 1, least significant bit first."""
    return np.array([2 * ((b >> k) & 1) - 1 for k in range(K)], dtype=np.int64)


def bmatrix(Xo, Po, beta):
    """src/PartitionedLSOpt.jl:22-31 -- scale column m of Xo by d_m = (Po @ beta)_m."""
    d = (Po * beta[None, :]).sum(axis=1)
    return Xo * d[None, :]


def nonneg_lsq(A, b):
    """The external call at Opt.jl:89 / Alt.jl:90 / BnB.jl:82 (alg=:nnls, Lawson-Hanson)."""
    x, _ = _scipy_nnls(A, b, maxiter=max(30 * A.shape[1], 300))
    return x


def opt_orthant(Xa, ya, Po, b):
    """One body of the loop at src/PartitionedLSOpt.jl:85-94: (optval, alpha_raw[M'], beta[K'])."""
    Kp = Po.shape[1]
    beta = index_to_beta(b, Kp)
    Xb = bmatrix(Xa, Po, beta)
    alpha = nonneg_lsq(Xb, ya)
    optval = float(np.linalg.norm(Xa @ ((Po * alpha[:, None]) @ beta) - ya))
    return optval, alpha, beta


def cleanup_result_opt(alpha_raw, beta, P):
    """src/PartitionedLSOpt.jl:34-44 on the tuple of :92 -- returns (alpha[M], beta[K], t).
    alpha_raw has M+1 entries, beta K+1 signs."""
    P = np.asarray(P, dtype=np.float64)
    a = np.asarray(alpha_raw, dtype=np.float64)[:-1]
    bsign = np.asarray(beta, dtype=np.float64)[:-1]
    t = float(beta[-1] * alpha_raw[-1])
    A = (P * a[:, None]).sum(axis=0)
    bb = bsign * A
    A = A.copy()
    A[A == 0.0] = 1.0
    aa = ((P * a[:, None]) / A[None, :]).sum(axis=1)
    return aa, bb, t


def fit_opt(X, y, P, eta=0.0, return_all=False):
    """src/PartitionedLSOpt.jl:73-104.  Returns dict(alpha, beta, t, opt, b_best, alpha_raw, objs)."""
    Xo, Po = homogeneous_coords(X, P)
    Xa, ya = regularize_problem(Xo, y, Po, float(eta))
    Kp = Po.shape[1]
    results = [opt_orthant(Xa, ya, Po, b) for b in range(2 ** Kp)]
    objs = np.array([r[0] for r in results])
    # Julia argmin: first minimal index; NaN wins (Opt.jl:96)
    nan = np.flatnonzero(np.isnan(objs))
    best = int(nan[0]) if len(nan) else int(np.argmin(objs))
    opt, araw, beta = results[best]
    a, bb, t = cleanup_result_opt(araw, beta, P)
    out = dict(alpha=a, beta=bb, t=t, opt=opt, b_best=best, alpha_raw=araw, objs=objs)
    if return_all:
        out["solutions"] = [(r[0],) + cleanup_result_opt(r[1], r[2], P) for r in results]
    return out


def predict(alpha, beta, t, P, X):
    """src/PartitionedLS.jl:132-134 -- X * (P .* alpha) * beta .+ t."""
    P = np.asarray(P, dtype=np.float64)
    return np.asarray(X, dtype=np.float64) @ ((P * np.asarray(alpha)[:, None]) @ np.asarray(beta)) + t


# ----------------------------------------------------------------------------------------
# BnB  (src/PartitionedLSBnB.jl)
# ----------------------------------------------------------------------------------------
def sum_max_0_ai_aj(Po, alpha):
    """src/PartitionedLSBnB.jl:42-57 -- nu_k = sum_{i<j in group k} max(0, -a_i a_j)."""
    K = Po.shape[1]
    out = np.zeros(K)
    for k in range(K):
        idx = np.flatnonzero(Po[:, k] != 0)
        a = alpha[idx]
        prod = -np.outer(a, a)
        iu = np.triu_indices(len(idx), 1)
        out[k] = np.maximum(0.0, prod[iu]).sum()
    return out


def lower_bound(Xa, ya, sigma_constr):
    """src/PartitionedLSBnB.jl:69-92.  sigma_constr: list of signed 1-based variable indices
    (+i: alpha_i >= 0, -i: alpha_i <= 0).  Returns (lb, alpha_signed)."""
    pos = [s - 1 for s in sigma_constr if s > 0]
    neg = [-s - 1 for s in sigma_constr if s < 0]
    M = Xa.shape[1]
    Xp = Xa.copy()
    Xm = -Xa.copy()
    Xp[:, neg] = 0
    Xm[:, pos] = 0
    XX = np.hstack([Xp, Xm])
    aa = nonneg_lsq(XX, ya)
    lb = float(np.linalg.norm(XX @ aa - ya))
    ap = aa[:M].copy()
    an = aa[M:].copy()
    ap[neg] = 0
    an[pos] = 0
    return lb, ap - an


def _fit_bnb_rec(Xa, ya, Po, mu, sigma):
    """src/PartitionedLSBnB.jl:94-132 (depth-first, positive child first)."""
    lb, alpha = lower_bound(Xa, ya, sigma)
    if lb >= mu:
        return np.inf, None, 1
    nu = sum_max_0_ai_aj(Po, alpha)
    if np.all(nu == 0):
        return float(np.linalg.norm(Xa @ alpha - ya)), alpha, 1
    k = int(np.argmax(nu))
    pk = [int(i) + 1 for i in np.flatnonzero(Po[:, k] == 1)]
    mup, ap, nop = _fit_bnb_rec(Xa, ya, Po, mu, sigma + pk)
    mum, am, nom = _fit_bnb_rec(Xa, ya, Po, min(mu, mup), sigma + [-i for i in pk])
    cands = [(mu, alpha), (mup, ap), (mum, am)]
    i = int(np.argmin([c[0] for c in cands]))
    return cands[i][0], cands[i][1], nop + nom + 1


def fit_bnb(X, y, P, eta=0.0):
    """src/PartitionedLSBnB.jl:30-40.  Returns dict(alpha, beta, t, opt, nopen, alpha_signed)."""
    Xo, Po = homogeneous_coords(X, P)
    Xa, ya = regularize_problem(Xo, y, Po, float(eta))
    opt, a_s, nopen = _fit_bnb_rec(Xa, ya, Po, np.inf, [])
    Pf = Po.astype(np.float64)
    beta = (Pf * a_s[:, None]).sum(axis=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        alpha = (Pf * a_s[:, None] / beta[None, :]).sum(axis=1)   # no zero guard upstream (q9)
    return dict(alpha=alpha[:-1], beta=beta[:-1], t=float(beta[-1]), opt=opt, nopen=nopen,
                alpha_signed=a_s)


# ----------------------------------------------------------------------------------------
# Alt  (src/PartitionedLSAlt.jl)
# ----------------------------------------------------------------------------------------
def checkalpha(a, Po):
    """src/PartitionedLSAlt.jl:5-20 -- an all-zero group becomes uniform 1/|group|."""
    a = a.copy()
    suma = (Po * a[:, None]).sum(axis=0)
    sumP = Po.sum(axis=0)
    for k in range(Po.shape[1]):
        if suma[k] == 0.0:
            a[Po[:, k] == 1] = 1.0 / sumP[k]
    return a


def fit_alt(X, y, P, beta0, eta=0.0, eps=1e-6, T=100):
    """src/PartitionedLSAlt.jl:50-124 with the initial beta supplied by the caller (the RNG
    stays on the host side; alpha_0 is dead upstream, q6).  Uses ya in the alpha-step (q5)."""
    Xo, Po = homogeneous_coords(X, P)
    Xa, ya = regularize_problem(Xo, y, Po, float(eta))
    Pf = Po.astype(np.float64)
    beta = np.asarray(beta0, dtype=np.float64).copy()
    alpha = np.zeros(Po.shape[0])
    old, opt, i = 1e20, 1e10, 1
    while i <= T and abs(old - opt) > eps * old:
        d = (Pf * beta[None, :]).sum(axis=1)
        alpha = nonneg_lsq(Xa * d[None, :], ya)
        alpha = checkalpha(alpha, Po)
        suma = (Pf * alpha[:, None]).sum(axis=0)
        Pa = (Pf * suma[None, :]).sum(axis=1)
        alpha = alpha / Pa
        beta = beta * suma
        Xalpha = Xa @ (Pf * alpha[:, None])
        beta = np.linalg.lstsq(Xalpha, ya, rcond=None)[0]
        old = opt
        opt = float(np.linalg.norm(Xa @ ((Pf * alpha[:, None]) @ beta) - ya))
        i += 1
    return dict(alpha=alpha[:-1], beta=beta[:-1], t=float(beta[-1] * alpha[-1]), opt=opt,
                iters=i - 1, alpha_full=alpha, beta_full=beta)


# ----------------------------------------------------------------------------------------
# Synthetic workloads -- shared by tests and bench
# ----------------------------------------------------------------------------------------
TOY_X = np.array([[1.0, 2.0, 3.0], [3.0, 3.0, 4.0], [8.0, 1.0, 3.0], [5.0, 3.0, 1.0]])
TOY_Y = np.array([1.0, 1.0, 2.0, 3.0])
TOY_P = np.array([[1, 0], [1, 0], [0, 1]], dtype=np.int64)


def make_synthetic(N, M, K, seed, mixed_sign=False, rho=0.0):
    """X ~ N(0,1) (optionally AR(1)-correlated columns), contiguous near-equal groups,
    w_m = |z_m| * s_group (or mixed sign), y = X w + 0.5 + noise."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, M))
    if rho:
        for m in range(1, M):
            X[:, m] = rho * X[:, m - 1] + np.sqrt(1 - rho * rho) * X[:, m]
    g = (np.arange(M) * K) // M
    s = rng.choice([-1.0, 1.0], size=K)
    z = rng.standard_normal(M)
    w = z if mixed_sign else np.abs(z) * s[g]
    y = X @ w + 0.5 + rng.standard_normal(N)
    P = np.zeros((M, K), dtype=np.int64)
    P[np.arange(M), g] = 1
    return np.asfortranarray(X), y, P
