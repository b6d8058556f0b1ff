"""world_size-2 gloo test of the multi-process Opt fit plumbing (partitionedls.jl_b200/dist.py) on the
CPU.  The per-rank compute is a test double with the stage-wise interface of the GPU context: numpy
Gram sums for the rank's row shard and oracle solves for its orthant range (tests may use the
oracle; the product path never does)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """Stage-wise interface of _abi.Context, computed on the CPU for one rank's shard."""

    def __init__(self, o, oc, X, y, P, eta, rows):
        self.o, self.oc, self.eta = o, oc, eta
        self.Xfull, self.yfull, self.P = X, y, P
        r0, r1 = rows
        Z = np.hstack([X[r0:r1], np.ones((r1 - r0, 1)), y[r0:r1, None]])
        self.Z = Z
        self.S = None

    def gram_build(self):
        self.S = np.ascontiguousarray(self.Z.T @ self.Z).reshape(-1)

    def gram_raw(self):
        return self.S, self.S.size

    def gram_finalize(self):
        zc = self.Z.shape[1]
        S = self.S.reshape(zc, zc)
        _, Po = self.o.homogeneous_coords(self.Xfull[:1], self.P)
        self.G = S[:-1, :-1] + self.eta * (Po @ Po.T)
        self.c = S[:-1, -1].copy()
        self.yy = S[-1, -1]
        self.Po = Po

    def opt_solve_range(self, b0, bn):
        # oracle solve of the rank's orthant range on the GLOBAL Gram (Cholesky surrogate data)
        L = np.linalg.cholesky(self.G)
        A = L.T
        bb = np.linalg.solve(L, self.c)
        best = None
        for b in range(b0, b0 + bn):
            beta = self.o.index_to_beta(b, self.Po.shape[1])
            d = self.Po @ beta
            alpha = self.o.nonneg_lsq(A * d[None, :], bb)
            w = d * alpha
            obj = float(np.sqrt(max(self.yy - 2 * self.c @ w + w @ self.G @ w, 0.0)))
            if best is None or (obj, b) < (best["obj_gram"], best["b_best"]):
                best = dict(alpha_raw=alpha, b_best=b, obj_gram=obj)
        return best

    def opt_solve_pairs(self, p0, pn):
        # the two reference orthants p and p + 2^K of every sign pattern p of the user groups (what the
        # GPU solves as ONE problem with a free intercept): the better of the two, first one on a tie
        K = self.Po.shape[1] - 1
        best = None
        for p in range(p0, p0 + pn):
            for b in (p, p + (1 << K)):
                r = self.opt_solve_range(b, 1)
                if best is None or (r["obj_gram"], r["b_best"]) < (best["obj_gram"], best["b_best"]):
                    best = r
        return best

    def residual_partial(self, alpha, b):
        d = self.Po @ self.o.index_to_beta(b, self.Po.shape[1])
        res = self.Z[:, :-1] @ (d * alpha) - self.Z[:, -1]
        return float(res @ res)

    def objective_finish(self, alpha, b, ssq):
        d = self.Po @ self.o.index_to_beta(b, self.Po.shape[1])
        return float(np.sqrt(ssq + self.eta * np.sum((self.Po.T @ (d * alpha)) ** 2)))


    # ---- Alt: restart shard on the full problem (test double; the plumbing is what is under test)
    def alt_fit_shard(self, beta0_cols, eps=1e-6, T=100):
        runs = [self.o.fit_alt(self.Xfull, self.yfull, self.P, beta0_cols[:, r], eta=self.eta, eps=eps, T=T)
                for r in range(beta0_cols.shape[1])]
        i = int(np.argmin([q["opt"] for q in runs]))
        return dict(alpha=runs[i]["alpha_full"], beta=runs[i]["beta_full"], opt=runs[i]["opt"], best_restart=i,
                    iters=runs[i]["iters"])

    def residual_partial_w(self, w):
        res = self.Z[:, :-1] @ w - self.Z[:, -1]
        return float(res @ res)

    def objective_finish_w(self, w, ssq):
        return float(np.sqrt(ssq + self.eta * np.sum((self.Po.T @ w) ** 2)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import __graft_entry__ as g
    from importlib import import_module
    g.load_package()
    distmod = import_module(g.PKG_NAME + ".dist")
    o, oc = g.load_oracle()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, y, P = o.make_synthetic(400, 10, 3, seed=5, mixed_sign=True)
        eta = 1e-2
        be = OracleBackend(o, oc, X, y, P, eta, distmod.shard_rows(400, rank, world))
        comm = distmod.TorchComm(device=None)
        b, obj, alpha = distmod.opt_fit_sharded(be, comm, Mp=11, Kp=4)                 # paired orthants (default)
        b_f, obj_f, alpha_f = distmod.opt_fit_sharded(be, comm, Mp=11, Kp=4, pairs=False)   # literal enumeration
        assert b_f == b and abs(obj_f - obj) <= 1e-12 * obj and np.allclose(alpha_f, alpha, rtol=1e-10, atol=1e-14)
        beta0 = (np.random.default_rng(3).random((4, 5)) - 0.5) * 10.0      # 5 restarts over 2 ranks
        ra = distmod.alt_fit_sharded(be, comm, be.Po, beta0, eps=1e-6, T=50)
        q.put((rank, b, obj, alpha, ra))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_fit_matches_single_process(oracle):
    import torch.multiprocessing as mp
    o, oc = oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X, y, P = o.make_synthetic(400, 10, 3, seed=5, mixed_sign=True)
    ref = oc.opt_fit(X, y, P, 1e-2)
    beta0 = (np.random.default_rng(3).random((4, 5)) - 0.5) * 10.0
    runs = [o.fit_alt(X, y, P, beta0[:, r], eta=1e-2, eps=1e-6, T=50) for r in range(5)]
    ibest = int(np.argmin([q_["opt"] for q_ in runs]))
    for rank, b, obj, alpha, ra in res:
        assert b == ref["b_best"]
        assert abs(obj - ref["obj_best"]) <= 1e-9 * ref["obj_best"]
        assert np.all(np.abs(alpha - ref["alpha_best"]) <= 1e-8 * np.abs(ref["alpha_best"]).max())
        assert ra["best_restart"] == ibest and abs(ra["opt"] - runs[ibest]["opt"]) <= 1e-9 * runs[ibest]["opt"]
        assert np.allclose(ra["alpha"], runs[ibest]["alpha_full"]) and np.allclose(ra["beta"], runs[ibest]["beta_full"])


def _worker_fail(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import __graft_entry__ as g
    from importlib import import_module
    g.load_package()
    distmod = import_module(g.PKG_NAME + ".dist")
    o, oc = g.load_oracle()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, y, P = o.make_synthetic(200, 6, 2, seed=6, mixed_sign=True)
        be = OracleBackend(o, oc, X, y, P, 1e-2, distmod.shard_rows(200, rank, world))
        if rank == 1:
            def boom(*a, **k):
                raise ValueError("solver failed on this rank")
            be.opt_solve_pairs = boom
        comm = distmod.TorchComm(device=None)
        try:
            distmod.opt_fit_sharded(be, comm, Mp=7, Kp=3)
            q.put((rank, "no error"))
        except RuntimeError as e:
            q.put((rank, str(e)))
    finally:
        dist.destroy_process_group()


def test_rank_local_failure_raises_on_every_rank_instead_of_hanging():
    """A local solve that fails on one rank must not leave the others blocked in the all-gather: the failing rank sends
    a flagged record and every rank raises after the collective."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_fail, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert "rank(s) [1]" in res[0] and "rank(s) [1]" in res[1] and "solver failed on this rank" in res[1]


def test_sharding_helpers(pkg):
    from importlib import import_module
    import __graft_entry__ as g
    d = import_module(g.PKG_NAME + ".dist")
    assert [d.shard_rows(10, r, 3) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]
    assert [d.shard_orthants(16, r, 4) for r in range(4)] == [(0, 4), (4, 4), (8, 4), (12, 4)]
    assert [d.shard_restarts(5, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 3), (3, 5)]
    assert [d.shard_restarts(2, r, 4) for r in range(4)] == [(0, 0), (0, 1), (1, 1), (1, 2)]
    with pytest.raises(ValueError):
        d.shard_orthants(16, 0, 3)
    rec = np.array([[1.0, 2.0, 5.0, 7.0], [3.0, 4.0, 5.0, 3.0], [0.0, 0.0, 6.0, 0.0]])
    a, b, obj, i = d.pick_winner(rec, 2)
    assert (b, obj, i) == (3, 5.0, 1) and list(a) == [3.0, 4.0]      # tie on objective -> lower b
    rec[2, 2] = np.nan
    assert d.pick_winner(rec, 2)[1] == 0                              # NaN sorts first
