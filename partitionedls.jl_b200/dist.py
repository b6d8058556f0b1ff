"""One-process-per-GPU Opt fit: the host-side plumbing around the stage-wise C ABI entry points.

Sharding:
  K1  rows of X are sharded; the raw Gram sums S = Z'Z are all-reduced once ((M+2)^2 doubles);
  K2  the orthant index space [0, 2^(K+1)) is split into one contiguous range per rank, no
      communication;
  K3  every rank's (objective, b, alpha) winner is all-gathered, the lexicographic minimum
      (objective, b) wins on every rank (first-minimum semantics of Opt.jl:96);
  K4  the winner's squared residual is summed over the row shards (one double all-reduced).

`backend` is anything with the stage-wise methods of `_abi.Context` (load, gram_build, gram_raw,
gram_finalize, opt_solve_range, residual_partial, objective_finish); `comm` wraps the collectives so
the same code runs over NCCL (GPU tensors) in bench.py and over gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_rows(N: int, rank: int, world: int):
    """Contiguous near-equal row ranges."""
    return (N * rank) // world, (N * (rank + 1)) // world


def shard_orthants(total: int, rank: int, world: int):
    """Contiguous equal orthant ranges; world must divide the (power of two) total."""
    if total % world:
        raise ValueError(f"world size {world} does not divide {total} orthants")
    n = total // world
    return n * rank, n


def pick_winner(records, Mp, tau=0.0):
    """records: array (world, Mp + 2) rows [alpha_raw (Mp), objective, b].  The argmin rule of the orthant enumeration
    (Opt.jl:96: first minimum, NaN objectives first -- Julia argmin semantics) with the library's tie tolerance: squared
    objectives closer than tau = 1e-13 * y'y count as equal and the lower index wins (include/pls.h: pls_gram_scalars)."""
    rec = np.asarray(records, dtype=np.float64)

    def better(a, c):
        oa, ba, ob, bb = a[Mp], a[Mp + 1], c[Mp], c[Mp + 1]
        na, nb = np.isnan(oa), np.isnan(ob)
        if na != nb:
            return bool(na)
        if na:
            return ba < bb
        if abs(oa - ob) * (oa + ob) <= tau:
            return ba < bb
        return oa < ob

    i = 0
    for q in range(1, len(rec)):
        if better(rec[q], rec[i]):
            i = q
    return rec[i, :Mp].copy(), int(rec[i, Mp + 1]), float(rec[i, Mp]), i


class TorchComm:
    """Collectives over torch.distributed (NCCL for CUDA tensors, gloo for CPU tensors).  The two small exchanges of a
    fit (winner records, winner residual) go through ONE preallocated pinned host buffer and ONE device buffer each, so
    a fit costs one host->device copy, one collective and one device->host copy per exchange and no allocation."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.device = torch, dist, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self._bufs = {}

    def _buf(self, key, n):
        """(pinned host tensor, device tensor or None) of n doubles, cached."""
        b = self._bufs.get((key, n))
        if b is None:
            t = self.torch
            host = t.empty(n, dtype=t.float64, pin_memory=self.device is not None)
            dev = t.empty(n, dtype=t.float64, device=self.device) if self.device is not None else None
            b = self._bufs[(key, n)] = (host, dev)
        return b

    def allreduce_sum_inplace_dev(self, dev_ptr: int, count: int):
        """Sum a device buffer of `count` doubles across ranks, in place (zero-copy view)."""
        class _Buf:
            pass
        b = _Buf()
        b.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (dev_ptr, False), "version": 3}
        t = self.torch.as_tensor(b, device=self.device)
        self.dist.all_reduce(t)
        self.torch.cuda.synchronize(self.device)

    def allreduce_sum(self, arr: np.ndarray) -> np.ndarray:
        arr = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
        host, dev = self._buf("ar", arr.size)
        host.numpy()[:] = arr
        if dev is None:
            self.dist.all_reduce(host)
            return host.numpy().copy()
        dev.copy_(host, non_blocking=True)
        self.dist.all_reduce(dev)
        host.copy_(dev)                      # blocking device->host copy: the one synchronisation of this exchange
        return host.numpy().copy()

    def allgather(self, arr: np.ndarray) -> np.ndarray:
        arr = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
        host, dev = self._buf("ag_in", arr.size)
        hout, dout = self._buf("ag_out", arr.size * self.world)
        host.numpy()[:] = arr
        if dev is None:
            self.dist.all_gather_into_tensor(hout, host)
            return hout.numpy().reshape(self.world, -1).copy()
        dev.copy_(host, non_blocking=True)
        self.dist.all_gather_into_tensor(dout, dev)
        hout.copy_(dout)
        return hout.numpy().reshape(self.world, -1).copy()


def opt_fit_sharded(backend, comm, Mp: int, Kp: int, *, reload=None, pairs=None):
    """One Opt fit over `comm.world` ranks on data already loaded in `backend` (or loaded by the
    `reload()` callback first -- the end-to-end variant).  Returns (b*, objective, alpha_raw).
    pairs (default: whenever the backend offers it and the problem fits): shard the 2^K sign patterns of
    the user groups and leave the intercept sign free -- each solve resolves two reference orthants.

    Exchanges per fit: the all-reduce of the raw Gram sums (in place, on the device), ONE all-gather of the ranks'
    winner records and ONE all-reduce of the winner's squared residual over the row shards.  A rank whose local solve
    fails does not leave the others waiting in a collective: it sends a record flagged as failed, and every rank raises
    after the gather."""
    if reload is not None:
        reload()
    backend.gram_build()
    ptr, count = backend.gram_raw()
    if isinstance(ptr, np.ndarray):          # CPU test double: host array, reduced out of place
        ptr[...] = comm.allreduce_sum(ptr)
    else:
        comm.allreduce_sum_inplace_dev(ptr, count)
    backend.gram_finalize()
    if pairs is None:
        pairs = hasattr(backend, "opt_solve_pairs") and Mp <= 1024 and Kp >= 2 and (1 << (Kp - 1)) >= comm.world
    rec = np.zeros(Mp + 3)
    err = None
    try:
        if pairs:
            b0, bn = shard_orthants(1 << (Kp - 1), comm.rank, comm.world)
            loc = backend.opt_solve_pairs(b0, bn)
        else:
            b0, bn = shard_orthants(1 << Kp, comm.rank, comm.world)
            loc = backend.opt_solve_range(b0, bn)
        rec[:Mp], rec[Mp], rec[Mp + 1] = loc["alpha_raw"], loc["obj_gram"], float(loc["b_best"])
    except Exception as e:                   # sentinel record: objective +inf, b = -1, error flag
        err = e
        rec[Mp], rec[Mp + 1], rec[Mp + 2] = np.inf, -1.0, 1.0
    allrec = comm.allgather(rec)
    if allrec[:, Mp + 2].any():
        bad = [int(r) for r in np.flatnonzero(allrec[:, Mp + 2])]
        raise RuntimeError(f"opt_fit_sharded: the local solve failed on rank(s) {bad}" + (f": {err}" if err is not None else ""))
    tau = 1e-13 * backend.gram_scalars()[0] if hasattr(backend, "gram_scalars") else 0.0
    alpha, b, _, _ = pick_winner(allrec[:, :Mp + 2], Mp, tau)
    ssq = comm.allreduce_sum(np.array([backend.residual_partial(alpha, b)]))[0]
    return b, backend.objective_finish(alpha, b, float(ssq)), alpha


def shard_restarts(R: int, rank: int, world: int):
    """Contiguous near-equal restart ranges (a rank may get none when R < world)."""
    return (R * rank) // world, (R * (rank + 1)) // world


def alt_fit_sharded(backend, comm, Po, beta0, eps=1e-6, T=100):
    """fit(Alt) with batched restarts over `comm.world` ranks: rows sharded for the Gram build and the
    final data-space loss, restarts sharded for the iteration (no exchange), the ranks' best
    (loss, restart, alpha, beta) records all-gathered and the lexicographic minimum taken.
    Po: (M+1) x (K+1) membership matrix incl. the intercept group; beta0: (K+1) x R.
    Returns dict(alpha, beta, opt, best_restart, iters)."""
    Po = np.asarray(Po, dtype=np.float64)
    Mp, Kp = Po.shape
    beta0 = np.asarray(beta0, dtype=np.float64).reshape(Kp, -1)
    backend.gram_build()
    ptr, count = backend.gram_raw()
    if isinstance(ptr, np.ndarray):
        ptr[...] = comm.allreduce_sum(ptr)
    else:
        comm.allreduce_sum_inplace_dev(ptr, count)
    backend.gram_finalize()
    r0, r1 = shard_restarts(beta0.shape[1], comm.rank, comm.world)
    rec = np.zeros(Mp + Kp + 3)
    rec[Mp + Kp], rec[Mp + Kp + 1] = np.inf, -1.0
    if r1 > r0:
        try:
            loc = backend.alt_fit_shard(beta0[:, r0:r1], eps=eps, T=T)
            rec[:Mp], rec[Mp:Mp + Kp] = loc["alpha"], loc["beta"]
            rec[Mp + Kp], rec[Mp + Kp + 1], rec[Mp + Kp + 2] = loc["opt"], r0 + loc["best_restart"], loc["iters"]
        except Exception:                    # every restart of this shard failed: "no candidate" (as multi.cu does); the
            pass                             # other ranks are not left waiting in the collective
    allrec = comm.allgather(rec)
    cand = [q for q in allrec if q[Mp + Kp + 1] >= 0]
    if not cand:
        raise RuntimeError("alt: every restart failed")
    best = min(cand, key=lambda q: (q[Mp + Kp], q[Mp + Kp + 1]))
    alpha, beta = best[:Mp].copy(), best[Mp:Mp + Kp].copy()
    w = alpha * (Po @ beta)
    ssq = comm.allreduce_sum(np.array([backend.residual_partial_w(w)]))[0]
    return dict(alpha=alpha, beta=beta, opt=backend.objective_finish_w(w, float(ssq)),
                best_restart=int(best[Mp + Kp + 1]), iters=int(best[Mp + Kp + 2]))
