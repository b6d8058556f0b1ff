"""ctypes loader for oracle/libpls_oracle.so (the C restatement).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libpls_oracle.so")
    src = os.path.join(_HERE, "pls_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpls_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libpls_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        L.pls_oracle_opt_fit.restype = C.c_int
        L.pls_oracle_opt_fit.argtypes = [dp, C.c_int64, C.c_int64, dp, ip, C.c_int64, C.c_double,
                                         ip, C.c_int64, C.c_int, dp, dp, ip, dp, dp]
        L.pls_oracle_nnls.restype = C.c_int
        L.pls_oracle_nnls.argtypes = [dp, C.c_int, C.c_int, dp, dp, dp, dp, dp, C.POINTER(C.c_int)]
        L.pls_oracle_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def nnls(A, b):
    A = np.array(A, dtype=np.float64, order="F")
    b = np.array(b, dtype=np.float64)
    m, n = A.shape
    x = np.zeros(n); w = np.zeros(n); z = np.zeros(m); idx = np.zeros(n, dtype=np.int32)
    rn = C.c_double()
    rc = lib().pls_oracle_nnls(_dp(A), m, n, _dp(b), _dp(x), C.byref(rn), _dp(w), _dp(z),
                               idx.ctypes.data_as(C.POINTER(C.c_int)))
    return x, rn.value, rc


def opt_fit(X, y, P, eta=0.0, b_list=None, nthreads=0, want_alpha=True):
    """Returns dict(objs, alphas (nb x M') or None, b_best, obj_best, alpha_best)."""
    X = np.asfortranarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    P = np.asfortranarray(P, dtype=np.int64)
    N, M = X.shape
    K = P.shape[1]
    if b_list is None:
        nb = 1 << (K + 1); bl = None
    else:
        bl = np.ascontiguousarray(b_list, dtype=np.int64); nb = len(bl)
    objs = np.zeros(nb)
    alphas = np.zeros((nb, M + 1)) if want_alpha else None
    bb = C.c_int64(); ob = C.c_double(); ab = np.zeros(M + 1)
    rc = lib().pls_oracle_opt_fit(
        _dp(X), N, M, _dp(y), P.ctypes.data_as(C.POINTER(C.c_int64)), K, float(eta),
        bl.ctypes.data_as(C.POINTER(C.c_int64)) if bl is not None else None, nb, int(nthreads),
        _dp(objs), _dp(alphas) if want_alpha else None, C.byref(bb), C.byref(ob), _dp(ab))
    if rc != 0:
        raise RuntimeError(f"pls_oracle_opt_fit failed rc={rc}")
    return dict(objs=objs, alphas=alphas, b_best=bb.value, obj_best=ob.value, alpha_best=ab)


def num_threads():
    return lib().pls_oracle_num_threads()
