"""TEST INFRASTRUCTURE ONLY -- ctypes wrappers around oracle/weave_kernels.c with the reference's
signatures (GPclust/utilities.py:27 `softmax_weave(X)`, :108 `multiple_mahalanobis(X1, X2, L)`)."""
import ctypes
import os

import numpy as np

from .build import LIB, build_oracle

_lib = None


def _load():
    global _lib
    if _lib is None:
        path = LIB if os.path.exists(LIB) else build_oracle()
        _lib = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.gpo_softmax.argtypes = [dp, dp, dp, dp, ctypes.c_long, ctypes.c_long]
        _lib.gpo_softmax.restype = ctypes.c_double
        _lib.gpo_multiple_mahalanobis.argtypes = [dp, dp, dp, ctypes.c_long, ctypes.c_long, ctypes.c_long, dp]
        _lib.gpo_multiple_mahalanobis.restype = None
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def softmax_weave(X):
    lib = _load()
    X = np.ascontiguousarray(X, dtype=np.float64)
    N, D = X.shape
    phi = np.empty_like(X)
    log_phi = np.empty_like(X)
    scratch = np.empty(N)
    H = lib.gpo_softmax(_p(X), _p(phi), _p(log_phi), _p(scratch), N, D)
    return phi, log_phi, H


def multiple_mahalanobis(X1, X2, L):
    lib = _load()
    X1 = np.ascontiguousarray(X1, dtype=np.float64)
    X2 = np.ascontiguousarray(X2, dtype=np.float64)
    L = np.ascontiguousarray(L, dtype=np.float64)
    N1, D = X1.shape
    N2, D2 = X2.shape
    assert D == D2 and L.shape == (D, D)
    out = np.empty((N1, N2))
    lib.gpo_multiple_mahalanobis(_p(X1), _p(X2), _p(L), N1, N2, D, _p(out))
    return out
