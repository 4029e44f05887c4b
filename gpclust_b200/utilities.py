"""The reference's L1 backend seam (GPclust/utilities.py == GPclust/np_utilities.py signatures), served by
the device library: numpy in, fresh numpy arrays out, `LinAlgError` on a failed factorisation.

    softmax(X) -> (phi, log_phi, H)                      np_utilities.py:25-40 / utilities.py:27-105
    multiple_mahalanobis(X1, X2, L) -> N1 x N2           utilities.py:108-172 / np_utilities.py:43-59
    multiple_pdinv(A) -> (inv D x D x K, halflogdet K)   np_utilities.py:6-22
    lngammad(v, D), ln_dirichlet_C(a)                    np_utilities.py:62-69 (K-sized host scalars)

These are the read-out / drop-in forms; the optimiser itself never materialises phi or the N x K
Mahalanobis matrix (see collapsed_mixture.py).
"""
import ctypes

import numpy as np
from numpy.linalg import LinAlgError
from scipy.special import gammaln

from . import _cabi
from ._cabi import lib, check, ptr


def _stream(torch):
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def softmax(X):
    torch = _cabi.require_cuda()
    X = _cabi.f64(X)
    assert X.ndim == 2
    N, K = X.shape
    xd = torch.from_numpy(X).cuda()
    phi = torch.empty_like(xd)
    logphi = torch.empty_like(xd)
    grid = max(1, min(4 * lib.gpc_default_grid(torch.cuda.current_device()), (N + 7) // 8))
    hpart = torch.zeros(grid, dtype=torch.float64, device=xd.device)
    H = torch.zeros(1, dtype=torch.float64, device=xd.device)
    check(lib.gpc_softmax_rows(ptr(xd), N, K, K, ptr(phi), ptr(logphi), K, ptr(hpart), grid, _stream(torch)), "gpc_softmax_rows")
    check(lib.gpc_reduce_partials(ptr(hpart), grid, 1, ptr(H), _stream(torch)), "gpc_reduce_partials")
    return phi.cpu().numpy(), logphi.cpu().numpy(), float(H.cpu()[0])


softmax_weave = softmax
softmax_numpy = softmax


def multiple_mahalanobis(X1, X2, L):
    """X1: N1 x D, X2: N2 x D, L: lower Cholesky factor D x D -> N1 x N2 squared distances."""
    torch = _cabi.require_cuda()
    X1, X2, L = _cabi.f64(X1), _cabi.f64(X2), _cabi.f64(L)
    N1, D = X1.shape
    N2, D2 = X2.shape
    assert D == D2
    assert L.shape == (D, D)
    a, b, l = torch.from_numpy(X1).cuda(), torch.from_numpy(X2).cuda(), torch.from_numpy(L).cuda()
    out = torch.empty((N1, N2), dtype=torch.float64, device=a.device)
    check(lib.gpc_multiple_mahalanobis(ptr(a), ptr(b), ptr(l), N1, N2, D, ptr(out), _stream(torch)), "gpc_multiple_mahalanobis")
    return out.cpu().numpy()


multiple_mahalanobis_numpy_loops = multiple_mahalanobis


def multiple_pdinv(A):
    """A: D x D x K stack of positive-definite matrices -> (inverses D x D x K, half log-determinants K)."""
    torch = _cabi.require_cuda()
    A = _cabi.f64(A)
    D, D2, K = A.shape
    assert D == D2
    a = torch.from_numpy(A).cuda()
    inv = torch.empty_like(a)
    hld = torch.empty(K, dtype=torch.float64, device=a.device)
    info = torch.zeros(1, dtype=torch.int32, device=a.device)
    check(lib.gpc_multiple_pdinv(ptr(a), D, K, ptr(inv), ptr(hld), ptr(info), _stream(torch)), "gpc_multiple_pdinv")
    if int(info.cpu()[0]) != 0:
        raise LinAlgError("not positive definite, even with jitter.")
    return inv.cpu().numpy(), hld.cpu().numpy()


def lngammad(v, D):
    """sum of log gamma functions, as appears in a Wishart distribution (np_utilities.py:62-64)"""
    return np.sum([gammaln((v + 1.0 - d) / 2.0) for d in range(1, D + 1)], 0)


def ln_dirichlet_C(a):
    """the log-normalizer of a Dirichlet distribution (np_utilities.py:67-69)"""
    return gammaln(a.sum()) - np.sum(gammaln(a))
