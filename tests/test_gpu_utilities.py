"""GPU parity of the L1 seam (gpclust_b200.utilities == GPclust/np_utilities.py signatures) and of the
covariance fills against the CPU oracle.  Tolerance: 1e-12 max-norm relative (FP64, same formulas)."""
import numpy as np
import pytest

from helpers import relerr
from oracle import gpy_shim as G
from oracle import vb_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(1, 1), (7, 3), (64, 8), (1000, 20), (513, 64), (300, 130)])
def test_softmax(N, K):
    from gpclust_b200 import utilities as U
    rng = np.random.RandomState(N + K)
    X = rng.randn(N, K) * 5.0
    phi, logphi, H = U.softmax(X)
    phi0, logphi0, H0 = O.softmax_rows(X)
    assert relerr(phi, phi0) < 1e-13
    assert relerr(logphi, logphi0) < 1e-13
    assert abs(H - H0) <= 1e-12 * max(1.0, abs(H0))
    np.testing.assert_allclose(phi.sum(1), 1.0, atol=1e-14)


def test_softmax_extreme_logits():
    from gpclust_b200 import utilities as U
    X = np.array([[1e4, -1e4, 0.0], [-745.0, -746.0, -800.0], [0.0, 0.0, 0.0]])
    phi, logphi, H = U.softmax(X)
    phi0, logphi0, H0 = O.softmax_rows(X)
    assert np.all(np.isfinite(phi)) and relerr(phi, phi0) < 1e-14
    assert abs(H - H0) < 1e-12


@pytest.mark.parametrize("N1,N2,D", [(5, 3, 2), (100, 7, 12), (257, 64, 64)])
def test_multiple_mahalanobis(N1, N2, D):
    from gpclust_b200 import utilities as U
    rng = np.random.RandomState(D)
    X1, X2 = rng.randn(N1, D), rng.randn(N2, D)
    A = rng.randn(D, D)
    L = np.linalg.cholesky(A.dot(A.T) + D * np.eye(D))
    got = U.multiple_mahalanobis(X1, X2, L)
    want = O.mahalanobis_pairs_substitution(X1, X2, L)
    assert got.shape == (N1, N2)
    assert relerr(got, want) < 1e-13
    if N1 * N2 <= 100:
        assert relerr(got, O.mahalanobis_pairs_solve(X1, X2, L)) < 1e-11


@pytest.mark.parametrize("D,K", [(1, 1), (2, 10), (5, 3), (64, 4)])
def test_multiple_pdinv(D, K):
    from gpclust_b200 import utilities as U
    rng = np.random.RandomState(D * 7 + K)
    A = np.zeros((D, D, K))
    for k in range(K):
        B = rng.randn(D, D)
        A[:, :, k] = B.dot(B.T) + D * np.eye(D)
    inv, hld = U.multiple_pdinv(A)
    inv0, hld0 = O.multiple_pdinv(A)
    assert relerr(inv, inv0) < 1e-11
    assert relerr(hld, hld0) < 1e-13


def test_multiple_pdinv_not_pd_raises():
    from gpclust_b200 import utilities as U
    A = np.zeros((2, 2, 1))
    A[:, :, 0] = [[1.0, 2.0], [2.0, -1.0]]
    with pytest.raises(np.linalg.LinAlgError):
        U.multiple_pdinv(A)


def test_multiple_pdinv_jitter_rescue():
    """A singular PSD matrix is rescued by GPy's jitter retry (mean(diag)*1e-6 ...), like gpy_shim.jitchol."""
    from gpclust_b200 import utilities as U
    v = np.array([[1.0, 2.0, 3.0]])
    A = v.T.dot(v)[:, :, None].copy()
    inv, hld = U.multiple_pdinv(A)
    inv0, hld0 = O.multiple_pdinv(A)
    assert relerr(hld, hld0) < 1e-6  # the rescued factor is ill-conditioned by construction


@pytest.mark.parametrize("kind", ["rbf", "matern32", "matern52", "bias", "white", "sum"])
def test_kern_fill(kind):
    from gpclust_b200 import kern as PK
    rng = np.random.RandomState(3)
    X = np.sort(rng.rand(40, 1) * 10.0, 0)
    X[5] = X[4]  # duplicated input (replicates, as in the drosophila design)
    X2 = rng.rand(13, 1) * 10.0
    mk = {
        "rbf": lambda M: M.RBF(1, 1.3, 2.0),
        "matern32": lambda M: M.Matern32(1, 0.7, 1.5),
        "matern52": lambda M: M.Matern52(1, 1.0, 12.0),
        "bias": lambda M: M.Bias(1, 0.3),
        "white": lambda M: M.White(1, 0.05),
        "sum": lambda M: M.Matern52(1, 0.1, 12.0) + M.White(1, 0.05) + M.Bias(1, 0.2),
    }[kind]
    kp, ko = mk(PK), mk(G)
    assert relerr(kp.K(X), ko.K(X)) < 1e-14
    assert np.max(np.abs(kp.K(X, X2) - ko.K(X, X2))) < 1e-14
    Xq = rng.randn(17, 3)
    if kind not in ("bias", "white"):
        assert relerr(kp.K(Xq), ko.K(Xq)) < 1e-13
