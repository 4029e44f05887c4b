"""Seeded random shapes through the CUDA path against the CPU oracle: ragged row counts (not multiples of 8 / 64), every
padded cluster width (8, 16, 32, 64 and the wide 128..512), D from 1 to 64, both priors, every beta rule.  State, bound,
gradients to the parity bar (1e-9, north_star); two loop iterations through the device-resident loop."""
import io

import numpy as np
import pytest

from helpers import mohgp_problem, oracle_kernels, product_kernels, relerr
from oracle.vb_oracle import MOGOracle, MOHGPOracle, mahalanobis_whitened_gemm

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _cases(n, seed, kmax):
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        K = int(rng.choice([1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 33, 63, 64, 65, 100, 129, 191, 257, 320, 385, 450, 512]))
        K = min(K, kmax)
        N = int(rng.choice([1, 3, 7, 8, 9, 63, 64, 65, 127, 200, 511, 513, 1000, 1777]))
        D = int(rng.choice([1, 2, 5, 8, 13, 31, 32, 33, 56, 64]))
        out.append((N, D, K, str(rng.choice(["symmetric", "DP"])), str(rng.choice(["HS", "PR", "FR", "steepest"])), float(rng.choice([0.5, 1.0, 3.0]))))
    return out


@pytest.mark.parametrize("N,D,K,prior_Z,method,alpha", _cases(28, 2024, 512))
def test_mohgp_random_shape(N, D, K, prior_Z, method, alpha):
    from gpclust_b200 import MOHGP
    X, Y, _ = mohgp_problem(N, D, max(1, min(K, 6)), seed=N + D + K)
    np.random.seed(5)
    phi0 = np.random.randn(N, K)
    kF, kY = oracle_kernels()
    mo = MOHGPOracle(X, kF, kY, Y, K=K, prior_Z=prior_Z, alpha=alpha, mahalanobis=mahalanobis_whitened_gemm)
    mo.set_vb_param(phi0.flatten())
    pF, pY = product_kernels()
    mp = MOHGP(X, pF, pY, Y, K=K, prior_Z=prior_Z, alpha=alpha, phi_init=phi0)
    assert relerr(mp.phi, mo.phi) < 1e-12
    assert relerr(mp.phi_hat, mo.phi_hat) < 1e-11
    assert abs(mp.bound() - mo.bound()) < TOL * abs(mo.bound())
    (g_p, ng_p), (g_o, ng_o) = mp.vb_grad_natgrad(), mo.vb_grad_natgrad()
    assert relerr(ng_p, ng_o) < TOL and relerr(g_p, g_o) < TOL
    mo.out = io.StringIO()
    mo.optimize(method=method, maxiter=3, verbose=False)
    mp.optimize(method=method, maxiter=3, verbose=False)
    to, tp = np.array(mo.trace), np.array(mp.optimize_trace)
    assert to.shape == tp.shape and np.array_equal(to[:, 0], tp[:, 0])
    assert np.max(np.abs(to[:, 1] - tp[:, 1]) / np.abs(to[:, 1])) < 1e-8


@pytest.mark.parametrize("N,D,K,prior_Z,method,alpha", [c for c in _cases(40, 77, 200) if c[1] <= 9][:12])
def test_mog_random_shape(N, D, K, prior_Z, method, alpha):
    from gpclust_b200 import MOG
    rng = np.random.RandomState(N * 7 + D)
    X = rng.randn(N, D) + 3.0 * rng.randint(0, 3, (N, 1))
    np.random.seed(6)
    phi0 = np.random.randn(N, K)
    mo = MOGOracle(X, K=K, prior_Z=prior_Z, alpha=alpha)
    mo.set_vb_param(phi0.flatten())
    mp = MOG(X, K=K, prior_Z=prior_Z, alpha=alpha, phi_init=phi0)
    assert relerr(mp.phi, mo.phi) < 1e-12
    assert abs(mp.bound() - mo.bound()) < TOL * abs(mo.bound())
    (g_p, ng_p), (g_o, ng_o) = mp.vb_grad_natgrad(), mo.vb_grad_natgrad()
    assert relerr(ng_p, ng_o) < TOL and relerr(g_p, g_o) < TOL
    mo.out = io.StringIO()
    mo.optimize(method=method, maxiter=3, verbose=False)
    mp.optimize(method=method, maxiter=3, verbose=False)
    to, tp = np.array(mo.trace), np.array(mp.optimize_trace)
    assert to.shape == tp.shape and np.array_equal(to[:, 0], tp[:, 0])
    assert np.max(np.abs(to[:, 1] - tp[:, 1]) / np.abs(to[:, 1])) < 1e-8
