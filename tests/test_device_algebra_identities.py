"""The algebra the device posterior rests on, checked in numpy against the oracle's literal Cholesky route
(MOHGP.py:70-90,124-153): with A = Ly^-1 Sf Ly^-T = Q diag(lam) Q^T (the same A for every cluster and iteration),

    V = Ly Q,  U = Ly^-T Q,  c_i = lam_i / (1 + phi_hat_k lam_i),  z = U^T ybar_k,  s_k = U z  (= Sy^-1 ybar_k)

    log_det_diff_k          = sum_i log1p(phi_hat_k lam_i)
    Lambda_inv_k            = V diag(c) V^T - (1e-6 / phi_hat_k) I
    mu_k                    = V (c o z) - (1e-6 / phi_hat_k) s_k
    s_k^T Lambda_inv_k s_k  = sum c z^2 - (1e-6 / phi_hat_k) |s_k|^2
    tr(Lambda_inv_k Sy_inv) = sum c - (1e-6 / phi_hat_k) tr(Sy_inv)

(CPU only: this pins the mathematics, the -m gpu tests pin the kernels that implement it -- k_mohgp_post against the oracle AND
against the literal on-device Cholesky route k_mohgp_cluster.)"""
import numpy as np
import pytest

from helpers import mohgp_problem, oracle_kernels, relerr
from oracle import gpy_shim as G
from oracle.vb_oracle import MOHGPOracle, mahalanobis_whitened_gemm


def _eigen_route(m):
    Ly = m.Sy_chol
    Ly_inv = m.Sy_chol_inv
    A = Ly_inv.dot(m.Sf).dot(Ly_inv.T)
    lam, Q = np.linalg.eigh(0.5 * (A + A.T))
    V, U = Ly.dot(Q), Ly_inv.T.dot(Q)
    out = {"log_det_diff": [], "Lambda_inv": [], "muk": [], "sLs": [], "trLSyi": []}
    for k in range(m.K):
        ph = m.phi_hat[k]
        c = lam / (1.0 + ph * lam)
        z = U.T.dot(m.ybark[:, k])
        s = U.dot(z)
        eps = 1e-6 / ph
        out["log_det_diff"].append(np.sum(np.log1p(ph * lam)))
        out["Lambda_inv"].append((V * c).dot(V.T) - eps * np.eye(m.D))
        out["muk"].append(V.dot(c * z) - eps * s)
        out["sLs"].append(np.sum(c * z * z) - eps * s.dot(s))
        out["trLSyi"].append(np.sum(c) - eps * np.trace(m.Sy_inv))
    return {k: np.array(v) for k, v in out.items()}, lam


@pytest.mark.parametrize("case", ["config3-like", "replicated time points (singular Sf)", "matern"])
def test_eigenbasis_posterior_equals_the_cholesky_route(case):
    if case == "config3-like":
        X, Y, _ = mohgp_problem(600, 24, 6, seed=1)
        kF, kY = oracle_kernels()
    elif case == "matern":
        X, Y, _ = mohgp_problem(300, 16, 4, seed=2)
        kF, kY = G.Matern52(1, 1.0, 3.0), G.Matern32(1, 0.2, 1.5) + G.White(1, 0.05) + G.Bias(1, 0.1)
    else:
        _, Y, _ = mohgp_problem(200, 14, 3, seed=3)
        X = np.repeat(np.arange(2.0, 9.0), 2)[:, None]           # every hour twice: Sf has rank 7 of 14
        kF, kY = G.Matern52(1, 1.0, 12.0), G.Matern52(1, 0.1, 12.0) + G.White(1, 0.05)
    np.random.seed(4)
    m = MOHGPOracle(X, kF, kY, Y, K=(6 if case == "config3-like" else 4), mahalanobis=mahalanobis_whitened_gemm)
    e, lam = _eigen_route(m)
    if case.startswith("replicated"):
        assert np.sum(np.abs(lam) < 1e-9 * lam.max()) >= 6       # the null space of Sf: c_i = 0 there, nothing is inverted
    assert relerr(e["log_det_diff"], m.log_det_diff) < 1e-11
    assert relerr(e["Lambda_inv"], m.Lambda_inv) < 1e-9
    assert relerr(e["muk"].T, m.muk) < 1e-9
    sLs = np.array([m.Syi_ybark[:, k].dot(m.Lambda_inv[k]).dot(m.Syi_ybark[:, k]) for k in range(m.K)])
    assert relerr(e["sLs"], sLs) < 1e-9
    assert relerr(e["sLs"].sum(), np.sum(m.Syi_ybarkybarkT_Syi * m.Lambda_inv)) < 1e-9      # the term of MOHGP.py:133
    tr = np.array([np.sum(m.Lambda_inv[k] * m.Sy_inv) for k in range(m.K)])                  # MOHGP.py:147
    assert relerr(e["trLSyi"], tr) < 1e-9


def test_eigenbasis_kernel_gradient_matrices():
    """B_inv_k = phi_hat_k U diag(1 / (1 + phi_hat_k lam)) U^T, so sum_k B_inv_k is ONE product with the
    summed diagonal -- against the per-cluster triangular solves of MOHGP.py:97-98."""
    X, Y, _ = mohgp_problem(400, 20, 5, seed=5)
    kF, kY = oracle_kernels()
    np.random.seed(6)
    m = MOHGPOracle(X, kF, kY, Y, K=5, mahalanobis=mahalanobis_whitened_gemm)
    Ly_inv = m.Sy_chol_inv
    A = Ly_inv.dot(m.Sf).dot(Ly_inv.T)
    lam, Q = np.linalg.eigh(0.5 * (A + A.T))
    U = Ly_inv.T.dot(Q)
    tmp = [G.dtrtrs(L, m.Sy_chol_inv, lower=1)[0] for L in m._C_chols]
    B_invs = [ph * t.T.dot(t) for ph, t in zip(m.phi_hat, tmp)]
    for k in range(m.K):
        Bk = m.phi_hat[k] * (U / (1.0 + m.phi_hat[k] * lam)).dot(U.T)
        assert relerr(Bk, B_invs[k]) < 1e-10
    dsum = np.sum([m.phi_hat[k] / (1.0 + m.phi_hat[k] * lam) for k in range(m.K)], 0)
    assert relerr((U * dsum).dot(U.T), sum(B_invs)) < 1e-10


def test_mog_feature_expansion_identities():
    """gpclust_b200/MOG.py: with f(x) = [x_i x_j (i <= j) | x_i] of the CENTRED data, phi^T F holds Ck and Xsumk, the quadratic form of
    MOG.py:74-77 is linear in f(x), and centring changes nothing but a shift of `mun` (MOG.py:52-61) -- numpy, against the oracle."""
    from helpers import load_mog_data
    from oracle.vb_oracle import MOGOracle
    rng = np.random.RandomState(0)
    X = np.hstack([load_mog_data()[:120], rng.randn(120, 1) * 3 + 5.0])   # D = 3, far from the origin in one coordinate
    np.random.seed(1)
    m = MOGOracle(X, K=4)
    N, D, K = X.shape[0], X.shape[1], m.K
    mean = X.mean(0)
    Xc = X - mean
    iu = np.triu_indices(D)
    F = np.hstack([Xc[:, iu[0]] * Xc[:, iu[1]], Xc])                      # N x (D(D+1)/2 + D)
    T = m.phi.T.dot(F)                                                    # the S pass's statistics
    # posterior from the statistics of the centred data (m0 shifted too)
    m0c = m.m0 - mean
    for k in range(K):
        Ck = np.zeros((D, D))
        Ck[iu] = T[k, :len(iu[0])]
        Ck = Ck + Ck.T - np.diag(np.diag(Ck))
        Xsum = T[k, len(iu[0]):]
        mun_c = (m.k0 * m0c + Xsum) / m.kNs[k]
        Sn = m.S0 + Ck + m.k0 * np.outer(m0c, m0c) - m.kNs[k] * np.outer(mun_c, mun_c)
        assert relerr(mun_c + mean, m.mun[:, k]) < 1e-12
        assert relerr(Sn, m.Sns[:, :, k]) < 1e-10
        # (x - mu)^T S (x - mu) = sum_{i<=j} (2 - delta_ij) S_ij x_i x_j - 2 (S mu)^T x + mu^T S mu: one row of W plus a bias
        S = m.Sns_inv[:, :, k]
        w = np.concatenate([np.where(iu[0] == iu[1], 1.0, 2.0) * S[iu], -2.0 * S.dot(mun_c)])
        q = F.dot(w) + mun_c.dot(S).dot(mun_c)
        x_m = X - m.mun[:, k]
        assert relerr(q, np.einsum("ni,ij,nj->n", x_m, S, x_m)) < 1e-10


def test_omgp_triangular_inverse_identities():
    """with B = L L^T and W = L^-1, everything OMGP.py:96-147 and :55-94 need comes from W and alpha = W^T (W Y):
    tr(B^-1 Y Y^T) = sum Y o alpha (the reference solves against the N x N YYT), logdet B = -2 sum log diag W, diag(B^-1) = squared
    column norms of W, and dL_dB = B^-1 (Y Y^T - B) B^-1 = alpha alpha^T - B^-1 contracts with any dK/dtheta tile by tile."""
    rng = np.random.RandomState(0)
    N, D = 50, 2
    X = np.sort(rng.uniform(0, 10, N))[:, None]
    Y = np.hstack([np.sin(X), np.cos(X)]) + 0.1 * rng.randn(N, D)
    phi = rng.dirichlet([1.0, 1.0, 1.0], N)[:, 0]
    kern, variance = G.RBF(1, 1.3, 0.9), 0.05
    B = kern.K(X) + np.diag(variance / (phi + 1e-6))
    L = np.linalg.cholesky(B)
    W = np.linalg.inv(L)
    alpha = W.T.dot(W.dot(Y))
    Bi, LB, _, logdet = G.pdinv(B)
    assert abs(np.sum(Y * alpha) - G.dpotrs(LB, Y.dot(Y.T))[0].trace()) < 1e-10 * abs(np.sum(Y * alpha))   # OMGP.py:113
    assert abs(-2.0 * np.sum(np.log(np.diag(W))) - logdet) < 1e-12 * abs(logdet)                            # OMGP.py:114
    assert relerr(np.sum(W * W, 0), np.diag(Bi)) < 1e-10                                                    # OMGP.py:137
    assert relerr(alpha, G.dpotrs(LB, Y)[0]) < 1e-10                                                        # OMGP.py:136
    # OMGP.py:71-77 (with the matrix the reference uses there: no 1e-6)
    B0 = kern.K(X) + np.diag(variance / phi)
    Bi0, LB0, _, _ = G.pdinv(B0)
    tmp = G.dpotrs(LB0, Y.dot(Y.T))[0]
    tmp[np.diag_indices(N)] -= 1.0
    dL_dB = G.dpotrs(LB0, tmp.T)[0]
    a0 = Bi0.dot(Y)
    assert relerr(a0.dot(a0.T) - Bi0, dL_dB) < 1e-9
    # contraction tile by tile (16 x 16 tiles of B^-1 = W^T W formed one at a time) == the dense contraction
    W0 = np.linalg.inv(np.linalg.cholesky(B0))
    dK = kern.K(X) / kern.variance                                        # dK / d variance
    acc = 0.0
    for i in range(0, N, 16):
        for j in range(0, N, 16):
            tile = W0[:, i:i + 16].T.dot(W0[:, j:j + 16])
            acc += np.sum((a0[i:i + 16].dot(a0[j:j + 16].T) - tile) * dK[i:i + 16, j:j + 16])
    kern.update_gradients_full(0.5 * dL_dB, X)
    assert abs(0.5 * acc - kern.variance_gradient) < 1e-9 * abs(kern.variance_gradient)
