"""GPU parity of the MOHGP hot path (through the Python mirror -> C-ABI -> sm_100a kernels) against the
CPU oracle on identical seeded inputs.  Tolerance (north_star): 1e-9 relative in FP64 for phi, bound and
gradients -- arrays are compared in max-norm-relative terms; cluster assignments
after convergence must be identical."""
import io

import numpy as np
import pytest

from helpers import mohgp_problem, oracle_kernels, product_kernels, relerr
from oracle.vb_oracle import MOHGPOracle, mahalanobis_pairs_substitution, mahalanobis_whitened_gemm

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _pair(N, D, K, prior_Z="symmetric", alpha=1.0, seed=0):
    from gpclust_b200 import MOHGP
    X, Y, _ = mohgp_problem(N, D, K, seed)
    np.random.seed(seed + 1)
    phi0 = np.random.randn(N, K)
    kF, kY = oracle_kernels()
    np.random.seed(seed + 1)
    mo = MOHGPOracle(X, kF, kY, Y, K=K, prior_Z=prior_Z, alpha=alpha, mahalanobis=mahalanobis_whitened_gemm if N * K > 40000 else None)
    pF, pY = product_kernels()
    np.random.seed(seed + 1)
    mp = MOHGP(X, pF, pY, Y, K=K, prior_Z=prior_Z, alpha=alpha)
    assert np.array_equal(mp.phi_, phi0) and np.array_equal(mo.phi_, phi0)
    return mo, mp


SHAPES = [(400, 12, 5), (1000, 64, 64), (777, 56, 20), (64, 8, 2), (130, 3, 1), (5000, 33, 17), (300, 64, 9), (5, 6, 3), (9, 2, 2)]


@pytest.mark.parametrize("N,D,K", SHAPES)
@pytest.mark.parametrize("prior_Z", ["symmetric", "DP"])
def test_state_bound_gradient(N, D, K, prior_Z):
    mo, mp = _pair(N, D, K, prior_Z, alpha=1.0 if prior_Z == "symmetric" else 3.0)
    assert relerr(mp.phi, mo.phi) < 1e-13
    assert relerr(mp.Hgrad, mo.Hgrad) < 1e-13
    assert relerr(mp.phi_hat, mo.phi_hat) < 1e-12
    assert abs(mp.H - mo.H) <= 1e-11 * abs(mo.H) + 1e-13
    assert relerr(mp.Sf, mo.Sf) < 1e-14 and relerr(mp.Sy, mo.Sy) < 1e-14
    assert relerr(mp.Sy_chol, mo.Sy_chol) < 1e-10
    assert relerr(mp.Sy_inv, mo.Sy_inv) < 1e-8  # Sy is ill-conditioned; the inverse is compared loosely, its uses tightly
    assert abs(mp.Sy_logdet - mo.Sy_logdet) < 1e-10 * abs(mo.Sy_logdet)
    assert relerr(mp.YTY, mo.YTY) < 1e-12
    assert relerr(mp.ybark, mo.ybark) < 1e-12
    assert relerr(mp.log_det_diff, mo.log_det_diff) < TOL
    assert relerr(mp.muk, mo.muk) < TOL
    assert relerr(mp.Lambda_inv, mo.Lambda_inv) < 1e-7  # the reference's own formula is noisy here
    assert abs(mp.mixing_prop_bound() - mo.mixing_prop_bound()) < 1e-11 * max(1.0, abs(mo.mixing_prop_bound()))
    assert relerr(mp.mixing_prop_bound_grad(), mo.mixing_prop_bound_grad()) < 1e-13
    assert abs(mp.bound() - mo.bound()) < TOL * abs(mo.bound())
    assert abs(mp.log_likelihood() - mo.bound()) < TOL * abs(mo.bound())
    g_p, ng_p = mp.vb_grad_natgrad()
    g_o, ng_o = mo.vb_grad_natgrad()
    assert g_p.shape == (N * K,) and ng_p.shape == (N * K,)
    assert relerr(ng_p, ng_o) < TOL
    assert relerr(g_p, g_o) < TOL


def test_gemm_form_equals_pairwise_substitution():
    """The G pass evaluates the Mahalanobis terms as one GEMM; the reference forward-substitutes per pair
    (utilities.py:147-170).  Same natural gradient to 1e-9 (the row-constant |L^-1 y|^2 is projected out)."""
    mo, mp = _pair(300, 24, 6)
    mo.mahalanobis = mahalanobis_pairs_substitution
    g_o, ng_o = mo.vb_grad_natgrad()
    g_p, ng_p = mp.vb_grad_natgrad()
    assert relerr(ng_p, ng_o) < TOL and relerr(g_p, g_o) < TOL


@pytest.mark.parametrize("method", ["HS", "PR", "FR", "steepest"])
def test_optimize_trace(method):
    N, D, K = 600, 16, 6
    mo, mp = _pair(N, D, K, seed=3)
    mo.out = io.StringIO()
    mo.optimize(method=method, maxiter=40, verbose=False)
    mp.optimize(method=method, maxiter=40, verbose=False)
    to, tp = mo.trace, mp.optimize_trace
    assert [r[0] for r in tp] == [r[0] for r in to]  # identical iteration accounting (incl. rejected steps)
    for (it, b, sq, beta), (_, b0, sq0, beta0) in zip(tp, to):
        assert abs(b - b0) < 1e-9 * abs(b0), (it, b, b0)
        assert abs(sq - sq0) < 1e-6 * max(abs(sq0), 1e-12), (it, sq, sq0)
    assert relerr(mp.phi, mo.phi) < 1e-6
    assert np.array_equal(mp.phi.argmax(1), mo.phi.argmax(1))


def test_converged_assignments_identical():
    N, D, K = 2000, 32, 8
    mo, mp = _pair(N, D, K, seed=5)
    mo.out = io.StringIO()
    mo.optimize(maxiter=200, verbose=False)
    it = mp.optimize(maxiter=200, verbose=False)
    assert it == mo.trace[-1][0]
    assert abs(mp.bound() - mo.bound()) < 1e-9 * abs(mo.bound())
    assert np.array_equal(mp.phi.argmax(1), mo.phi.argmax(1))
    # bound is monotone non-decreasing over accepted iterations
    b = [r[1] for r in mp.optimize_trace]
    assert all(b2 >= b1 - 1e-9 * abs(b1) for b1, b2 in zip(b, b[1:]))


def test_set_get_roundtrip_and_randomize():
    mo, mp = _pair(200, 10, 4)
    x = mp.get_vb_param()
    assert x.shape == (200 * 4,)
    b = mp.bound()
    mp.set_vb_param(x.copy())
    assert mp.bound() == b  # bit-reproducible re-evaluation
    assert np.array_equal(mp.get_vb_param(), x)
    np.random.seed(11)
    mp.randomize()
    np.random.seed(11)
    mo.randomize()
    assert abs(mp.bound() - mo.bound()) < 1e-9 * abs(mo.bound())


def test_tiny_phi_hat_guard():
    """MOHGP.py:85: clusters with phi_hat <= 1e-6 take Lambda_inv = Sf."""
    mo, mp = _pair(150, 8, 3)
    x = mo.get_vb_param().reshape(150, 3).copy()
    x[:, 2] = -60.0
    mo.set_vb_param(x.flatten())
    mp.set_vb_param(x.flatten())
    assert mo.phi_hat[2] < 1e-6
    assert relerr(mp.Lambda_inv[2], mo.Sf) < 1e-14
    assert abs(mp.bound() - mo.bound()) < 1e-9 * abs(mo.bound())
    g_p, ng_p = mp.vb_grad_natgrad()
    g_o, ng_o = mo.vb_grad_natgrad()
    assert relerr(ng_p, ng_o) < TOL


def test_not_pd_raises_linalgerror():
    from gpclust_b200 import MOHGP, kern
    X, Y, _ = mohgp_problem(100, 8, 3)
    with pytest.raises(np.linalg.LinAlgError):
        MOHGP(X, kern.RBF(1, 1.0, 2.0), kern.Bias(1, -1.0), Y, K=3)


def test_matern_dp_drosophila_like():
    """Config 1 stand-in: D=56 time points with replicated inputs, Matern52 kernels, DP prior, K=20."""
    from gpclust_b200 import MOHGP, kern
    from oracle import gpy_shim as G
    import os
    from helpers import GOLDEN
    rows = [l.strip().split(",") for l in open(os.path.join(GOLDEN, "kalinka09_pdata.csv")).read().strip().split("\n")]
    hours = np.array([float(r[2]) for r in rows if r[0] == "me"])  # melanogaster: 8 replicates x hours 2..20 (no header row)
    assert hours.size == 56
    X = hours[:56, None]
    D = X.shape[0]
    rng = np.random.RandomState(0)
    kF, kY = G.Matern52(1, 1.0, 12.0), G.Matern52(1, 0.1, 12.0) + G.White(1, 0.05)
    Sf, Sy = kF.K(X), kY.K(X)
    K, N = 20, 500
    means = np.linalg.cholesky(Sf + 1e-6 * np.eye(D)).dot(rng.randn(D, 6)).T
    Y = means[rng.randint(0, 6, N)] + np.linalg.cholesky(Sy + 1e-8 * np.eye(D)).dot(rng.randn(D, N)).T
    Y = (Y - Y.mean(1)[:, None]) / Y.std(1)[:, None]
    np.random.seed(2)
    mo = MOHGPOracle(X, kF, kY, Y, K=K, prior_Z="DP", alpha=1.0)
    np.random.seed(2)
    mp = MOHGP(X, kern.Matern52(1, 1.0, 12.0).fix(), (kern.Matern52(1, 0.1, 12.0) + kern.White(1, 0.05)).fix(), Y, K=K, prior_Z="DP", alpha=1.0)
    assert abs(mp.bound() - mo.bound()) < TOL * abs(mo.bound())
    mo.out = io.StringIO()
    mo.optimize(maxiter=30, verbose=False)
    mp.optimize(maxiter=30, verbose=False)
    assert [r[0] for r in mp.optimize_trace] == [r[0] for r in mo.trace]
    assert abs(mp.bound() - mo.bound()) < 1e-8 * abs(mo.bound())


@pytest.mark.parametrize("N,D,K", [(1000, 64, 64), (30000, 64, 3), (777, 56, 20), (130, 3, 1), (5000, 33, 17)])
def test_eigenbasis_posterior_equals_literal_cholesky_route(N, D, K):
    """The hot loop evaluates the per-cluster posterior in the eigenbasis of A (gpc_mohgp_post);
    the reference's literal Cholesky / TRSM / SYRK route (gpc_mohgp_cluster, MOHGP.py:79-90) must give the same
    posterior means, G-pass weights and per-cluster scalars -- also at large phi_hat (N/K = 1e4)."""
    mo, mp = _pair(N, D, K)
    dense = mp._dense_posterior()
    assert relerr(mp._muk_t[:K].cpu().numpy(), dense["muk_t"][:K].cpu().numpy()) < TOL
    assert relerr(mp._Wt.cpu().numpy(), dense["Wt"].cpu().numpy()) < TOL
    assert relerr(mp._Syi_ybark_t[:K].cpu().numpy(), dense["s"][:K].cpu().numpy()) < TOL
    pk, pd = mp._per_k[:K].cpu().numpy(), dense["per_k"][:K].cpu().numpy()
    for col in range(4):  # log_det_diff, s^T Lambda_inv s, tr(Lambda_inv Sy_inv), mu^T Sy^-1 mu
        assert relerr(pk[:, col], pd[:, col]) < TOL, col
    assert relerr(mp.muk, mo.muk) < TOL


def test_duplicate_time_points_singular_Sf():
    """Replicated time points (drosophila design) make Sf singular; A then has zero eigenvalues -- never factorised."""
    from gpclust_b200 import MOHGP, kern
    from oracle import gpy_shim as G
    rng = np.random.RandomState(3)
    X = np.repeat(np.arange(2.0, 10.0), 4)[:, None]  # 32 time points, 8 distinct
    D, N, K = X.shape[0], 500, 6
    kF, kY = G.Matern52(1, 1.0, 12.0), G.Matern52(1, 0.1, 12.0) + G.White(1, 0.05)
    Y = rng.randn(N, D)
    np.random.seed(4)
    mo = MOHGPOracle(X, kF, kY, Y, K=K, prior_Z="DP", alpha=2.0)
    np.random.seed(4)
    mp = MOHGP(X, kern.Matern52(1, 1.0, 12.0).fix(), (kern.Matern52(1, 0.1, 12.0) + kern.White(1, 0.05)).fix(), Y, K=K, prior_Z="DP", alpha=2.0)
    assert abs(mp.bound() - mo.bound()) < TOL * abs(mo.bound())
    assert relerr(mp.muk, mo.muk) < TOL
    g_p, ng_p = mp.vb_grad_natgrad()
    g_o, ng_o = mo.vb_grad_natgrad()
    assert relerr(ng_p, ng_o) < TOL and relerr(g_p, g_o) < TOL


@pytest.mark.parametrize("method", ["HS", "PR", "FR", "steepest"])
def test_device_resident_loop_equals_host_driven_loop(method):
    """optimize() runs the loop from the device-resident state (accept / reject, rotation, termination on the device);
    a callback forces the host-driven loop that follows collapsed_vb.py:147-303 line by line.  Same trace, same state."""
    mo, m_dev = _pair(900, 20, 7, seed=5)
    _, m_host = _pair(900, 20, 7, seed=5)
    n_dev = m_dev.optimize(method=method, maxiter=35, verbose=False)
    n_host = m_host.optimize(method=method, maxiter=35, verbose=False, callback=lambda: None)
    assert n_dev == n_host
    td, th = np.array(m_dev.optimize_trace), np.array(m_host.optimize_trace)
    assert td.shape == th.shape and np.array_equal(td[:, 0], th[:, 0])
    assert np.max(np.abs(td[:, 1] - th[:, 1]) / np.abs(th[:, 1])) < 1e-13
    assert np.allclose(td[:, 2:], th[:, 2:], rtol=1e-9, atol=1e-12)
    assert relerr(m_dev.phi_, m_host.phi_) < 1e-12 and relerr(m_dev.phi, m_host.phi) < 1e-12
    assert abs(m_dev.bound() - m_host.bound()) < 1e-13 * abs(m_host.bound())
    # a second optimize() continues from the adopted state on either path
    m_dev.optimize(method=method, maxiter=6, verbose=False, callback=lambda: None)
    m_host.optimize(method=method, maxiter=6, verbose=False)
    assert abs(m_dev.bound() - m_host.bound()) < 1e-12 * abs(m_host.bound())


def test_device_loop_terminates_on_ftol_and_reports_it(capsys):
    mo, mp = _pair(300, 12, 3, seed=7)
    mo.out = io.StringIO()
    mo.optimize(maxiter=400, ftol=1e-3, verbose=False)
    n = mp.optimize(maxiter=400, ftol=1e-3, verbose=True)
    out = capsys.readouterr().out
    assert "vb converged (ftol)" in out and out.count("iteration ") == len(mp.optimize_trace)
    assert [r[0] for r in mp.optimize_trace] == [r[0] for r in mo.trace] and n == mo.trace[-1][0]


def test_unsupported_sizes_fail_loudly():
    from gpclust_b200 import MOHGP
    X, Y, _ = mohgp_problem(100, 8, 2, 0)
    pF, pY = product_kernels()
    with pytest.raises(ValueError):
        MOHGP(X, pF, pY, Y, K=513)


@pytest.mark.parametrize("N,D,K,prior_Z", [(700, 24, 65, "symmetric"), (1500, 64, 128, "DP"), (900, 40, 200, "symmetric"), (8000, 32, 512, "symmetric"),
                                           (90, 16, 70, "symmetric"), (5000, 64, 300, "DP"), (3000, 8, 449, "symmetric")])
def test_more_than_64_clusters(N, D, K, prior_Z):
    """Config 5 of BASELINE.json uses K up to 512: wide mode -- thread-block clusters of KP/64 CTAs, one 64-cluster chunk per CTA,
    row scalars exchanged through distributed shared memory (gpc_natgrad_wide / gpc_step_stats_wide and their loop variants).
    Same parity bar as the K <= 64 path; includes fewer row groups than clusters (N = 90) and every cluster size 2..8."""
    mo, mp = _pair(N, D, K, prior_Z, alpha=1.0 if prior_Z == "symmetric" else 2.0)
    assert relerr(mp.phi, mo.phi) < 1e-13
    assert relerr(mp.phi_hat, mo.phi_hat) < 1e-12
    assert abs(mp.H - mo.H) <= 1e-11 * abs(mo.H)
    assert relerr(mp.ybark, mo.ybark) < 1e-12
    assert relerr(mp.muk, mo.muk) < TOL
    assert abs(mp.bound() - mo.bound()) < TOL * abs(mo.bound())
    g_p, ng_p = mp.vb_grad_natgrad()
    g_o, ng_o = mo.vb_grad_natgrad()
    assert relerr(ng_p, ng_o) < TOL and relerr(g_p, g_o) < TOL
    mo.out = io.StringIO()
    iters = 12 if K < 512 else 6  # few data per cluster make the trajectory chaotic: keep the comparison short
    mo.optimize(maxiter=iters, verbose=False)
    mp.optimize(maxiter=iters, verbose=False)
    to, tp = np.array(mo.trace), np.array(mp.optimize_trace)
    # with a handful of rows per cluster the accept / reject decisions (bound comparisons at near-ties) eventually differ between
    # any two summation orders: the leading rows must agree to the parity bar, the rest only while the decisions coincide
    n = min(len(to), len(tp))
    same = np.cumprod(to[:n, 0] == tp[:n, 0]).astype(bool)
    lead = int(same.sum())
    assert lead >= min(5, n), (to[:, 0], tp[:, 0])
    assert np.max(np.abs(to[:lead, 1] - tp[:lead, 1]) / np.abs(to[:lead, 1])) < (TOL if N >= 20 * K else 1e-7)
    if lead == len(to) == len(tp):
        assert relerr(mp.phi, mo.phi) < 1e-7


def test_wide_mode_equals_per_chunk_launches(monkeypatch):
    """The two implementations of K > 64 (clusters of CTAs vs one launch per 64-cluster chunk around row-coupled elementwise kernels)
    agree to rounding on state, bound and gradients, and the wide device-resident loop reproduces the host-driven one bit for bit."""
    X, Y, _ = mohgp_problem(2500, 48, 150, seed=3)
    np.random.seed(8)
    phi0 = np.random.randn(2500, 150)
    from gpclust_b200 import MOHGP
    models = {}
    for wide in ("1", "0"):
        monkeypatch.setenv("GPCLUST_B200_WIDE", wide)
        pF, pY = product_kernels()
        models[wide] = MOHGP(X, pF, pY, Y, K=150, phi_init=phi0)
    mw, mc = models["1"], models["0"]
    assert mw._wide and not mc._wide
    assert abs(mw.bound() - mc.bound()) < 1e-13 * abs(mc.bound())
    (gw, nw), (gc, nc) = mw.vb_grad_natgrad(), mc.vb_grad_natgrad()
    assert relerr(nw, nc) < 1e-12 and relerr(gw, gc) < 1e-12
    pF, pY = product_kernels()
    monkeypatch.setenv("GPCLUST_B200_WIDE", "1")
    mh = MOHGP(X, pF, pY, Y, K=150, phi_init=phi0)
    mw.optimize(maxiter=20, verbose=False)                          # device-resident loop (graph replay)
    mh.optimize(maxiter=20, verbose=False, callback=lambda: None)   # host-driven loop
    tw, th = np.array(mw.optimize_trace), np.array(mh.optimize_trace)
    assert tw.shape == th.shape and np.array_equal(tw[:, :2], th[:, :2])
    mc.optimize(maxiter=20, verbose=False)
    tc = np.array(mc.optimize_trace)
    assert tc.shape == tw.shape and np.max(np.abs(tc[:, 1] - tw[:, 1]) / np.abs(tw[:, 1])) < 1e-9


def test_double_failure_restores_accepted_point():
    """collapsed_vb.py:170-193 when BOTH the conjugate step and the steepest retry fail to factorise.  kernF = Bias(-c) makes
    C_k = I + phi_hat_k A indefinite as soon as one cluster grows past 1.5x its initial mass, which the very first step does (beta = 0:
    the retry is the same step and fails too).  The reference warns, restores phi_old, counts two evaluations and stops on ftol
    because the bound is unchanged.  The device-resident loop reports GPC_DONE_FAULT; the host must then recompute the state at the
    accepted point (not keep the failed trial's statistics) and finish exactly like the reference and the host-driven loop."""
    import warnings
    from gpclust_b200 import MOHGP, kern
    from gpclust_b200.collapsed_vb import LinAlgWarning
    from oracle import gpy_shim as G
    N, D, K = 500, 16, 5
    X, Y, _ = mohgp_problem(N, D, K, 2)
    kY = G.RBF(1, 0.1, 1.0) + G.White(1, 0.05)
    L = np.linalg.cholesky(kY.K(X) + 1e-6 * np.eye(D))
    v = np.sum(np.linalg.solve(L, np.ones(D)) ** 2)
    np.random.seed(3)
    phi0 = np.random.randn(N, K)
    e = np.exp(phi0 - phi0.max(1)[:, None])
    c = 1.0 / (v * 1.5 * (e / e.sum(1)[:, None]).sum(0).max())
    np.random.seed(3)
    mo = MOHGPOracle(X, G.Bias(1, -c), kY, Y, K=K)
    mo.out = io.StringIO()
    b0 = mo.bound()
    mo.optimize(maxiter=30, verbose=False)
    assert [r[0] for r in mo.trace] == [2] and mo.trace[0][1] == b0  # the oracle takes the double-failure path
    res = {}
    for name, kw in (("dev", {}), ("host", {"callback": lambda: None})):
        m = MOHGP(X, kern.Bias(1, -c).fix(), (kern.RBF(1, 0.1, 1.0) + kern.White(1, 0.05)).fix(), Y, K=K, phi_init=phi0)
        bp = m.bound()
        assert abs(bp - b0) < TOL * abs(b0)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            n = m.optimize(maxiter=30, verbose=False, **kw)
        assert any(issubclass(x.category, LinAlgWarning) for x in w), name
        assert n == 2 and [r[0] for r in m.optimize_trace] == [2], (name, n, m.optimize_trace)
        assert m.bound() == bp and m.optimize_trace[0][1] == bp and np.array_equal(m.phi_, phi0)  # at the accepted point, state recomputed
        assert abs(m.optimize_trace[0][2] - mo.trace[0][2]) < 1e-7 * abs(mo.trace[0][2]) and m.optimize_trace[0][3] == 0.0
        # the model is usable afterwards: same gradient as the oracle at the (unchanged) point
        g_p, ng_p = m.vb_grad_natgrad()
        g_o, ng_o = mo.vb_grad_natgrad()
        assert relerr(ng_p, ng_o) < TOL and relerr(g_p, g_o) < TOL
        res[name] = m.optimize_trace
    assert res["dev"] == res["host"]


def test_predict_components_against_oracle():
    """MOHGP.predict_components (MOHGP.py:156-174) after optimisation, D = 64 time points: K mean vectors and K covariance matrices
    at new inputs inside and outside the design range."""
    mo, mp = _pair(1500, 64, 6, seed=4)
    mo.out = io.StringIO()
    mo.optimize(maxiter=15, verbose=False)
    mp.optimize(maxiter=15, verbose=False)
    Xnew = np.linspace(-2.0, 12.0, 23)[:, None]
    mu_o, var_o = mo.predict_components(Xnew)
    mu_p, var_p = mp.predict_components(Xnew)
    assert len(mu_p) == len(var_p) == 6 and mu_p[0].shape == (23,) and var_p[0].shape == (23, 23)
    assert relerr(np.array(mu_p), np.array(mu_o)) < 1e-7
    assert relerr(np.array(var_p), np.array(var_o)) < 1e-7


@pytest.mark.parametrize("D,kind", [(64, "psd"), (64, "plusminus"), (33, "psd_singular"), (7, "plusminus"), (56, "clustered")])
def test_eigensolver_reconstructs_A(D, kind):
    """gpc_mohgp_eig: A = Q diag(lam) Q^T.  Stage 1 is a one-sided (Hestenes) Jacobi, exact for positive semi-definite A (every A of this
    path: A = Ly^-1 Sf Ly^-T) but blind to +lam / -lam pairs; its result is residual-checked and a two-sided Jacobi is the fall-back.
    Covers: PSD, singular PSD (replicated time points), clustered spectrum, and an indefinite matrix WITH +lam / -lam pairs."""
    import ctypes
    import torch
    from gpclust_b200._cabi import lib, check, ptr
    rng = np.random.RandomState(D)
    Qr, _ = np.linalg.qr(rng.randn(D, D))
    if kind == "psd":
        lam = rng.uniform(0.01, 5.0, D)
    elif kind == "psd_singular":
        lam = np.concatenate([rng.uniform(0.1, 3.0, D // 4), np.zeros(D - D // 4)])
    elif kind == "clustered":
        lam = np.concatenate([np.full(D // 2, 2.0), 2.0 + 1e-9 * rng.rand(D - D // 2)])
    else:
        half = rng.uniform(0.5, 2.0, D // 2)
        lam = np.concatenate([half, -half, np.zeros(D - 2 * (D // 2))])
    A = (Qr * lam).dot(Qr.T)
    A = 0.5 * (A + A.T)
    dev = torch.device("cuda", torch.cuda.current_device())
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    eye = np.eye(D)
    ws = torch.zeros(int(lib.gpc_mohgp_eig_ws_len(D)), dtype=torch.float64, device=dev)
    A_d, I_d = t(A), t(eye)  # kept alive across the (asynchronous) launch
    check(lib.gpc_mohgp_eig(ptr(A_d), ptr(I_d), ptr(I_d), ptr(I_d), ptr(I_d), D, ptr(ws), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
          "gpc_mohgp_eig")
    w = ws.cpu().numpy()
    sweeps, resid = w[64 + 4 * D * D + 2], w[64 + 4 * D * D + 3]
    lam_d, U = w[:D], w[64:64 + D * D].reshape(D, D)  # Ly = I: U = Q
    assert np.max(np.abs(U.T.dot(U) - eye)) < 1e-13
    assert np.max(np.abs((U * lam_d).dot(U.T) - A)) < 1e-12 * max(1.0, np.abs(A).max())
    assert np.max(np.abs(np.sort(lam_d) - np.sort(lam))) < 1e-12 * max(1.0, np.abs(lam).max())
    assert sweeps < 30 and resid < 1e-12, (sweeps, resid)
