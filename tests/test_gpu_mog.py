"""GPU parity of the MOG path: oracle comparison on seeded inputs and the reference's seeded notebook
trace (tests/golden/mog_notebook_golden.json) replayed THROUGH the GPU path, including try_split /
systematic_splits (K: 4 -> 5 -> ... -> 11) and the Student-t predictive densities."""
import io
import re

import numpy as np
import pytest

from helpers import load_golden, load_mog_data, relerr
from oracle.vb_oracle import MOGOracle

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _pair(X, K, prior_Z="symmetric", alpha=1.0, seed=0, **kw):
    from gpclust_b200 import MOG
    np.random.seed(seed)
    mo = MOGOracle(X, K=K, prior_Z=prior_Z, alpha=alpha, **kw)
    np.random.seed(seed)
    mp = MOG(X, K=K, prior_Z=prior_Z, alpha=alpha, **kw)
    return mo, mp


@pytest.mark.parametrize("N,D,K", [(506, 2, 10), (100, 1, 2), (1000, 3, 7), (333, 5, 20), (2000, 9, 33), (64, 2, 64), (3000, 2, 70), (4000, 4, 130)])
@pytest.mark.parametrize("prior_Z", ["symmetric", "DP"])
def test_state_bound_gradient(N, D, K, prior_Z):
    rng = np.random.RandomState(N + D)
    centres = rng.randn(4, D) * 3.0 + 5.0  # off-centre data: exercises the centring
    X = centres[rng.randint(0, 4, N)] + rng.randn(N, D) * 0.5
    mo, mp = _pair(X, K, prior_Z, alpha=2.0)
    assert relerr(mp.phi, mo.phi) < 1e-13
    assert relerr(mp.phi_hat, mo.phi_hat) < 1e-12
    assert relerr(mp.kNs, mo.kNs) < 1e-12 and relerr(mp.vNs, mo.vNs) < 1e-12
    assert relerr(mp.Xsumk, mo.Xsumk) < 1e-11
    assert relerr(mp.mun, mo.mun) < 1e-11
    assert relerr(mp.Sns, mo.Sns) < 1e-9
    assert relerr(mp.Sns_inv, mo.Sns_inv) < 1e-8
    assert relerr(mp.Sns_halflogdet, mo.Sns_halflogdet) < TOL
    assert abs(mp.bound() - mo.bound()) < TOL * abs(mo.bound())
    g_p, ng_p = mp.vb_grad_natgrad()
    g_o, ng_o = mo.vb_grad_natgrad()
    assert relerr(ng_p, ng_o) < TOL
    assert relerr(g_p, g_o) < TOL


def test_priors_passed_through():
    rng = np.random.RandomState(1)
    X = rng.randn(200, 2)
    kw = dict(prior_m=np.array([0.3, -0.2]), prior_kappa=0.5, prior_S=np.array([[2.0, 0.3], [0.3, 1.0]]), prior_v=5.0)
    mo, mp = _pair(X, 3, **kw)
    assert abs(mp.bound() - mo.bound()) < TOL * abs(mo.bound())
    assert relerr(mp.vb_grad_natgrad()[1], mo.vb_grad_natgrad()[1]) < TOL


def _check_trace(tp, gold_rows, tol_bound, tol_grad):
    assert [r[0] for r in tp] == [g["iteration"] for g in gold_rows]
    for (it, b, sq, beta), g in zip(tp, gold_rows):
        assert abs(b - float(g["bound"])) <= tol_bound * abs(float(g["bound"])), (it, b, g["bound"])
        assert abs(sq - float(g["grad"])) <= tol_grad * abs(float(g["grad"])), (it, sq, g["grad"])


def test_notebook_golden_replay_on_gpu(capsys):
    """notebooks/mixture of Gaussians.ipynb cells [0]-[23] through the GPU path."""
    from gpclust_b200 import MOG
    gold = load_golden()
    X = load_mog_data()
    np.random.seed(0)
    m = MOG(X, K=10)
    m.optimize(verbose=False)
    _check_trace(m.optimize_trace, gold["8"]["trace"], 1e-9, 1e-6)

    m = MOG(X, K=4, prior_Z="DP", alpha=10.0)
    m.optimize(verbose=False)
    _check_trace(m.optimize_trace, gold["11"]["trace"], 1e-9, 1e-6)

    traces = []
    orig = m.optimize

    def rec(*a, **k):
        r = orig(*a, **k)
        traces.extend(m.optimize_trace)
        return r
    m.optimize = rec
    assert m.try_split(1, verbose=False) is True
    m.optimize = orig
    assert m.K == 5
    _check_trace(traces, gold["13"]["trace"], 1e-7, 1e-3)
    assert abs(m.last_split_increase - 58.7935883981) < 1e-5

    m.systematic_splits(verbose=False)
    assert m.K == 11
    msgs = " ".join(gold["21"]["messages"])
    nums = [float(v) for v in re.findall(r"-?\d+\.\d+", msgs)]
    np.testing.assert_allclose(m.mun[:, 0], nums[:2], atol=1e-7)
    np.testing.assert_allclose(m.Sns[:, :, 0].ravel(), nums[2:6], atol=1e-7)
    dens = float(re.findall(r"-?\d+\.\d+", gold["23"]["messages"][0])[0])
    tp = np.array([[0.1, 0.0]])
    np.testing.assert_allclose(m.predict(tp), [dens], atol=1e-6)
    txt = " ".join(gold["23"]["messages"][1:])
    gold_cw = [float(v) for v in re.findall(r"\d\.\d+e[+-]\d+", txt)]
    np.testing.assert_allclose(m.predict_components(tp).round(3).ravel(), gold_cw, atol=1e-9)


def test_verbose_print_format(capsys):
    from gpclust_b200 import MOG
    X = load_mog_data()
    np.random.seed(0)
    m = MOG(X, K=3)
    m.optimize(maxiter=3)
    out = capsys.readouterr().out
    assert re.search(r"iteration 1 bound=\S+ grad=\S+, beta=0", out)
    assert "maxiter exceeded" in out


def test_remove_empty_and_reorder():
    X = load_mog_data()
    mo, mp = _pair(X, 6, "DP", 1.0, seed=4)
    for m in (mo, mp):
        x = m.get_vb_param().reshape(-1, 6).copy()
        x[:, 4] = -80.0
        m.set_vb_param(x.flatten())
        m.remove_empty_clusters()
        m.reorder()
    assert mp.K == mo.K == 5
    assert relerr(mp.phi_hat, mo.phi_hat) < 1e-12
    assert np.all(np.diff(mp.phi_hat) <= 0)
    assert abs(mp.bound() - mo.bound()) < TOL * abs(mo.bound())
