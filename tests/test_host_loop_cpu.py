"""The PRODUCT's host code that needs no GPU, checked on CPU:

* `gpclust_b200.collapsed_vb.CollapsedVB.optimize` / `_cg_iteration` (the host-driven loop used with `callback=` or free kernel
  parameters; collapsed_vb.py:63-310) -- its four device hooks are implemented here in numpy on top of the oracle's MOG model,
  and the loop must print the reference notebook's stored iteration lines (numbering incl. the +2 jumps of rejected steps,
  values, termination message);
* the slice of GPy / paramz the host mirror carries: the Logexp transform behind `optimizer_array`, `_objective_grads`,
  `optimize_parameters` (collapsed_vb.py:313-323);
* the host side of `gpclust_b200.kern`: `update_gradients_full` of every covariance part against finite differences of an
  independent evaluation of K (the oracle's shim), `param_list` / `fix` bookkeeping.
"""
import numpy as np
import pytest
from numpy.linalg import LinAlgError

from helpers import load_golden, load_mog_data
from oracle import gpy_shim as G
from oracle.vb_oracle import MOGOracle
from test_oracle_golden import _assert_trace

from gpclust_b200 import kern
from gpclust_b200.collapsed_vb import CollapsedVB, LinAlgWarning


class _NumpyHooks(CollapsedVB):
    """CollapsedVB with the device hooks done in numpy (the arithmetic of collapsed_vb.py:152-172 on host vectors)."""

    def __init__(self, model, fail_at=()):
        CollapsedVB.__init__(self, "numpy-hooks")
        self.m = model
        self.sd_old, self.prev, self.fail_at, self.n_set = 0.0, None, set(fail_at), 0

    def bound(self):
        return self.m.bound()

    def _set(self, x):
        self.n_set += 1
        if self.n_set in self.fail_at:
            raise LinAlgError("not positive definite, even with jitter.")
        self.m.set_vb_param(x)

    def _cg_gradient(self, with_old):
        g, ng = self.m.vb_grad_natgrad()
        self.g, self.ng = -g, -ng
        self.x_acc = self.m.get_vb_param().copy()
        self.sq = float(np.dot(self.ng, self.g))
        if with_old:
            self.num = float(np.dot(self.ng - self.prev[1], self.g))
            self.den = float(np.dot(self.ng - self.prev[1], self.prev[0]))

    def _cg_trial(self, beta_mode, step_length, sq_old):
        if beta_mode in ("zero", "steepest_retry"):
            beta = 0.0
        elif beta_mode == "FR":
            beta = self.sq / sq_old
        elif beta_mode == "PR":
            beta = self.num / sq_old
        else:
            beta = self.num / self.den
        if np.isnan(beta) or beta < 0.0:
            beta = 0.0
        self.sd = -self.ng + beta * self.sd_old
        try:
            self._set(self.x_acc + step_length * self.sd)
            return self.m.bound(), self.sq, beta, False
        except LinAlgError:
            return np.nan, self.sq, beta, True

    def _cg_restore(self):
        self.m.set_vb_param(self.x_acc)
        return self.m.bound()

    def _cg_accept(self):
        self.sd_old, self.prev = self.sd, (self.g, self.ng)


def test_host_loop_prints_the_notebook_traces(capsys):
    gold = load_golden()
    X = load_mog_data()
    np.random.seed(0)
    h = _NumpyHooks(MOGOracle(X, K=10))
    n = h.optimize()
    _assert_trace(h.optimize_trace, gold["8"]["trace"], 2e-11, 2e-9, 2e-9)
    assert n == 65 and h.optimize_trace[-1][0] == 65
    out = capsys.readouterr().out
    assert out.count("iteration ") == 61 and out.rstrip().endswith("vb converged (ftol)")
    assert "iteration 1 bound=840.2460148" in out

    h = _NumpyHooks(MOGOracle(X, K=4, prior_Z="DP", alpha=10.0))  # continues the global numpy stream, as the notebook does
    h.optimize(verbose=False)
    _assert_trace(h.optimize_trace, gold["11"]["trace"], 2e-11, 2e-9, 2e-9)
    assert capsys.readouterr().out == ""


@pytest.mark.parametrize("method", ["steepest", "FR", "PR", "HS"])
def test_host_loop_equals_the_oracle_loop_for_every_method(method):
    X = load_mog_data()[:150]
    np.random.seed(5)
    a = MOGOracle(X, K=5)
    b = MOGOracle(X, K=5)
    b.set_vb_param(a.get_vb_param().copy())
    import io
    a.out = io.StringIO()
    a.optimize(method=method, maxiter=25, verbose=False)
    h = _NumpyHooks(b)
    h.optimize(method=method, maxiter=25, verbose=False)
    assert [r[0] for r in h.optimize_trace] == [r[0] for r in a.trace]
    np.testing.assert_allclose(np.array(h.optimize_trace)[:, 1:], np.array(a.trace)[:, 1:], rtol=1e-9, atol=1e-12)


def test_maxiter_counts_bound_evaluations_and_callback_runs_once_per_pass(capsys):
    X = load_mog_data()[:100]
    np.random.seed(2)
    h = _NumpyHooks(MOGOracle(X, K=3))
    calls = []
    n = h.optimize(maxiter=7, ftol=-1.0, gtol=-1.0, callback=lambda: calls.append(1))
    assert n >= 7 and n == h.optimize_trace[-1][0]
    assert len(calls) == len(h.optimize_trace)          # collapsed_vb.py:147-150: once per pass of the loop body
    assert capsys.readouterr().out.rstrip().endswith("maxiter exceeded")
    with pytest.raises(AssertionError):
        h.optimize(method="BFGS")                       # collapsed_vb.py:110


def test_failed_trial_points_follow_collapsed_vb_182_193():
    """A LinAlgError at the conjugate trial point -> steepest retry; at both -> warning, old parameters, two evaluations."""
    X = load_mog_data()[:100]
    np.random.seed(2)
    h = _NumpyHooks(MOGOracle(X, K=3), fail_at={2})     # second trial point of the run fails to factorise
    h.optimize(maxiter=3, ftol=-1.0, gtol=-1.0, verbose=False)
    assert [r[0] for r in h.optimize_trace][:2] == [1, 3]
    np.random.seed(2)
    h = _NumpyHooks(MOGOracle(X, K=3), fail_at={2, 3})  # ... and so does the steepest retry
    with pytest.warns(LinAlgWarning):
        h.optimize(maxiter=3, ftol=-1.0, gtol=-1.0, verbose=False)
    (it1, b1, _, _), (it2, b2, _, _) = h.optimize_trace[:2]
    assert (it1, it2) == (1, 3) and b2 == b1            # the bound of the restored point is the old bound


def test_logexp_transform_round_trip_and_gradient_factor():
    f = np.array([1e-8, 1e-3, 0.5, 1.0, 20.0, 35.9, 36.1, 500.0])
    x = CollapsedVB._logexp_finv(f)
    np.testing.assert_allclose(CollapsedVB._logexp_f(x), f, rtol=1e-12)
    assert np.all(np.isfinite(CollapsedVB._logexp_f(np.array([-1e4, -745.0, 0.0, 1e4]))))
    # d f / d x = 1 - exp(-f)  (paramz Logexp.gradfactor), used by _objective_grads
    xx = np.array([-3.0, 0.0, 2.0])
    fd = (CollapsedVB._logexp_f(xx + 1e-6) - CollapsedVB._logexp_f(xx - 1e-6)) / 2e-6
    np.testing.assert_allclose(fd, -np.expm1(-CollapsedVB._logexp_f(xx)), rtol=1e-8)


class _Quadratic(CollapsedVB):
    """bound(theta) = -sum (theta - target)^2 over two free positive parameters and one fixed one."""

    def __init__(self):
        CollapsedVB.__init__(self, "q")
        self.k = kern.RBF(1, 2.0, 3.0) + kern.White(1, 0.5).fix()
        self.target = np.array([0.7, 1.9])
        self.changes = 0

    def _kernel_param_list(self):
        return self.k.param_list()

    def parameters_changed(self):
        self.changes += 1
        theta = np.array([getattr(o, n) for o, n in self._kernel_param_list()])
        for (o, n), t, v in zip(self._kernel_param_list(), self.target, theta):
            setattr(o, n + "_gradient", -2.0 * (v - t))

    def bound(self):
        theta = np.array([getattr(o, n) for o, n in self._kernel_param_list()])
        return -float(np.sum((theta - self.target) ** 2))


def test_optimize_parameters_is_lbfgs_on_the_logexp_vector():
    q = _Quadratic()
    assert [n for _, n in q._kernel_param_list()] == ["variance", "lengthscale"]  # the fixed White variance is not a parameter
    np.testing.assert_allclose(CollapsedVB._logexp_f(q.optimizer_array), [2.0, 3.0], rtol=1e-12)
    q.parameters_changed()
    x0 = q.optimizer_array.copy()
    f0, g0 = q._objective_grads(x0)
    for i in range(2):
        e = np.zeros(2)
        e[i] = 1e-6
        fd = (q._objective_grads(x0 + e)[0] - q._objective_grads(x0 - e)[0]) / 2e-6
        assert abs(fd - g0[i]) < 1e-6 * max(1.0, abs(fd))
    q.optimizer_array = x0
    gain = q.optimize_parameters()
    assert gain > 0 and q.bound() > -1e-6
    np.testing.assert_allclose([q.k.parts[0].variance, q.k.parts[0].lengthscale], q.target, atol=1e-3)
    assert q.k.parts[1].variance == 0.5
    n = q.changes
    q.optimizer_array = q.optimizer_array  # unchanged values do not fire parameters_changed
    assert q.changes == n


@pytest.mark.parametrize("make", [lambda K: K.RBF(1, 0.7, 1.3), lambda K: K.Matern32(1, 1.2, 0.8), lambda K: K.Matern52(1, 0.4, 2.1),
                                  lambda K: K.Bias(1, 0.3), lambda K: K.White(1, 0.2),
                                  lambda K: K.RBF(1, 0.7, 1.3) + K.Matern52(1, 0.4, 2.1) + K.White(1, 0.2) + K.Bias(1, 0.1)])
def test_kernel_gradients_against_finite_differences_of_the_shim(make):
    """`update_gradients_full` (host D x D algebra of the product's kernels) == d/dtheta sum(dL_dK * K) with K evaluated by the
    oracle's independent shim; the same kernel built from the shim's classes gives the same numbers."""
    rng = np.random.RandomState(0)
    X = np.sort(rng.uniform(0, 5, 9))[:, None]
    X[3] = X[2]                                   # duplicated inputs (replicates), as in the kalinka design
    dL = rng.randn(9, 9)
    kp, ks = make(kern), make(G)
    kp.update_gradients_full(dL, X)
    ks.update_gradients_full(dL, X)
    plist, slist = kp.param_list(), ks.param_list()
    assert [n for _, n in plist] == [n for _, n in slist]
    for (o, n), (so, sn) in zip(plist, slist):
        v0 = getattr(so, sn)
        h = 1e-6 * max(1.0, abs(v0))
        setattr(so, sn, v0 + h)
        up = np.sum(dL * ks.K(X))
        setattr(so, sn, v0 - h)
        dn = np.sum(dL * ks.K(X))
        setattr(so, sn, v0)
        fd = (up - dn) / (2 * h)
        assert abs(getattr(o, n + "_gradient") - fd) < 1e-7 * max(1.0, abs(fd)), (n, getattr(o, n + "_gradient"), fd)
        assert abs(getattr(o, n + "_gradient") - getattr(so, sn + "_gradient")) < 1e-12 * max(1.0, abs(fd))
    # cross-covariance: White contributes nothing (GPy), the others use X2
    X2 = rng.uniform(0, 5, (4, 1))
    kp.update_gradients_full(dL[:, :4], X, X2)
    ks.update_gradients_full(dL[:, :4], X, X2)
    for (o, n), (so, sn) in zip(plist, slist):
        assert abs(getattr(o, n + "_gradient") - getattr(so, sn + "_gradient")) < 1e-12 * max(1.0, abs(getattr(so, sn + "_gradient")))


def test_kernel_bookkeeping():
    k = kern.RBF(1, 1.0, 2.0) + (kern.Matern32(1, 0.5, 1.0) + kern.White(1, 0.05))
    assert [s[0] for s in k.parts_spec()] == ["rbf", "matern32", "white"] and len(k.parts) == 3  # nested sums are flattened
    np.testing.assert_allclose(k.param_array, [1.0, 2.0, 0.5, 1.0, 0.05])
    assert len(k.param_list()) == 5
    k.parts[1].fix()
    assert [n for _, n in k.param_list()] == ["variance", "lengthscale", "variance"]
    assert k.fix().param_list() == [] and len(k.unfix().param_list()) == 5
    c = k.copy()
    c.parts[0].variance = 9.0
    assert k.parts[0].variance == 1.0
    with pytest.raises(NotImplementedError):
        kern.RBF(2, 1.0, [1.0, 2.0], ARD=True)
    with pytest.raises(ValueError):
        kern.Add([kern.Bias(1, 1.0)] * 9)
