"""Generate golden vectors by running the REFERENCE'S OWN Python modules (unmodified, imported from
/root/reference) on small seeded inputs.  Run in the build container only; the fixtures it writes
(tests/golden/reference_*.npz) are committed and are what travels to the GPU box.

The reference imports GPy at module level (not installable here).  Only thin pieces of GPy are touched on
the hot path, so a stub `GPy` package is placed in sys.modules before the import:

    GPy.core.Model                        plain base class: link_parameter(s) / unlink_parameter set / drop
                                          attributes, optimizer_array is empty (kernel hyper-parameters fixed,
                                          so optimize_parameters() returns 0. exactly as collapsed_vb.py:322-323)
    GPy.core.parameterization.param.Param, ...transformations.Logexp
                                          a (1,) float array with a name -- what GPy's scalar Param is
    GPy.util.linalg.*, GPy.util.diag.subtract, GPy.kern.*
                                          oracle/gpy_shim.py (thin LAPACK wrappers / closed-form covariances)

Everything else that executes -- CollapsedVB.optimize, CollapsedMixture, MOHGP, MOG, OMGP, np_utilities (the
Python-3 fallbacks, since `scipy.weave` does not exist) -- is the reference's code.
"""
import contextlib
import io
import json
import os
import re
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GPCLUST_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import gpy_shim as G  # noqa: E402


def install_gpy_stub():
    class Model(object):
        def __init__(self, name):
            self.name = name
            self._linked = []

        def link_parameter(self, p):
            self._linked.append(p)
            if isinstance(p, Param):
                setattr(self, p.name, p)

        def link_parameters(self, *ps):
            for p in ps:
                self.link_parameter(p)

        def unlink_parameter(self, p):
            self._linked = [q for q in self._linked if q is not p]

        @property
        def optimizer_array(self):
            return np.zeros(0)

        @optimizer_array.setter
        def optimizer_array(self, value):  # collapsed_mixture.py:174 restores it after a failed split: nothing is free here
            assert np.asarray(value).size == 0

        def randomize(self):
            pass

    class Param(np.ndarray):
        def __new__(cls, name, value, *transforms):
            obj = np.atleast_1d(np.asarray(value, dtype=np.float64)).view(cls)
            obj.name = name
            return obj

        def __array_finalize__(self, obj):
            self.name = getattr(obj, "name", None)

    class Logexp(object):
        pass

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def subtract(A, b):  # GPy.util.diag.subtract: in-place on the diagonal
        A[np.diag_indices(A.shape[0])] -= b
        return A

    linalg = mod("GPy.util.linalg", mdot=G.mdot, pdinv=G.pdinv, backsub_both_sides=G.backsub_both_sides, dpotrs=G.dpotrs,
                 jitchol=G.jitchol, dtrtrs=G.dtrtrs, dpotri=G.dpotri, tdot_numpy=G.tdot, tdot=G.tdot)
    diag = mod("GPy.util.diag", subtract=subtract)
    util = mod("GPy.util", linalg=linalg, diag=diag)
    param = mod("GPy.core.parameterization.param", Param=Param)
    transformations = mod("GPy.core.parameterization.transformations", Logexp=Logexp)
    parameterization = mod("GPy.core.parameterization", param=param, transformations=transformations)
    model = mod("GPy.core.model", Model=Model)
    core = mod("GPy.core", Model=Model, model=model, parameterization=parameterization)
    kern = mod("GPy.kern", RBF=G.RBF, Matern32=G.Matern32, Matern52=G.Matern52, Bias=G.Bias, White=G.White)
    mod("GPy", core=core, util=util, kern=kern)


LINE = re.compile(r"iteration (\d+) bound=(\S+) grad=(\S+), beta=(\S+)")


def run_optimize(m, **kw):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        m.optimize(verbose=True, **kw)
    rows = [(int(a), float(b), float(c), float(d)) for a, b, c, d in LINE.findall(buf.getvalue().replace("\r", "\n"))]
    return np.array(rows, dtype=np.float64).reshape(-1, 4)


def mohgp_case(GPclust, name, N, D, K, prior_Z, alpha, seed, kernF, kernY, X=None, method="HS", maxiter=25):
    rng = np.random.RandomState(seed)
    X = np.linspace(0, 10, D)[:, None] if X is None else X
    Sf, Sy = kernF.K(X), kernY.K(X)
    z = rng.randint(0, K, N)
    means = (np.linalg.cholesky(Sf + 1e-8 * np.eye(D)).dot(rng.randn(D, K))).T
    Y = means[z] + (np.linalg.cholesky(Sy + 1e-8 * np.eye(D)).dot(rng.randn(D, N))).T
    np.random.seed(seed + 1)
    m = GPclust.MOHGP(X, kernF, kernY, Y, K=K, alpha=alpha, prior_Z=prior_Z)
    out = {"X": X, "Y": Y, "phi0": m.phi_.copy(), "K": K, "alpha": alpha, "prior_Z": prior_Z, "method": method,
           "kernF": json.dumps(kernF.spec()), "kernY": json.dumps(kernY.spec())}
    g, ng = m.vb_grad_natgrad()
    out.update(bound0=m.bound(), grad0=g, natgrad0=ng, phi_init=m.phi.copy(), phi_hat0=m.phi_hat.copy(), H0=m.H, muk0=m.muk.copy(),
               Lambda_inv0=m.Lambda_inv.copy(), log_det_diff0=m.log_det_diff.copy(), ybark0=m.ybark.copy(),
               Sy_chol=m.Sy_chol.copy(), Sy_logdet=m.Sy_logdet, mix0=m.mixing_prop_bound(), mixgrad0=m.mixing_prop_bound_grad().copy())

    def kgrads():
        # the reference's own update_kern_grads (MOHGP.py:92-122): d bound / d (kernel parameter), kernF's then kernY's
        m.update_kern_grads()
        return np.array([getattr(o, n + "_gradient") for o, n in m.kernF.param_list() + m.kernY.param_list()])

    out.update(kgrad0=kgrads())
    out["trace"] = run_optimize(m, method=method, maxiter=maxiter)
    g, ng = m.vb_grad_natgrad()
    out.update(bound1=m.bound(), grad1=g, natgrad1=ng, phi1=m.phi.copy(), phi_1=m.phi_.copy(), muk1=m.muk.copy(), kgrad1=kgrads())
    # the reference's own MOHGP.predict_components (MOHGP.py:156-174) at the optimised point, inside and outside the design range
    Xnew = np.linspace(X.min() - 1.0, X.max() + 1.0, 9)[:, None]
    mu, var = m.predict_components(Xnew)
    out.update(pred_Xnew=Xnew, pred_mu=np.array(mu), pred_var=np.array(var))
    np.savez_compressed(os.path.join(HERE, "reference_mohgp_%s.npz" % name), **out)
    print("mohgp", name, "bound0 %.12g -> bound1 %.12g, %d trace rows" % (out["bound0"], out["bound1"], len(out["trace"])))


def omgp_case(GPclust, name, N, Dy, K, prior_Z, alpha, seed, variance, kernels, maxiter=12):
    rng = np.random.RandomState(seed)
    X = np.sort(rng.uniform(0, 10, N))[:, None]
    comp = rng.randint(0, K, N)
    Y = np.zeros((N, Dy))
    for d in range(Dy):
        for i in range(K):
            sel = comp == i
            Y[sel, d] = np.sin(X[sel, 0] * (0.6 + 0.5 * i) + d + i) * (1.0 + 0.3 * i) + 0.1 * rng.randn(sel.sum())
    np.random.seed(seed + 1)
    m = GPclust.OMGP(X, Y, K=K, kernels=kernels, variance=variance, alpha=alpha, prior_Z=prior_Z)
    out = {"X": X, "Y": Y, "phi0": m.phi_.copy(), "K": K, "alpha": alpha, "prior_Z": prior_Z, "variance": variance,
           "kernels": json.dumps([k.spec() for k in m.kern])}
    g, ng = m.vb_grad_natgrad()
    out.update(bound0=m.bound(), grad0=g, natgrad0=ng, phi_init=m.phi.copy())

    def kgrads():
        # the reference's own update_kern_grads (OMGP.py:55-94): [d/d variance, then (variance, lengthscale) per component]
        m.update_kern_grads()
        v = [float(np.asarray(m.variance.gradient).reshape(-1)[0])]
        for k in m.kern:
            v.extend([k.variance_gradient, k.lengthscale_gradient])
        return np.array(v)

    out.update(kgrad0=kgrads())
    out["trace"] = run_optimize(m, method="HS", maxiter=maxiter)
    g, ng = m.vb_grad_natgrad()
    out.update(bound1=m.bound(), grad1=g, natgrad1=ng, phi1=m.phi.copy(), phi_1=m.phi_.copy(), kgrad1=kgrads())
    Xnew = np.linspace(0, 10, 7)[:, None]
    mu, va = m.predict(Xnew, 0)
    out.update(Xnew=Xnew, pred_mu0=mu, pred_va0=va)
    np.savez_compressed(os.path.join(HERE, "reference_omgp_%s.npz" % name), **out)
    print("omgp", name, "bound0 %.12g -> bound1 %.12g, %d trace rows" % (out["bound0"], out["bound1"], len(out["trace"])))


def mog_case(GPclust, name, N, D, K, prior_Z, alpha, seed, maxiter=30):
    rng = np.random.RandomState(seed)
    centres = rng.randn(K, D) * 4.0
    X = centres[rng.randint(0, K, N)] + rng.randn(N, D)
    np.random.seed(seed + 1)
    m = GPclust.MOG(X, K=K, prior_Z=prior_Z, alpha=alpha)
    out = {"X": X, "phi0": m.phi_.copy(), "K": K, "alpha": alpha, "prior_Z": prior_Z}
    g, ng = m.vb_grad_natgrad()
    out.update(bound0=m.bound(), grad0=g, natgrad0=ng, mun0=m.mun.copy(), Sns0=m.Sns.copy())
    out["trace"] = run_optimize(m, method="HS", maxiter=maxiter)
    g, ng = m.vb_grad_natgrad()
    out.update(bound1=m.bound(), grad1=g, natgrad1=ng, phi1=m.phi.copy(), phi_1=m.phi_.copy())
    np.savez_compressed(os.path.join(HERE, "reference_mog_%s.npz" % name), **out)
    print("mog", name, "bound0 %.12g -> bound1 %.12g, %d trace rows" % (out["bound0"], out["bound1"], len(out["trace"])))


def splits_case(GPclust, name, N, D, K0, K_true, prior_Z, alpha, seed, kernF, kernY):
    """The reference's own merge-split search on MOHGP (collapsed_mixture.py:96-205): optimize with the steepest method (VBEM),
    then systematic_splits() from K0 < K_true clusters.  The stdout of the whole search is kept: which splits were attempted,
    which succeeded and by how much the bound changed.  (Used by the CPU oracle test only -- the file name keeps it out of the
    `reference_mohgp_*` parametrisations.)"""
    rng = np.random.RandomState(seed)
    X = np.linspace(0, 10, D)[:, None]
    Sf, Sy = kernF.K(X), kernY.K(X)
    z = rng.randint(0, K_true, N)
    means = (np.linalg.cholesky(Sf + 1e-8 * np.eye(D)).dot(rng.randn(D, K_true))).T * 2.0
    Y = means[z] + (np.linalg.cholesky(Sy + 1e-8 * np.eye(D)).dot(rng.randn(D, N))).T
    np.random.seed(seed + 1)
    m = GPclust.MOHGP(X, kernF, kernY, Y, K=K0, alpha=alpha, prior_Z=prior_Z)
    out = {"X": X, "Y": Y, "phi0": m.phi_.copy(), "K": K0, "alpha": alpha, "prior_Z": prior_Z,
           "kernF": json.dumps(kernF.spec()), "kernY": json.dumps(kernY.spec()), "bound0": m.bound()}
    out["trace_steepest"] = run_optimize(m, method="steepest", maxiter=40)
    out.update(bound1=m.bound(), phi1=m.phi.copy())
    np.random.seed(seed + 2)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        m.systematic_splits(verbose=True)
    text = buf.getvalue().replace("\r", "\n")
    events = re.findall(r"attempting to split cluster\s+(\d+)|split (failed|suceeded), bound changed by:\s+(\S+)", text)
    attempts = [int(a) for a, _, _ in events if a]
    results = [(1.0 if r == "suceeded" else 0.0, float(v)) for a, r, v in events if r]
    out.update(split_attempts=np.array(attempts, dtype=np.int64), split_results=np.array(results, dtype=np.float64).reshape(-1, 2),
               K2=m.K, bound2=m.bound(), phi2=m.phi.copy(), phi_hat2=m.phi_hat.copy(),
               rng_after=np.random.randn(3))  # the legacy global stream must have been consumed identically
    np.savez_compressed(os.path.join(HERE, "reference_splits_%s.npz" % name), **out)
    print("splits", name, "K %d -> %d, bound %.12g -> %.12g -> %.12g, %d attempts, %d succeeded" % (
        K0, m.K, out["bound0"], out["bound1"], out["bound2"], len(attempts), int(sum(r[0] for r in results))))


def config1_case(GPclust):
    """BASELINE.json configs[0] stand-in (data/kalinka09.csv itself is not in the reference tree): the replicate /
    time design of the melanogaster rows of data/kalinka09_pdata.csv (56 samples, 8 replicates per hour), Matern52 kernels and DP prior
    as in drosophila.ipynb, N = 500 series drawn from the model and row-normalised, K = 20 -- the same problem as
    tests/test_gpu_mohgp.py::test_matern_dp_drosophila_like, run here through the reference's own MOHGP / CollapsedVB."""
    rows = [l.strip().split(",") for l in open(os.path.join(HERE, "kalinka09_pdata.csv")).read().strip().split("\n")]
    hours = np.array([float(r[2]) for r in rows if r[0] == "me"])
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
    m = GPclust.MOHGP(X, kF, kY, Y, K=K, prior_Z="DP", alpha=1.0)
    out = {"X": X, "Y": Y, "phi0": m.phi_.copy(), "K": K, "alpha": 1.0, "prior_Z": "DP", "method": "HS",
           "kernF": json.dumps(kF.spec()), "kernY": json.dumps(kY.spec())}
    g, ng = m.vb_grad_natgrad()
    out.update(bound0=m.bound(), natgrad0=ng, phi_hat0=m.phi_hat.copy(), muk0=m.muk.copy(), log_det_diff0=m.log_det_diff.copy())
    out["trace"] = run_optimize(m, method="HS", maxiter=30)
    out.update(bound1=m.bound(), phi1=m.phi.copy(), muk1=m.muk.copy())
    np.savez_compressed(os.path.join(HERE, "reference_config1_kalinka.npz"), **out)
    print("config1 kalinka stand-in: bound0 %.12g -> bound1 %.12g, %d trace rows" % (out["bound0"], out["bound1"], len(out["trace"])))


def omgp_demo_case(GPclust):
    """The reference's own demo data for OMGP (data/split_data_test.csv, OMGP_demo.ipynb cells [2]-[12]: N = 250 points of two
    overlapping curves, K = 2, variance 0.01, DP prior) -- the notebook's stored outputs are unseeded, so the run is repeated here with a
    seed through the reference's own OMGP / CollapsedVB.  CPU oracle pin (the file name keeps it out of the `reference_omgp_*` cases)."""
    raw = np.loadtxt(os.path.join(HERE, "split_data_test.csv"), delimiter=",", skiprows=1)
    X, Y = raw[:, 1:2], raw[:, 2:3]
    np.random.seed(11)
    m = GPclust.OMGP(X, Y, K=2, variance=0.01, prior_Z="DP")
    out = {"X": X, "Y": Y, "phi0": m.phi_.copy(), "K": 2, "alpha": 1.0, "prior_Z": "DP", "variance": 0.01,
           "kernels": json.dumps([k.spec() for k in m.kern])}
    g, ng = m.vb_grad_natgrad()
    out.update(bound0=m.bound(), grad0=g, natgrad0=ng, log_likelihood0=m.log_likelihood())
    out["trace"] = run_optimize(m, method="HS", maxiter=15)
    g, ng = m.vb_grad_natgrad()
    out.update(bound1=m.bound(), grad1=g, natgrad1=ng, phi1=m.phi.copy())
    Xnew = np.linspace(X.min(), X.max(), 11)[:, None]
    mu, va = m.predict(Xnew, 1)
    out.update(Xnew=Xnew, pred_mu1=mu, pred_va1=va)
    np.savez_compressed(os.path.join(HERE, "reference_demo_omgp.npz"), **out)
    print("omgp demo data: bound0 %.12g -> bound1 %.12g, %d trace rows" % (out["bound0"], out["bound1"], len(out["trace"])))


def quirks_case(GPclust):
    """Edge cases of SURVEY.md appendix A through the reference's own code (CPU oracle pin):
      A.3   a cluster nobody belongs to (phi_hat <= 1e-6): Lambda_inv_k := Sf, the rest computed normally; remove_empty_clusters()
      A.9   reorder() sorts the clusters under the DP prior and is a no-op under the symmetric one
      A.6   maxiter may be a float; the iteration count at the exit
      A.10  randomize() re-draws phi_ from the legacy global stream
      A.11  OMGP.do_computations grows the kernel list by ONE per call: with one kernel given and K = 3 the first bound() sums over
            two components only, the next set_vb_param() adds the third"""
    out = {}
    rng = np.random.RandomState(21)
    D, N, K = 8, 50, 4
    X = np.linspace(0, 10, D)[:, None]
    kF, kY = G.RBF(1, 1.0, 2.0), G.RBF(1, 0.1, 1.0) + G.White(1, 0.05)
    Y = (np.linalg.cholesky(kF.K(X) + 1e-8 * np.eye(D)).dot(rng.randn(D, 3))).T[rng.randint(0, 3, N)] + 0.3 * rng.randn(N, D)
    np.random.seed(22)
    m = GPclust.MOHGP(X, kF, kY, Y, K=K, prior_Z="DP", alpha=2.0)
    phi_ = m.phi_.copy()
    phi_[:, 2] = -60.0                                   # nobody belongs to cluster 2: phi_hat[2] ~ 1e-25
    m.set_vb_param(phi_.flatten())
    g, ng = m.vb_grad_natgrad()
    out.update(e_X=X, e_Y=Y, e_phi0=phi_, e_phi_hat=m.phi_hat.copy(), e_bound=m.bound(), e_grad=g, e_natgrad=ng, e_Lambda_inv=m.Lambda_inv.copy(),
               e_muk=m.muk.copy(), e_log_det_diff=m.log_det_diff.copy(), e_Sf=m.Sf.copy())
    m.reorder()                                          # DP: biggest first, the empty one last
    out.update(r_phi_hat=m.phi_hat.copy(), r_bound=m.bound(), r_phi_=m.phi_.copy())
    m.remove_empty_clusters()
    out.update(x_K=m.K, x_bound=m.bound(), x_phi_hat=m.phi_hat.copy())
    out["f_trace"] = run_optimize(m, method="HS", maxiter=3.5)
    np.random.seed(23)
    m.randomize()
    out.update(z_phi_=m.phi_.copy(), z_bound=m.bound())
    np.random.seed(22)
    ms = GPclust.MOHGP(X, kF, kY, Y, K=K, prior_Z="symmetric", alpha=2.0)
    before = ms.phi_.copy()
    ms.reorder()
    out.update(s_noop=np.array_equal(before, ms.phi_), s_bound=ms.bound())

    Xo = np.sort(rng.uniform(0, 10, 30))[:, None]
    Yo = np.sin(Xo) + 0.1 * rng.randn(30, 1)
    np.random.seed(24)
    mo = GPclust.OMGP(Xo, Yo, K=3, kernels=[G.RBF(1, 1.0, 1.5)], variance=0.05)
    out.update(o_X=Xo, o_Y=Yo, o_phi0=mo.phi_.copy(), o_nkern0=len(mo.kern), o_bound0=mo.bound())
    mo.set_vb_param(mo.get_vb_param())
    out.update(o_nkern1=len(mo.kern), o_bound1=mo.bound())
    mo.set_vb_param(mo.get_vb_param())
    g, ng = mo.vb_grad_natgrad()
    out.update(o_nkern2=len(mo.kern), o_bound2=mo.bound(), o_natgrad2=ng)
    np.savez_compressed(os.path.join(HERE, "reference_quirks.npz"), **out)
    print("quirks: empty-cluster bound %.12g, reordered %.12g, K after removal %d, float maxiter exit at iteration %d, OMGP kernels %d -> %d -> %d" % (
        out["e_bound"], out["r_bound"], out["x_K"], int(out["f_trace"][-1, 0]), out["o_nkern0"], out["o_nkern1"], out["o_nkern2"]))


def main():
    install_gpy_stub()
    sys.path.insert(0, REF)
    import GPclust
    assert os.path.realpath(GPclust.__file__).startswith(os.path.realpath(REF)), GPclust.__file__

    mohgp_case(GPclust, "sym", N=60, D=12, K=4, prior_Z="symmetric", alpha=1.0, seed=0,
               kernF=G.RBF(1, 1.0, 2.0), kernY=G.RBF(1, 0.1, 1.0) + G.White(1, 0.05))
    mohgp_case(GPclust, "dp", N=45, D=9, K=5, prior_Z="DP", alpha=3.0, seed=1,
               kernF=G.Matern52(1, 1.0, 3.0), kernY=G.Matern32(1, 0.2, 1.5) + G.White(1, 0.05) + G.Bias(1, 0.1), method="PR")
    # replicated time points (drosophila design: 8 replicates per hour, rows of data/kalinka09_pdata.csv) -> singular Sf
    hours = np.repeat(np.arange(2.0, 9.0), 2)[:, None]
    mohgp_case(GPclust, "replicates", N=40, D=14, K=3, prior_Z="DP", alpha=1.0, seed=2,
               kernF=G.Matern52(1, 1.0, 12.0), kernY=G.Matern52(1, 0.1, 12.0) + G.White(1, 0.05), X=hours, method="FR", maxiter=15)
    omgp_case(GPclust, "k2", N=36, Dy=1, K=2, prior_Z="symmetric", alpha=1.0, seed=3, variance=0.01,
              kernels=[G.RBF(1, 1.0, 1.0), G.RBF(1, 1.0, 2.0)])
    omgp_case(GPclust, "k3", N=30, Dy=2, K=3, prior_Z="DP", alpha=2.0, seed=4, variance=0.05, kernels=None)
    mog_case(GPclust, "d3", N=80, D=3, K=4, prior_Z="symmetric", alpha=1.0, seed=5)
    mog_case(GPclust, "dp", N=70, D=2, K=5, prior_Z="DP", alpha=4.0, seed=6)
    config1_case(GPclust)
    quirks_case(GPclust)
    omgp_demo_case(GPclust)
    splits_case(GPclust, "mohgp", N=90, D=10, K0=2, K_true=4, prior_Z="symmetric", alpha=1.0, seed=7,
                kernF=G.RBF(1, 1.0, 2.0), kernY=G.RBF(1, 0.1, 1.0) + G.White(1, 0.05))
    splits_case(GPclust, "mohgp_dp", N=70, D=8, K0=3, K_true=5, prior_Z="DP", alpha=2.0, seed=8,
                kernF=G.Matern52(1, 1.0, 3.0), kernY=G.Matern32(1, 0.1, 2.0) + G.White(1, 0.05))


if __name__ == "__main__":
    main()
