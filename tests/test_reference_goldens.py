"""Pins against outputs of the REFERENCE'S OWN modules (tests/golden/reference_*.npz, produced by
tests/golden/make_reference_goldens.py, which imports /root/reference/GPclust unmodified with a stub GPy).

  * CPU (`-m "not gpu"`): the oracle restatement reproduces the reference's bound, gradients, responsibilities and
    optimisation trace on the same inputs.
  * GPU (`-m gpu`): the product path (Python mirror -> C-ABI -> sm_100a kernels) does, within 1e-9 relative
    (north_star), max-norm-relative for arrays.
"""
import glob
import io
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, relerr
from oracle import gpy_shim as G
from oracle.vb_oracle import MOGOracle, MOHGPOracle, OMGPOracle, mahalanobis_pairs_solve

TOL = 1e-9


def _load(path):
    d = np.load(path)
    return {k: d[k] for k in d.files}


def _cases(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, "reference_%s_*.npz" % prefix)))


def _kern(spec, mod):
    """spec: [(kind, variance, lengthscale), ...] -> a kernel of `mod` (oracle.gpy_shim or gpclust_b200.kern)."""
    ctor = {"rbf": mod.RBF, "matern32": mod.Matern32, "matern52": mod.Matern52}
    parts = []
    for kind, var, ls in spec:
        if kind in ctor:
            parts.append(ctor[kind](1, var, ls))
        elif kind == "bias":
            parts.append(mod.Bias(1, var))
        else:
            parts.append(mod.White(1, var))
    out = parts[0]
    for p in parts[1:]:
        out = out + p
    if hasattr(out, "fix"):
        out.fix()  # the fixtures were produced with the kernel hyper-parameters held fixed
    return out


def _scalar(v):
    return v.item() if hasattr(v, "item") else v


def _check_trace(trace, ref):
    trace = np.asarray(trace, dtype=np.float64).reshape(-1, 4)
    assert trace.shape == ref.shape, (trace.shape, ref.shape)
    assert np.array_equal(trace[:, 0], ref[:, 0])  # identical iteration numbering (accept / reject decisions)
    # the reference prints str(float) -> full precision; bounds to 1e-9 relative
    assert np.max(np.abs(trace[:, 1] - ref[:, 1]) / np.abs(ref[:, 1])) < TOL
    assert relerr(trace[:, 2], ref[:, 2]) < 1e-7  # squared gradient norms
    assert np.allclose(trace[:, 3], ref[:, 3], rtol=1e-6, atol=1e-9)  # beta: ratio of small differences


def _check_model(m, d, trace, after=True):
    if after:
        assert abs(m.bound() - _scalar(d["bound1"])) < TOL * abs(_scalar(d["bound1"]))
        assert relerr(m.phi, d["phi1"]) < 1e-8
        assert np.array_equal(np.argmax(m.phi, 1), np.argmax(d["phi1"], 1))
        # at convergence the gradient is rounding noise (1e-14): measure the error against the scale of the start gradient
        g, ng = m.vb_grad_natgrad()
        for got, ref, start in ((ng, d["natgrad1"], d["natgrad0"]), (g, d["grad1"], d["grad0"])):
            scale = max(np.max(np.abs(ref)), 1e-4 * np.max(np.abs(start)))
            assert np.max(np.abs(got - ref)) < 1e-7 * scale
    _check_trace(trace, d["trace"])


def _check_start(m, d):
    assert np.array_equal(m.phi_, d["phi0"])
    assert abs(m.bound() - _scalar(d["bound0"])) < TOL * abs(_scalar(d["bound0"]))
    g, ng = m.vb_grad_natgrad()
    assert relerr(ng, d["natgrad0"]) < TOL and relerr(g, d["grad0"]) < TOL


# ------------------------------------------------------------------------------------------------ oracle (CPU)
@pytest.mark.parametrize("path", _cases("mohgp"))
def test_oracle_mohgp_matches_reference(path):
    d = _load(path)
    seed_state = np.random.get_state()
    m = MOHGPOracle(d["X"], _kern(json.loads(str(d["kernF"])), G), _kern(json.loads(str(d["kernY"])), G), d["Y"], K=int(d["K"]),
                    alpha=float(d["alpha"]), prior_Z=str(d["prior_Z"]), mahalanobis=mahalanobis_pairs_solve)
    np.random.set_state(seed_state)
    m.set_vb_param(d["phi0"].flatten())
    _check_start(m, d)
    assert relerr(m.muk, d["muk0"]) < TOL and relerr(m.log_det_diff, d["log_det_diff0"]) < TOL
    assert relerr(m.Lambda_inv, d["Lambda_inv0"]) < 1e-9 and relerr(m.ybark, d["ybark0"]) < 1e-13
    assert relerr(m.phi, d["phi_init"]) < 1e-14 and abs(m.H - _scalar(d["H0"])) < 1e-12 * abs(_scalar(d["H0"]))
    assert abs(m.mixing_prop_bound() - _scalar(d["mix0"])) < 1e-12 * max(1.0, abs(_scalar(d["mix0"])))
    assert relerr(m.mixing_prop_bound_grad(), d["mixgrad0"]) < 1e-14
    assert relerr(m.kernel_gradients(), d["kgrad0"]) < TOL  # MOHGP.update_kern_grads (MOHGP.py:92-122) at the initial point
    m.out = io.StringIO()
    m.optimize(method=str(d["method"]), maxiter=int(d["trace"][-1, 0]) if len(d["trace"]) else 25, verbose=False)
    _check_model(m, d, m.trace)
    _check_predict(m, d)
    assert relerr(m.kernel_gradients(), d["kgrad1"]) < 1e-7  # at the optimised point (rounding of the loop amplified)


def _check_predict(m, d, tol=1e-7):
    """MOHGP.predict_components (MOHGP.py:156-174) at the optimised point against the reference's own output: K mean vectors and
    K covariance matrices at 9 new inputs (inside and outside the design range).  The tolerance is that of the optimised phi
    (rounding of ~10-25 loop passes), measured against the largest entry."""
    mu, var = m.predict_components(d["pred_Xnew"])
    assert len(mu) == len(var) == int(d["K"])
    assert relerr(np.array(mu), d["pred_mu"]) < tol
    assert relerr(np.array(var), d["pred_var"]) < tol


@pytest.mark.parametrize("path", _cases("omgp"))
def test_oracle_omgp_matches_reference(path):
    d = _load(path)
    kernels = [_kern(s, G) for s in json.loads(str(d["kernels"]))]
    m = OMGPOracle(d["X"], d["Y"], K=int(d["K"]), kernels=kernels, variance=float(d["variance"]), alpha=float(d["alpha"]),
                   prior_Z=str(d["prior_Z"]))
    m.set_vb_param(d["phi0"].flatten())
    _check_start(m, d)
    assert relerr(m.kernel_gradients(), d["kgrad0"]) < TOL  # OMGP.update_kern_grads (OMGP.py:55-94) at the initial point
    m.out = io.StringIO()
    m.optimize(method="HS", maxiter=12, verbose=False)
    _check_model(m, d, m.trace)
    assert relerr(m.kernel_gradients(), d["kgrad1"]) < 1e-7  # after 8 accepted steps (rounding of the loop amplified)


@pytest.mark.parametrize("path", _cases("mog"))
def test_oracle_mog_matches_reference(path):
    d = _load(path)
    m = MOGOracle(d["X"], K=int(d["K"]), prior_Z=str(d["prior_Z"]), alpha=float(d["alpha"]))
    m.set_vb_param(d["phi0"].flatten())
    _check_start(m, d)
    assert relerr(m.mun, d["mun0"]) < 1e-12 and relerr(m.Sns, d["Sns0"]) < 1e-12
    m.out = io.StringIO()
    m.optimize(method="HS", maxiter=30, verbose=False)
    _check_model(m, d, m.trace)


def test_oracle_quirks_match_reference():
    """SURVEY.md appendix A edge cases through the reference's own code (tests/golden/reference_quirks.npz): an empty cluster
    (phi_hat <= 1e-6 -> Lambda_inv_k := Sf), reorder() under DP / symmetric priors, remove_empty_clusters(), a float maxiter,
    randomize(), and OMGP's kernel list growing by one per do_computations()."""
    d = _load(os.path.join(GOLDEN, "reference_quirks.npz"))
    kF, kY = G.RBF(1, 1.0, 2.0), G.RBF(1, 0.1, 1.0) + G.White(1, 0.05)
    np.random.seed(22)
    m = MOHGPOracle(d["e_X"], kF, kY, d["e_Y"], K=4, prior_Z="DP", alpha=2.0, mahalanobis=mahalanobis_pairs_solve)
    m.set_vb_param(d["e_phi0"].flatten())
    assert m.phi_hat[2] < 1e-20 and relerr(m.phi_hat, d["e_phi_hat"]) < 1e-13
    assert abs(m.bound() - _scalar(d["e_bound"])) < TOL * abs(_scalar(d["e_bound"]))
    assert np.array_equal(m.Lambda_inv[2], d["e_Sf"]) and relerr(m.Lambda_inv, d["e_Lambda_inv"]) < TOL        # A.3
    assert relerr(m.muk, d["e_muk"]) < TOL and relerr(m.log_det_diff, d["e_log_det_diff"]) < TOL
    g, ng = m.vb_grad_natgrad()
    assert relerr(ng, d["e_natgrad"]) < TOL and relerr(g, d["e_grad"]) < TOL
    m.reorder()                                                                                               # A.9 (DP)
    assert np.all(np.diff(m.phi_hat) <= 0) and np.array_equal(m.phi_, d["r_phi_"])
    assert relerr(m.phi_hat, d["r_phi_hat"]) < 1e-13 and abs(m.bound() - _scalar(d["r_bound"])) < TOL * abs(_scalar(d["r_bound"]))
    m.remove_empty_clusters()
    assert m.K == int(d["x_K"]) == 3 and relerr(m.phi_hat, d["x_phi_hat"]) < 1e-13
    assert abs(m.bound() - _scalar(d["x_bound"])) < TOL * abs(_scalar(d["x_bound"]))
    m.out = io.StringIO()
    m.optimize(method="HS", maxiter=3.5, verbose=False)                                                       # A.6: float maxiter
    _check_trace(m.trace, d["f_trace"])
    np.random.seed(23)
    m.randomize()                                                                                             # A.10
    assert np.array_equal(m.phi_, d["z_phi_"]) and abs(m.bound() - _scalar(d["z_bound"])) < TOL * abs(_scalar(d["z_bound"]))
    np.random.seed(22)
    ms = MOHGPOracle(d["e_X"], kF, kY, d["e_Y"], K=4, prior_Z="symmetric", alpha=2.0, mahalanobis=mahalanobis_pairs_solve)
    before = ms.phi_.copy()
    ms.reorder()                                                                                              # A.9 (symmetric: no-op)
    assert bool(d["s_noop"]) and np.array_equal(before, ms.phi_) and abs(ms.bound() - _scalar(d["s_bound"])) < TOL * abs(_scalar(d["s_bound"]))

    np.random.seed(24)
    mo = OMGPOracle(d["o_X"], d["o_Y"], K=3, kernels=[G.RBF(1, 1.0, 1.5)], variance=0.05)                     # A.11
    assert np.array_equal(mo.phi_, d["o_phi0"])
    for tag in ("0", "1", "2"):
        assert len(mo.kern) == int(d["o_nkern" + tag])
        assert abs(mo.bound() - _scalar(d["o_bound" + tag])) < TOL * abs(_scalar(d["o_bound" + tag]))
        if tag != "2":
            mo.set_vb_param(mo.get_vb_param())
    assert relerr(mo.vb_grad_natgrad()[1], d["o_natgrad2"]) < TOL


def test_oracle_omgp_demo_data_matches_reference():
    """The reference's own OMGP demo data (data/split_data_test.csv, N = 250, K = 2, variance 0.01, DP prior; OMGP_demo.ipynb) through
    the reference's own code with a seed: start bound (the notebook's unseeded cell [6] prints 921.947; a seeded start gives 921.42),
    gradients, trace, final responsibilities and predict()."""
    d = _load(os.path.join(GOLDEN, "reference_demo_omgp.npz"))
    kernels = [_kern(s, G) for s in json.loads(str(d["kernels"]))]
    m = OMGPOracle(d["X"], d["Y"], K=int(d["K"]), kernels=kernels, variance=float(d["variance"]), alpha=float(d["alpha"]),
                   prior_Z=str(d["prior_Z"]))
    m.set_vb_param(d["phi0"].flatten())
    _check_start(m, d)
    assert 900.0 < m.bound() < 940.0 and abs(m.log_likelihood() - _scalar(d["log_likelihood0"])) < TOL * abs(m.bound())
    m.out = io.StringIO()
    m.optimize(method="HS", maxiter=15, verbose=False)
    _check_model(m, d, m.trace)


def test_oracle_config1_standin_matches_reference():
    """BASELINE.json configs[0] stand-in (kalinka replicate / time design, D = 56, Matern52, DP prior, N = 500, K = 20) through the
    reference's own MOHGP / CollapsedVB.optimize: the oracle reproduces the start bound, natural gradient, posterior means, the
    28-line optimisation trace and the final assignments.  (The GPU path runs the same problem against the oracle in
    tests/test_gpu_mohgp.py::test_matern_dp_drosophila_like.)"""
    d = _load(os.path.join(GOLDEN, "reference_config1_kalinka.npz"))
    m = MOHGPOracle(d["X"], _kern(json.loads(str(d["kernF"])), G), _kern(json.loads(str(d["kernY"])), G), d["Y"], K=int(d["K"]),
                    alpha=float(d["alpha"]), prior_Z=str(d["prior_Z"]))
    m.set_vb_param(d["phi0"].flatten())
    assert abs(m.bound() - _scalar(d["bound0"])) < TOL * abs(_scalar(d["bound0"]))
    assert relerr(m.vb_grad_natgrad()[1], d["natgrad0"]) < TOL
    assert relerr(m.muk, d["muk0"]) < TOL and relerr(m.log_det_diff, d["log_det_diff0"]) < TOL and relerr(m.phi_hat, d["phi_hat0"]) < 1e-13
    m.out = io.StringIO()
    m.optimize(method="HS", maxiter=30, verbose=False)
    _check_trace(m.trace, d["trace"])
    assert abs(m.bound() - _scalar(d["bound1"])) < TOL * abs(_scalar(d["bound1"]))
    assert relerr(m.phi, d["phi1"]) < 1e-7 and np.array_equal(np.argmax(m.phi, 1), np.argmax(d["phi1"], 1))
    assert relerr(m.muk, d["muk1"]) < 1e-7


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "reference_splits_*.npz"))))
def test_oracle_merge_split_search_matches_reference(path):
    """collapsed_mixture.py:96-205 through the reference's own code (MOHGP, steepest optimisation, then systematic_splits()):
    the oracle must attempt the same clusters in the same order, take the same accept / reject decision for every split with the
    same change of the bound, end with the same K and assignments, and leave the legacy numpy stream in the same state."""
    import re
    d = _load(path)
    m = MOHGPOracle(d["X"], _kern(json.loads(str(d["kernF"])), G), _kern(json.loads(str(d["kernY"])), G), d["Y"], K=int(d["K"]),
                    alpha=float(d["alpha"]), prior_Z=str(d["prior_Z"]), mahalanobis=mahalanobis_pairs_solve)
    m.set_vb_param(d["phi0"].flatten())
    assert abs(m.bound() - _scalar(d["bound0"])) < TOL * abs(_scalar(d["bound0"]))
    m.out = io.StringIO()
    m.optimize(method="steepest", maxiter=40, verbose=False)
    _check_trace(m.trace, d["trace_steepest"])
    assert abs(m.bound() - _scalar(d["bound1"])) < TOL * abs(_scalar(d["bound1"]))
    seed = {"reference_splits_mohgp.npz": 7, "reference_splits_mohgp_dp.npz": 8}[os.path.basename(path)]
    np.random.seed(seed + 2)
    m.out = io.StringIO()
    m.systematic_splits(verbose=True)
    text = m.out.getvalue()
    events = re.findall(r"attempting to split cluster\s+(\d+)|split (failed|suceeded), bound changed by:\s+(\S+)", text)
    attempts = [int(a) for a, _, _ in events if a]
    results = np.array([(1.0 if r == "suceeded" else 0.0, float(v)) for a, r, v in events if r]).reshape(-1, 2)
    assert attempts == list(d["split_attempts"])
    assert np.array_equal(results[:, 0], d["split_results"][:, 0])
    # bound increments of whole optimisations (up to 100 + 5000 evaluations each): rounding of the loop amplified
    assert np.allclose(results[:, 1], d["split_results"][:, 1], rtol=1e-6, atol=1e-6)
    assert m.K == int(d["K2"])
    assert abs(m.bound() - _scalar(d["bound2"])) < 1e-8 * abs(_scalar(d["bound2"]))
    assert np.array_equal(np.argmax(m.phi, 1), np.argmax(d["phi2"], 1))
    assert relerr(m.phi_hat, d["phi_hat2"]) < 1e-6
    assert np.array_equal(np.random.randn(3), d["rng_after"])


# ------------------------------------------------------------------------------------------------ product (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("path", _cases("mohgp"))
def test_gpu_mohgp_matches_reference(path):
    from gpclust_b200 import MOHGP, kern
    d = _load(path)
    m = MOHGP(d["X"], _kern(json.loads(str(d["kernF"])), kern), _kern(json.loads(str(d["kernY"])), kern), d["Y"], K=int(d["K"]),
              alpha=float(d["alpha"]), prior_Z=str(d["prior_Z"]), phi_init=d["phi0"])
    _check_start(m, d)
    assert relerr(m.muk, d["muk0"]) < TOL and relerr(m.log_det_diff, d["log_det_diff0"]) < TOL
    assert relerr(m.Lambda_inv, d["Lambda_inv0"]) < 1e-7 and relerr(m.ybark, d["ybark0"]) < 1e-12
    assert relerr(m.phi, d["phi_init"]) < 1e-13
    assert relerr(m.mixing_prop_bound_grad(), d["mixgrad0"]) < 1e-13
    m.optimize(method=str(d["method"]), maxiter=int(d["trace"][-1, 0]), verbose=False)
    _check_model(m, d, m.optimize_trace)
    assert relerr(m.muk, d["muk1"]) < 1e-8
    _check_predict(m, d)


@pytest.mark.gpu
@pytest.mark.parametrize("path", _cases("mog"))
def test_gpu_mog_matches_reference(path):
    from gpclust_b200 import MOG
    d = _load(path)
    np.random.seed(0)
    m = MOG(d["X"], K=int(d["K"]), prior_Z=str(d["prior_Z"]), alpha=float(d["alpha"]))
    m.set_vb_param(d["phi0"].flatten())
    _check_start(m, d)
    assert relerr(m.mun, d["mun0"]) < 1e-10 and relerr(m.Sns, d["Sns0"]) < 1e-10
    m.optimize(method="HS", maxiter=30, verbose=False)
    _check_model(m, d, m.optimize_trace)


@pytest.mark.gpu
@pytest.mark.parametrize("path", _cases("omgp"))
def test_gpu_omgp_matches_reference(path):
    from gpclust_b200 import OMGP, kern
    d = _load(path)
    spec = json.loads(str(d["kernels"]))
    default = all(s == [["rbf", 1.0, 1.0]] for s in spec)
    np.random.seed(0)
    m = OMGP(d["X"], d["Y"], K=int(d["K"]), kernels=None if default else [_kern(s, kern) for s in spec], variance=float(d["variance"]),
             alpha=float(d["alpha"]), prior_Z=str(d["prior_Z"]))
    m.set_vb_param(d["phi0"].flatten())
    _check_start(m, d)

    def kgrads():  # gradient vector with every parameter free; the loop itself runs with them fixed, as in the fixture
        m.variance_fixed = False
        for k in m.kern:
            k.unfix()
        g = m._kernel_gradients()
        m.variance_fixed = True
        for k in m.kern:
            k.fix()
        return g

    assert relerr(kgrads(), d["kgrad0"]) < TOL  # OMGP.update_kern_grads (OMGP.py:55-94)
    m.optimize(method="HS", maxiter=12, verbose=False)
    _check_model(m, d, m.optimize_trace)
    assert relerr(kgrads(), d["kgrad1"]) < 1e-7
    mu, va = m.predict(d["Xnew"], 0)
    assert relerr(mu, d["pred_mu0"]) < 1e-7 and relerr(va, d["pred_va0"]) < 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "reference_splits_*.npz"))))
def test_gpu_merge_split_search_matches_reference(path):
    """The reference's own MOHGP merge-split search (optimize(method='steepest') + systematic_splits(), K 2 -> 4 and 3 -> 5) replayed
    through the GPU path: same attempted clusters, same accept / reject per split, bound increments, final K, assignments and
    legacy-RNG state (tools/check_split_fixture.py; measured on B200: profiles/split_fixture_r02.jsonl, bounds to 2e-14)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from check_split_fixture import replay
    rec = replay(path)
    assert rec["ok"], rec
