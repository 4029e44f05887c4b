"""Memory-safety and race evidence without compute-sanitizer (closed on this pool: profiles/sanitizer_r02_closed.txt).

  * guard zones: every device array of a model is carved out of a larger allocation with 4 KB canary margins; after the kernels have
    run (fused loop kernel, launch sequence, wide mode, ragged N and K, MOG) no canary word may have changed -- an out-of-bounds store
    of any kernel into a neighbouring margin fails the test;
  * determinism: the fused loop kernel (grid barriers, ring re-carving, ticket / scalar CTA, in-kernel reductions) is run twice over
    150 evaluations and once as the launch sequence: the three traces must be IDENTICAL bit for bit -- a data race in any of the
    hand-rolled protocols shows up as a run-to-run difference."""
import numpy as np
import pytest

from helpers import mohgp_problem, product_kernels

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,D,K,fused", [(1003, 64, 64, "1"), (517, 24, 7, "1"), (517, 24, 7, "0"), (64, 8, 2, "1"), (901, 33, 130, "1"), (5, 3, 2, "1"),
                                         (4097, 56, 20, "1")])
def test_guard_zones_intact(N, D, K, fused, monkeypatch):
    from gpclust_b200 import MOHGP, _cabi
    monkeypatch.setenv("GPCLUST_B200_FUSED", fused)
    monkeypatch.setattr(_cabi, "GUARD", True)
    _cabi.check_guards(clear=True)
    X, Y, _ = mohgp_problem(N, D, K, seed=N)
    np.random.seed(1)
    kF, kY = product_kernels()
    m = MOHGP(X, kF, kY, Y, K=K)
    m.optimize(maxiter=12, verbose=False)
    m.vb_grad_natgrad()
    _ = m.phi, m.Lambda_inv, m.muk
    m.optimize(maxiter=5, verbose=False, callback=lambda: None)  # host-driven loop: the stand-alone kernels
    import torch
    torch.cuda.synchronize()
    bad = _cabi.check_guards(clear=True)
    assert not bad, bad


def test_guard_zones_intact_mog(monkeypatch):
    from gpclust_b200 import MOG, _cabi
    monkeypatch.setattr(_cabi, "GUARD", True)
    _cabi.check_guards(clear=True)
    rng = np.random.RandomState(0)
    X = rng.randn(777, 3) + rng.randint(0, 3, (777, 1)) * 3.0
    np.random.seed(2)
    m = MOG(X, K=6)
    m.optimize(maxiter=15, verbose=False)
    import torch
    torch.cuda.synchronize()
    bad = _cabi.check_guards(clear=True)
    assert not bad, bad


@pytest.mark.parametrize("N,D,K,prior", [(3000, 64, 64, "symmetric"), (1201, 24, 9, "DP"), (40000, 32, 33, "symmetric")])
def test_fused_loop_deterministic_and_equal_to_launch_sequence(N, D, K, prior, monkeypatch):
    from gpclust_b200 import MOHGP
    X, Y, _ = mohgp_problem(N, D, K, seed=7)
    np.random.seed(8)
    phi0 = np.random.randn(N, K)
    traces, finals = [], []
    for fused in ("1", "1", "0"):
        monkeypatch.setenv("GPCLUST_B200_FUSED", fused)
        kF, kY = product_kernels()
        m = MOHGP(X, kF, kY, Y, K=K, prior_Z=prior, alpha=1.0 if prior == "symmetric" else 2.0, phi_init=phi0)
        assert m._fused_ok() == (fused == "1")
        m.optimize(maxiter=150, ftol=-1.0, gtol=-1.0, verbose=False)
        traces.append(np.array(m.optimize_trace))
        finals.append(m.phi_.copy())
    assert traces[0].shape == traces[1].shape == traces[2].shape
    assert np.array_equal(traces[0], traces[1]) and np.array_equal(finals[0], finals[1])  # run to run
    assert np.array_equal(traces[0], traces[2]) and np.array_equal(finals[0], finals[2])  # fused kernel vs launch sequence
