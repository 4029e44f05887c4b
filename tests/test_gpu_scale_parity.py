"""Oracle parity at the shape the metric is quoted on (BASELINE.json config 3: N = 1M series, D = 64, K = 64, FP64, HS) and at
N = 200k: bound, natural gradient, five loop-body passes (accept / reject decisions, bounds), responsibilities and assignments
against MOHGPOracle evaluated on the same inputs (/root/reference/GPclust/MOHGP.py:124-153, collapsed_vb.py:147-193).
phi_hat is ~15 600 per cluster here -- the regime where the reference's own Lambda_inv = (Sy - t^T t) / phi_hat (MOHGP.py:85)
cancels hardest.  Tolerance: 1e-9 relative (north_star)."""
import numpy as np
import pytest

from helpers import headline_parity, product_kernels

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.mark.parametrize("N", [200000, 1000000])
def test_oracle_parity_at_headline_shape(N):
    from bench import synth
    from gpclust_b200 import MOHGP
    D, K = 64, 64
    X, Y, phi0 = synth(N, D, K)
    kF, kY = product_kernels()
    m = MOHGP(X, kF, kY, Y, K=K, phi_init=phi0)
    r = headline_parity(X, Y, phi0, K, m, passes=5, method="HS")
    assert r["bound_rel"] < TOL, r
    assert r["natgrad_rel"] < TOL and r["grad_rel"] < TOL, r
    assert r["trace_iterations_equal"], r
    assert r["bound_rel_after"] < TOL, r
    assert r["phi_rel_after"] < TOL, r
    assert r["argmax_equal"], r
