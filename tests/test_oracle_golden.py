"""Pin the CPU oracle against the reference's ONLY seeded known-answer outputs: the stored stdout of
notebooks/mixture of Gaussians.ipynb (cells [8], [11], [13], [16]-[23]); fixtures extracted by
tests/golden/make_golden.py.  Iteration numbering must be identical and every printed value must agree to ~1e-10 relative."""
import io
import re

import numpy as np
import pytest

from helpers import load_golden, load_mog_data
from oracle.vb_oracle import MOGOracle


def _assert_trace(trace, gold_rows, tol_bound, tol_grad, tol_beta=None):
    """Iteration numbering (incl. the +2 jumps of rejected steps) must be identical; the printed values
    (12 significant digits in the notebook) must agree to the stated relative tolerance.  Measured here:
    cells [8]/[11] bound 4e-12, grad 3e-10, beta 3e-10; the split cell [13] (120 evaluations after a
    re-initialisation) bound 1.5e-9, grad 8e-7 -- rounding-level differences between this numpy/LAPACK
    and the 2014 one, amplified by the optimiser."""
    assert len(trace) == len(gold_rows)
    for (it, bound, grad, beta), g in zip(trace, gold_rows):
        assert it == g["iteration"]
        assert abs(bound - float(g["bound"])) <= tol_bound * abs(float(g["bound"])), (it, bound, g["bound"])
        assert abs(grad - float(g["grad"])) <= tol_grad * abs(float(g["grad"])), (it, grad, g["grad"])
        if tol_beta is not None:
            assert abs(beta - float(g["beta"])) <= tol_beta * max(abs(float(g["beta"])), 1e-300), (it, beta, g["beta"])


@pytest.fixture(scope="module")
def replay():
    """Replays notebook cells [0]-[23] once with the legacy global numpy stream (np.random.seed(0))."""
    gold = load_golden()
    X = load_mog_data()
    np.random.seed(0)
    out = {}
    m = MOGOracle(X, K=10)
    m.out = io.StringIO()
    m.optimize()
    out["cell8"] = list(m.trace)

    m = MOGOracle(X, K=4, prior_Z="DP", alpha=10.0)
    m.out = io.StringIO()
    m.optimize()
    out["cell11"] = list(m.trace)

    m.trace = []
    ok = m.try_split(1)
    out["cell13"] = (ok, list(m.trace), m.out.getvalue(), m.K, m.last_split_increase)

    m.systematic_splits()
    out["final"] = m
    return gold, out


def test_cell8_symmetric_k10(replay):
    gold, out = replay
    _assert_trace(out["cell8"], gold["8"]["trace"], 2e-11, 2e-9, 2e-9)
    assert "%.9g" % out["cell8"][0][1] == "840.246015" and "%.9g" % out["cell8"][-1][1] == "1403.95252"
    assert out["cell8"][-1][0] == 65


def test_cell11_dp_k4(replay):
    gold, out = replay
    _assert_trace(out["cell11"], gold["11"]["trace"], 2e-11, 2e-9, 2e-9)


def test_cell13_try_split(replay):
    gold, out = replay
    ok, trace, text, K, inc = out["cell13"]
    assert ok is True
    _assert_trace(trace, gold["13"]["trace"], 1e-8, 1e-5)
    assert abs(inc - 58.7935883981) < 1e-6
    assert "(K=5)" in text


def test_cells21_23_posterior_and_density(replay):
    gold, out = replay
    m = out["final"]
    assert m.K == 11
    msgs = " ".join(gold["21"]["messages"])
    nums = [float(v) for v in re.findall(r"-?\d+\.\d+", msgs)]
    mean, cov = nums[:2], nums[2:6]
    np.testing.assert_allclose(m.mun[:, 0], mean, atol=5e-9)
    np.testing.assert_allclose(m.Sns[:, :, 0].ravel(), cov, atol=5e-9)
    dens = float(re.findall(r"-?\d+\.\d+", gold["23"]["messages"][0])[0])
    test_point = np.array([[0.1, 0.0]])
    np.testing.assert_allclose(m.predict(test_point), [dens], atol=5e-9)
    cw = m.predict_components(test_point).round(3)
    txt = " ".join(gold["23"]["messages"][1:])
    gold_cw = [float(v) for v in re.findall(r"\d\.\d+e[+-]\d+", txt)]
    np.testing.assert_allclose(cw.ravel(), gold_cw, atol=1e-12)
