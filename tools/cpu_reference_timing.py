"""CPU timing of the REFERENCE ITSELF (/root/reference/GPclust imported unmodified, GPy replaced by the stub of
tests/golden/make_reference_goldens.py) -- runs in the build container only (the reference does not travel to the GPU box).
What runs is what a Python-3 user of the reference gets: np_utilities.multiple_mahalanobis_numpy_loops (a Python double loop with one
np.linalg.solve per (n, k) pair, np_utilities.py:43-59), softmax_numpy, numpy / LAPACK for everything else, and the dead N x D x D
`YYT` allocation of MOHGP.py:52.
  config 3 shape (D = 64, K = 64, RBF + White) on a row sample, scaled linearly in N (every N-dependent term is linear)
  config 1 stand-in (N = 500 series, D = 56 replicated time points of data/kalinka09_pdata.csv, K = 20, Matern52, DP prior)
Writes profiles/cpu_reference_r02.json."""
import io
import json
import os
import sys
import time
import contextlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np

import make_reference_goldens as mrg
from oracle import gpy_shim as G

mrg.install_gpy_stub()
sys.path.insert(0, mrg.REF)
import GPclust  # noqa: E402  (the reference's own package)

assert os.path.realpath(GPclust.__file__).startswith(os.path.realpath(mrg.REF))
out = {"cores": os.cpu_count(), "host": "build container (no GPU)", "reference": "GPclust imported unmodified from /root/reference, stub GPy"}


def time_optimize(m, maxiter):
    buf = io.StringIO()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(buf):
        m.optimize(maxiter=maxiter, verbose=True)
    dt = time.perf_counter() - t0
    rows = mrg.LINE.findall(buf.getvalue().replace("\r", "\n"))
    return dt, len(rows), (int(rows[-1][0]) if rows else 0)


# ---- config 3 shape on a row sample
from bench import synth
N_s = int(os.environ.get("REF_ROWS", "1500"))
X, Y, phi0 = synth(N_s, 64, 64)
np.random.seed(0)
m = GPclust.MOHGP(X, G.RBF(1, 1.0, 2.0), G.RBF(1, 0.1, 1.0) + G.White(1, 0.05), Y, K=64)
m.set_vb_param(phi0.flatten())
dt, passes, evals = time_optimize(m, 3)
per_pass = dt / max(passes, 1)
out["config3_reference_as_is_python3"] = {
    "rows": N_s, "passes": passes, "bound_evaluations": evals, "s_per_pass_on_sample": per_pass,
    "s_per_pass_at_1M_rows": per_pass * 1e6 / N_s, "iterations_per_s_at_1M_rows": 1.0 / (per_pass * 1e6 / N_s),
    "note": "np_utilities.multiple_mahalanobis_numpy_loops: one np.linalg.solve per (n, k) pair; scaled linearly in N"}
print(json.dumps(out["config3_reference_as_is_python3"]), flush=True)

# ---- config 1 stand-in
rows = [l.strip().split(",") for l in open(os.path.join(ROOT, "tests", "golden", "kalinka09_pdata.csv")).read().strip().split("\n")]
hours = np.array([float(r[2]) for r in rows if r[0] == "me"])[:56, None]
D = hours.shape[0]
rng = np.random.RandomState(0)
kF, kY = G.Matern52(1, 1.0, 12.0), G.Matern52(1, 0.1, 12.0) + G.White(1, 0.05)
Sf, Sy = kF.K(hours), kY.K(hours)
means = np.linalg.cholesky(Sf + 1e-6 * np.eye(D)).dot(rng.randn(D, 6)).T
Yk = means[rng.randint(0, 6, 500)] + np.linalg.cholesky(Sy + 1e-8 * np.eye(D)).dot(rng.randn(D, 500)).T
Yk = (Yk - Yk.mean(1)[:, None]) / Yk.std(1)[:, None]
np.random.seed(2)
m = GPclust.MOHGP(hours, kF, kY, Yk, K=20, prior_Z="DP", alpha=1.0)
dt, passes, evals = time_optimize(m, 6)
out["config1_standin_reference_as_is_python3"] = {
    "N": 500, "D": int(D), "K": 20, "kernels": "Matern52 / Matern52 + White, DP prior (drosophila.ipynb)", "passes": passes, "bound_evaluations": evals,
    "s_per_pass": dt / max(passes, 1), "iterations_per_s": max(passes, 1) / dt, "final_bound": float(m.bound())}
print(json.dumps(out["config1_standin_reference_as_is_python3"]), flush=True)
json.dump(out, open(os.path.join(ROOT, "profiles", "cpu_reference_r02.json"), "w"), indent=1)
