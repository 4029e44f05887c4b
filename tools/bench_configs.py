"""BASELINE.json configs 1 and 2 on the GPU path (information only: both are far too small to load a B200 -- the metric is config 3):
  config 1 stand-in   MOHGP, N = 500 series, D = 56 replicated time points (data/kalinka09_pdata.csv design), K = 20, Matern52, DP prior
  config 2            MOG on notebooks/twoD_clustering_example.numpy (506 x 2), K = 10
Prints one JSON line per config: iterations/s of optimize() through the public API (device-resident loop) and the oracle's rate on the
host cores beside it; profiles/cpu_reference_r02.json holds the reference's own CPU rate for config 1 (measured in the build container)."""
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from helpers import GOLDEN, load_mog_data
from oracle import gpy_shim as G
from oracle.vb_oracle import MOGOracle, MOHGPOracle
from gpclust_b200 import MOG, MOHGP, kern


def rate(make, maxiter, reps=3):
    best = None
    for _ in range(reps):
        m = make()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.optimize(maxiter=maxiter, ftol=-1.0, gtol=-1.0, verbose=False)
        b = m.bound()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        r = len(m.optimize_trace) / dt
        best = r if best is None else max(best, r)
    return best, b


rows = [l.strip().split(",") for l in open(os.path.join(GOLDEN, "kalinka09_pdata.csv")).read().strip().split("\n")]
hours = np.array([float(r[2]) for r in rows if r[0] == "me"])[:56, None]
D = hours.shape[0]
rng = np.random.RandomState(0)
kFo, kYo = G.Matern52(1, 1.0, 12.0), G.Matern52(1, 0.1, 12.0) + G.White(1, 0.05)
means = np.linalg.cholesky(kFo.K(hours) + 1e-6 * np.eye(D)).dot(rng.randn(D, 6)).T
Y = means[rng.randint(0, 6, 500)] + np.linalg.cholesky(kYo.K(hours) + 1e-8 * np.eye(D)).dot(rng.randn(D, 500)).T
Y = (Y - Y.mean(1)[:, None]) / Y.std(1)[:, None]
np.random.seed(2)
phi0 = np.random.randn(500, 20)


def mk1():
    return MOHGP(hours, kern.Matern52(1, 1.0, 12.0).fix(), (kern.Matern52(1, 0.1, 12.0) + kern.White(1, 0.05)).fix(), Y, K=20, prior_Z="DP", phi_init=phi0)


r1, b1 = rate(mk1, 60)
np.random.seed(2)
mo = MOHGPOracle(hours, kFo, kYo, Y, K=20, prior_Z="DP")
mo.out = io.StringIO()
t0 = time.perf_counter()
mo.optimize(maxiter=60, ftol=-1.0, gtol=-1.0, verbose=False)
ro = len(mo.trace) / (time.perf_counter() - t0)
print(json.dumps({"config": "1 (stand-in): MOHGP N=500 D=56 K=20 Matern52 DP", "gpu_iterations_per_s": r1, "oracle_cpu_iterations_per_s": ro, "cores": os.cpu_count(),
                  "bound_gpu": b1, "bound_oracle": mo.bound(), "bound_rel_diff": abs(b1 - mo.bound()) / abs(mo.bound())}), flush=True)

Xm = load_mog_data()
np.random.seed(0)
phi0m = np.random.randn(Xm.shape[0], 10)
r2, b2 = rate(lambda: MOG(Xm, K=10, phi_init=phi0m), 60)
np.random.seed(0)
mo = MOGOracle(Xm, K=10)
mo.set_vb_param(phi0m.flatten())
mo.out = io.StringIO()
t0 = time.perf_counter()
mo.optimize(maxiter=60, ftol=-1.0, gtol=-1.0, verbose=False)
ro = len(mo.trace) / (time.perf_counter() - t0)
print(json.dumps({"config": "2: MOG twoD_clustering_example (506 x 2) K=10", "gpu_iterations_per_s": r2, "oracle_cpu_iterations_per_s": ro, "cores": os.cpu_count(),
                  "bound_gpu": b2, "bound_oracle": mo.bound(), "bound_rel_diff": abs(b2 - mo.bound()) / abs(mo.bound())}), flush=True)
