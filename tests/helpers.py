"""Shared test helpers: synthetic MOHGP data (SURVEY.md section 8d recipe), comparison in max-norm-relative terms."""
import io
import json
import os

import numpy as np

from oracle import gpy_shim as G

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relerr(a, b):
    """max-norm relative error |a-b|_inf / |b|_inf (SURVEY.md section 7 'Numerical parity at 1e-9')."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    denom = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (denom if denom > 0 else 1.0))


def mohgp_problem(N, D, K, seed=0, kern_mod=None):
    """Synthetic hierarchical-GP data: K cluster mean functions drawn from kernF, series = mean + kernY noise.
    Returns (X, Y, spec) with spec describing the kernels so both the oracle shim and the product can build them."""
    rng = np.random.RandomState(seed)
    X = np.linspace(0, 10, D)[:, None]
    spec = {"F": ("rbf", 1.0, 2.0), "Y": [("rbf", 0.1, 1.0), ("white", 0.05, 1.0)]}
    kF = G.RBF(1, 1.0, 2.0)
    kY = G.RBF(1, 0.1, 1.0) + G.White(1, 0.05)
    Sf, Sy = kF.K(X), kY.K(X)
    z = rng.randint(0, K, N)
    means = (np.linalg.cholesky(Sf + 1e-8 * np.eye(D)).dot(rng.randn(D, K))).T
    Y = means[z] + (np.linalg.cholesky(Sy).dot(rng.randn(D, N))).T
    return X, np.ascontiguousarray(Y), spec


def oracle_kernels():
    return G.RBF(1, 1.0, 2.0), G.RBF(1, 0.1, 1.0) + G.White(1, 0.05)


def product_kernels(fixed=True):
    """fixed=True: kernel hyper-parameters held fixed (the oracle's default; the reference would need kern.fix())."""
    from gpclust_b200 import kern
    kF, kY = kern.RBF(1, 1.0, 2.0), kern.RBF(1, 0.1, 1.0) + kern.White(1, 0.05)
    return (kF.fix(), kY.fix()) if fixed else (kF, kY)


def load_golden():
    with open(os.path.join(GOLDEN, "mog_notebook_golden.json")) as f:
        return json.load(f)["cells"]


def load_mog_data():
    return np.load(os.path.join(GOLDEN, "twoD_clustering_example.npy"))


def quiet():
    return io.StringIO()
