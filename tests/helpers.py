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


def headline_parity(X, Y, phi0, K, model, passes=5, method="HS", gather=None, is_root=True):
    """Oracle parity of the product path at an arbitrary (large) shape -- used by tests/test_gpu_scale_parity.py and by
    bench.py's `parity` record (N = 1M, D = 64, K = 64, every rank count).

    `model` is the product MOHGP holding this rank's rows (all rows when there is one rank); `gather(a)` returns the
    row-concatenation of the ranks' arrays on the root (None elsewhere).  The oracle (MOHGPOracle with the whitened-GEMM form of
    multiple_mahalanobis, /root/reference/GPclust/MOHGP.py:124-153) is evaluated on the root over ALL rows:
      bound_rel               |bound - bound_oracle| / |bound_oracle| at the initial point
      natgrad_rel             max-norm relative error of the natural gradient at the initial point (grad_rel: of the gradient)
      trace_iterations_equal  same accept / reject decisions over `passes` loop-body passes (collapsed_vb.py:147-193)
      bound_rel_after         largest relative error of the bound over those passes
      phi_rel_after           max-norm relative error of the responsibilities after them; argmax_equal: same assignments
    Collective when `gather` is given: every rank must call it."""
    import time
    from oracle.vb_oracle import MOHGPOracle, mahalanobis_whitened_gemm
    gather = gather or (lambda a: a)
    t0 = time.perf_counter()
    b0 = model.bound()
    g, ng = model.vb_grad_natgrad()
    g, ng = gather(g.reshape(-1, K)), gather(ng.reshape(-1, K))
    rows = model.run_passes(passes, method=method)
    phi = gather(model.phi)
    b1 = model.bound()
    if not is_root:
        return None
    N = Y.shape[0]
    kF, kY = oracle_kernels()
    mo = MOHGPOracle.__new__(MOHGPOracle)  # as MOHGPOracle(X, kF, kY, Y, K) but starting from phi0 instead of a fresh draw
    mo.D, mo.Y, mo.X, mo.mahalanobis = Y.shape[1], Y, X, mahalanobis_whitened_gemm
    mo.N, mo.K, mo.prior_Z, mo.alpha, mo.hyperparam_interval, mo.trace, mo.out = N, K, "symmetric", 1.0, 50, [], io.StringIO()
    mo.phi_ = np.array(phi0, dtype=np.float64).reshape(N, K)
    mo._softmax_state()
    mo.kernF, mo.kernY = kF, kY
    mo.parameters_changed()
    out = {"N": int(N), "D": int(Y.shape[1]), "K": int(K), "passes": int(passes), "method": method}
    bo = mo.bound()
    out["bound_rel"] = abs(b0 - bo) / abs(bo)
    go, ngo = mo.vb_grad_natgrad()
    out["natgrad_rel"] = relerr(ng.reshape(-1), ngo)
    out["grad_rel"] = relerr(g.reshape(-1), go)
    del go, ngo, g, ng
    st = {"iteration": 0, "bound_old": bo, "searchDir_old": 0.0}
    orows = []
    for _ in range(passes):
        bound, sq, beta, grad, natgrad, sd = mo._loop_pass(st, method, 1.0)
        orows.append((st["iteration"], float(bound), float(sq), float(beta)))
        mo._carry(st, bound, sq, grad, natgrad, sd)
    out["trace_iterations_equal"] = [int(r[0]) for r in rows] == [r[0] for r in orows]
    out["bound_rel_after"] = max(abs(a[1] - b[1]) / abs(b[1]) for a, b in zip(rows, orows)) if rows else None
    out["beta_abs_after"] = max(abs(a[3] - b[3]) for a, b in zip(rows, orows)) if rows else None
    out["phi_rel_after"] = relerr(phi, mo.phi)
    out["argmax_equal"] = bool(np.array_equal(np.argmax(phi, 1), np.argmax(mo.phi, 1)))
    out["final_bound"] = float(b1)
    out["final_bound_oracle"] = float(mo.bound())
    out["seconds"] = time.perf_counter() - t0
    return out
