"""Print max-norm-relative differences between the GPU path and the CPU oracle for a few MOHGP shapes, including
large-phi_hat cases where the reference's own (Sy - t^T t)/phi_hat is at its rounding-noise floor.
Usage: python tools/parity_probe.py [N D K ...triples]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from helpers import mohgp_problem, oracle_kernels, product_kernels, relerr
from oracle.vb_oracle import MOHGPOracle, mahalanobis_whitened_gemm
from gpclust_b200 import MOHGP

args = [int(a) for a in sys.argv[1:]]
shapes = list(zip(args[0::3], args[1::3], args[2::3])) or [(1000, 64, 64), (5000, 33, 17), (20000, 64, 2), (60000, 64, 4), (40000, 24, 3)]
for N, D, K in shapes:
    X, Y, _ = mohgp_problem(N, D, K, 0)
    kF, kY = oracle_kernels()
    np.random.seed(1)
    mo = MOHGPOracle(X, kF, kY, Y, K=K, mahalanobis=mahalanobis_whitened_gemm)
    pF, pY = product_kernels()
    np.random.seed(1)
    mp = MOHGP(X, pF, pY, Y, K=K)
    out = {"N": N, "D": D, "K": K, "phi_hat_max": float(mo.phi_hat.max())}
    for it in range(2):
        g_p, ng_p = mp.vb_grad_natgrad()
        g_o, ng_o = mo.vb_grad_natgrad()
        out["it%d" % it] = dict(bound=abs(mp.bound() - mo.bound()) / abs(mo.bound()), muk=relerr(mp.muk, mo.muk),
                                ldd=relerr(mp.log_det_diff, mo.log_det_diff), Linv=relerr(mp.Lambda_inv, mo.Lambda_inv),
                                ng=relerr(ng_p, ng_o), g=relerr(g_p, g_o))
        import io
        mo.out = io.StringIO()
        mo.optimize(maxiter=6, verbose=False)
        mp.optimize(maxiter=6, verbose=False)
        out["phi_hat_max%d" % it] = float(mo.phi_hat.max())
    print(out, flush=True)
