"""Small shapes through every hand-rolled synchronisation protocol, for compute-sanitizer (one tool per run):
   fused loop kernel (grid barriers, ring mbarriers, phase re-carving of shared memory)      MOHGP K <= 64
   wide mode (thread-block clusters, st.async mailboxes in distributed shared memory)          MOHGP K = 130
   launch-sequence loop (graph replay off), MOG posterior, OMGP blocked Cholesky
Usage: compute-sanitizer --tool memcheck|racecheck python tools/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from helpers import mohgp_problem, product_kernels
from gpclust_b200 import MOG, MOHGP, OMGP, kern

os.environ["GPCLUST_B200_GRAPH"] = "0"
out = []
for (N, D, K, fused) in [(1500, 64, 64, "1"), (700, 24, 7, "1"), (700, 24, 7, "0"), (900, 32, 130, "1")]:
    os.environ["GPCLUST_B200_FUSED"] = fused
    X, Y, _ = mohgp_problem(N, D, K, seed=1)
    np.random.seed(2)
    kF, kY = product_kernels()
    m = MOHGP(X, kF, kY, Y, K=K)
    n = m.optimize(maxiter=6, verbose=False)
    g, ng = m.vb_grad_natgrad()
    out.append(("MOHGP N=%d D=%d K=%d fused=%s" % (N, D, K, fused), n, m.bound(), float(np.abs(ng).max())))
rng = np.random.RandomState(0)
Xm = rng.randn(600, 2) * 0.3 + rng.randint(0, 3, (600, 1)) * 2.0
np.random.seed(3)
m = MOG(Xm, K=5)
out.append(("MOG", m.optimize(maxiter=6, verbose=False), m.bound(), 0.0))
Xo = np.sort(rng.uniform(0, 10, 200))[:, None]
Yo = np.sin(Xo) + 0.1 * rng.randn(200, 1)
np.random.seed(4)
m = OMGP(Xo, Yo, K=2, kernels=[kern.RBF(1, 1.0, 1.0).fix(), kern.RBF(1, 1.0, 2.0).fix()], variance=0.05)
m.variance_fixed = True
out.append(("OMGP", m.optimize(maxiter=3, verbose=False), m.bound(), 0.0))
for o in out:
    print("%-34s iterations %d bound %.10g max|natgrad| %.3e" % o)
print("SANITIZE_SMALL_OK")
