import sys, io, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
np.set_printoptions(linewidth=200, precision=10)
import test_gpu_omgp as T
from gpclust_b200 import OMGP
orig = OMGP._factor_component
def patched(self, phi, i, eps):
    r = orig(self, phi, i, eps)
    return r
for (N, Dy, K, kernels) in [(50, 1, 2, None), (130, 2, 3, T._mixed_kernels)]:
    mo, mp = T._pair(N, Dy, K, free=True, kernels=kernels, variance=0.05)
    print("init oracle ", mo.kernel_gradients())
    print("init product", mp._kernel_gradients())
    mo.hyper_free = False
    mp.variance_fixed = True
    for k in mp.kern: k.fix()
    mo.out = io.StringIO()
    mo.optimize(maxiter=6, verbose=False); mp.optimize(maxiter=6, verbose=False)
    mp.variance_fixed = False
    for k in mp.kern: k.unfix()
    print("phi min", mo.phi.min(), mp.phi.min(), "relerr phi", np.abs(mo.phi-mp.phi).max())
    print("post oracle ", mo.kernel_gradients())
    print("post product", mp._kernel_gradients())
    # per-component factor check
    import torch
    p = mp._cur
    for i in range(K):
        mp._factor_component(p.phi, i, 0.0)
        Kmat = mo.kern[i].K(mo.X)
        B = Kmat + np.diag(1.0 / (mo.phi[:, i] / mo.variance))
        L = np.linalg.cholesky(B)
        Winv = np.linalg.inv(L)
        Wd = mp._W[:N, :N].cpu().numpy()
        print(" comp", i, "cond diag range", B.diagonal().min(), B.diagonal().max(), "W relerr", np.abs(Wd - Winv).max() / np.abs(Winv).max())
