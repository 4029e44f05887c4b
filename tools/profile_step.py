"""Short driver for ncu: config-3 shaped MOHGP (rows configurable), a few warm-up passes and a few timed
passes of the VB loop body.  Usage: python tools/profile_step.py [rows] [passes]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from gpclust_b200 import MOHGP, kern

nrows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
method = sys.argv[3] if len(sys.argv) > 3 else "HS"
X, Y, phi0 = bench.synth(nrows, bench.D, bench.K)
m = MOHGP(X, kern.RBF(1, 1.0, 2.0).fix(), (kern.RBF(1, 0.1, 1.0) + kern.White(1, 0.05)).fix(), Y, K=bench.K, phi_init=phi0)
m.bound()
m.run_passes(2, method=method)
torch.cuda.synchronize()
t0 = time.perf_counter()
rows = m.run_passes(passes, method=method)
torch.cuda.synchronize()
print("rows %d: %.3f ms/pass, bound %.6f, launches %d" % (Y.shape[0], (time.perf_counter() - t0) / passes * 1e3, rows[-1][1], m.kernel_launches))
