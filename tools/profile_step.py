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

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 4
method = sys.argv[3] if len(sys.argv) > 3 else "HS"
X, Y, phi0 = bench.synth(rows, bench.D, bench.K)
m = MOHGP(X, kern.RBF(1, 1.0, 2.0), kern.RBF(1, 0.1, 1.0) + kern.White(1, 0.05), Y, K=bench.K, phi_init=phi0)
st = {"b": m.bound(), "sq": 0.0, "first": True}
for i in range(passes + 2):
    if i == 2:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
    b, sq, beta, ev = m._cg_iteration(st["first"], method, 1.0, st["b"], st["sq"])
    st.update(b=b, sq=sq, first=False)
torch.cuda.synchronize()
print("rows %d: %.3f ms/pass, bound %.6f, launches %d" % (rows, (time.perf_counter() - t0) / passes * 1e3, b, m.kernel_launches))
