"""Wall-clock breakdown of the end-to-end call sequence of bench.py (construct from plain numpy arrays -- or pinned ones with --pin --,
optimize, read back)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from gpclust_b200 import MOHGP, kern

X, Y, phi0 = bench.synth(1000000, bench.D, bench.K)
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lo, hi = rank * 1000000 // world, (rank + 1) * 1000000 // world
    Y, phi0 = np.ascontiguousarray(Y[lo:hi]), np.ascontiguousarray(phi0[lo:hi])
if "--pin" in sys.argv:
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    Y, phi0 = pin(Y), pin(phi0)


def T(label, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if rank == 0:
        print("  %-28s %8.1f ms" % (label, (t1 - t0) * 1e3), flush=True)
    return t1


for rep in range(3):
    if world > 1:
        dist.barrier()
    if rank == 0:
        print("rep", rep, flush=True)
    t = time.perf_counter()
    m = MOHGP(X, kern.RBF(1, 1.0, 2.0).fix(), (kern.RBF(1, 0.1, 1.0) + kern.White(1, 0.05)).fix(), Y, K=bench.K, phi_init=phi0)
    t = T("construct", t)
    m.optimize(method="HS", maxiter=20, ftol=-1.0, gtol=-1.0, verbose=False)
    t = T("optimize(20)", t)
    phi = m.phi
    t = T("m.phi", t)
    b = m.bound()
    t = T("m.bound()", t)
    del m
if world > 1:
    dist.destroy_process_group()
