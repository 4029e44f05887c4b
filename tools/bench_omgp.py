"""Config 4 (BASELINE.json): OMGP synthetic data association, K = 8 overlapping GPs.  Times one bound + gradient evaluation
(all K components: build, blocked Cholesky, triangular inverse, per-component terms) for growing N and prints one JSON line
per N, with the oracle (numpy/LAPACK) timed beside it at the sizes it finishes in seconds.
Usage: python tools/bench_omgp.py [N ...]      (under torchrun the K components are shared out over the ranks, one per GPU at 8)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from gpclust_b200 import OMGP, kern

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
K = int(os.environ.get("OMGP_K", "8"))  # config 4 has K = 8; OMGP_K trims the component count for the max-fit run (time is linear in K)
sizes = [int(a) for a in sys.argv[1:]] or [2048, 8192, 16384]
for N in sizes:
    rng = np.random.RandomState(0)
    X = np.sort(rng.uniform(0, 10, N))[:, None]
    comp = rng.randint(0, K, N)
    Y = (np.sin(X[:, 0] * (0.5 + 0.3 * comp) + comp) * (1 + 0.2 * comp) + 0.1 * rng.randn(N))[:, None]
    np.random.seed(0)
    t0 = time.perf_counter()
    m = OMGP(X, Y, K=K, kernels=[kern.RBF(1, 1.0, 0.5 + 0.25 * i) for i in range(K)], variance=0.01)
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    x = m._x.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 2 if N < 60000 else 1
    for _ in range(reps):
        p = m._evaluate(x)
        ng, gr = m._natgrad(p, None, None)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    flops = K * (2.0 / 3.0) * float(m._Npad) ** 3  # Cholesky N^3/3 + triangular inverse N^3/3 per component
    line = {"workload": "OMGP bound+gradient evaluation, K=%d RBF components, D=1" % K, "n_gpus": world, "N": N, "Npad": m._Npad, "s_per_evaluation": dt,
            "fp64_tflops": flops / dt / 1e12, "first_evaluation_s": t_first, "bound": p.bound,
            "matrix_gb": 2 * 8.0 * m._Npad ** 2 / 1e9}
    if N <= 4096 and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle import gpy_shim as G
        from oracle.vb_oracle import OMGPOracle
        np.random.seed(0)
        mo = OMGPOracle(X, Y, K=K, kernels=[G.RBF(1, 1.0, 0.5 + 0.25 * i) for i in range(K)], variance=0.01)
        t0 = time.perf_counter()
        b = mo.bound()
        g = mo.vb_grad_natgrad()
        line["cpu_oracle_s_per_evaluation"] = time.perf_counter() - t0
        line["bound_rel_diff_vs_oracle"] = abs(b - m.bound()) / abs(b)
    if rank == 0:
        print(json.dumps(line), flush=True)
    del m
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
