"""Config 5 (BASELINE.json): MOHGP at K in {64,...,512}.  Prints one JSON line per (N, K): ms per loop-body pass on this
rank's rows (host-driven loop for K > 64, device-resident loop for K <= 64) and the achieved HBM / FP64-tensor rates.
Usage: python tools/bench_config5.py N K [K ...]   (single process; under torchrun rows are sharded like bench.py)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from gpclust_b200 import MOHGP, kern

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
N = N_total // world  # rows of this rank (its own draws: timing run, not a parity run)
Ks = [int(a) for a in sys.argv[2:]] or [64, 128, 256, 512]
D = 64
for K in Ks:
    rng = np.random.RandomState(0)
    X = np.linspace(0, 10, D)[:, None]
    from oracle import gpy_shim as G
    Sf = G.RBF(1, 1.0, 2.0).K(X)
    Sy = (G.RBF(1, 0.1, 1.0) + G.White(1, 0.05)).K(X)
    means = (np.linalg.cholesky(Sf + 1e-8 * np.eye(D)).dot(rng.randn(D, K))).T
    rng = np.random.RandomState(1 + rank)
    z = rng.randint(0, K, N)
    Y = means[z] + (np.linalg.cholesky(Sy).dot(rng.randn(D, N))).T
    torch.manual_seed(rank)
    phi0 = torch.randn((N, K), dtype=torch.float64, device="cuda")  # device RNG: the host stream would take minutes at K=512
    m = MOHGP(X, kern.RBF(1, 1.0, 2.0).fix(), (kern.RBF(1, 0.1, 1.0) + kern.White(1, 0.05)).fix(), Y, K=K, phi_init=phi0)
    del phi0
    passes, warm = 32, 16
    if m._device_loop_ok():
        m.run_passes(warm)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rows = m.run_passes(passes)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / passes
        rej = int(rows[-1][0] - rows[0][0] + 1) - passes
        bound = rows[-1][1]
    else:
        st = {"b": m.bound(), "sq": 0.0, "first": True}
        rej = 0
        for i in range(warm + passes):
            if i == warm:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                rej = 0
            b, sq, beta, ev = m._cg_iteration(st["first"], "HS", 1.0, st["b"], st["sq"])
            st.update(b=b, sq=sq, first=False)
            rej += ev - 1
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / passes
        bound = b
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    alg_bytes = 8.0 * N * world * (9 * K + 2 * D)
    if rank == 0:
      print(json.dumps({"workload": "MOHGP HS-CG pass, FP64", "n_gpus": world, "N": N * world, "rows_per_gpu": N, "D": D, "K": K, "chunks": m._chunks, "ms_per_pass": dt * 1e3,
                      "rejected": rej, "iterations_per_s": 1.0 / dt, "alg_GBps": alg_bytes / dt / 1e9,
                      "fp64_tensor_TFLOPs": 4.0 * N * world * K * D / dt / 1e12, "bound": bound}), flush=True)
    del m
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
