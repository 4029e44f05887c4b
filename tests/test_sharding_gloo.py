"""World-size-2 gloo test (CPU) of the multi-rank host logic: row partition, the shared initial-condition
stream, and the statistic / dot-product all-reduces that combine per-shard results.  The per-shard numbers
come from the CPU oracle (the CUDA kernels need a GPU); what is checked is that sharding + all-reduce
reproduce the single-process statistics, bound and CG dots."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import mohgp_problem, oracle_kernels


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gpclust_b200.sharding import RowShards
        from oracle.vb_oracle import MOHGPOracle, softmax_rows
        N, D, K = 301, 12, 5
        X, Y, _ = mohgp_problem(N, D, K, seed=2)
        sh = RowShards()
        lo, hi = RowShards.even_split(N, world, rank)
        start, total = sh.partition(hi - lo)
        assert (start, total) == (lo, N)
        np.random.seed(7)
        phi_loc = sh.draw_rows(hi - lo, start, total, K)
        np.random.seed(7)
        phi_full = np.random.randn(N, K)
        assert np.array_equal(phi_loc, phi_full[lo:hi])

        # per-shard sufficient statistics [T | phi_hat | H] -> all-reduce == full-data statistics
        phi, logphi, H = softmax_rows(phi_loc)
        stats = np.concatenate([phi.T.dot(Y[lo:hi]).ravel(), phi.sum(0), [H]])
        t = torch.from_numpy(stats.copy())
        sh.allreduce(t)
        phi_f, _, H_f = softmax_rows(phi_full)
        full = np.concatenate([phi_f.T.dot(Y).ravel(), phi_f.sum(0), [H_f]])
        np.testing.assert_allclose(t.numpy(), full, rtol=1e-12, atol=1e-12)

        # the bound and the CG dots assembled from all-reduced pieces equal the single-process values
        kF, kY = oracle_kernels()
        np.random.seed(7)
        m = MOHGPOracle(X, kF, kY, Y, K=K)
        g, ng = m.vb_grad_natgrad()
        g2, ng2 = g.reshape(N, K)[lo:hi], ng.reshape(N, K)[lo:hi]
        dots = torch.tensor([float((g2 * ng2).sum())], dtype=torch.float64)
        sh.allreduce(dots)
        assert abs(dots.item() - float(np.dot(g, ng))) < 1e-10 * abs(float(np.dot(g, ng)))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_single_process_is_world_one():
    from gpclust_b200.sharding import RowShards
    sh = RowShards()
    assert (sh.world, sh.rank) == (1, 0)
    assert sh.partition(17) == (0, 17)
    np.random.seed(3)
    a = sh.draw_rows(4, 0, 4, 3)
    np.random.seed(3)
    assert np.array_equal(a, np.random.randn(4, 3))
    assert RowShards.even_split(10, 3, 0) == (0, 3) and RowShards.even_split(10, 3, 2) == (6, 10)
