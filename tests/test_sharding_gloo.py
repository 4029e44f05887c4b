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

        # OMGP: components shared out over the ranks (component i -> rank i % world); the owners' per-component bound terms,
        # gradient columns and a failure flag combine into the single-process values (OMGP.py:96-147 through the oracle)
        from gpclust_b200.sharding import ComponentShards
        from oracle.vb_oracle import OMGPOracle
        from oracle import gpy_shim as G
        cs = ComponentShards()
        assert [cs.owns(i) for i in range(4)] == [(i % world) == rank for i in range(4)]
        rng = np.random.RandomState(5)
        Xo = np.sort(rng.uniform(0, 10, 40))[:, None]
        Yo = np.sin(Xo) + 0.1 * rng.randn(40, 1)
        np.random.seed(11)
        mo = OMGPOracle(Xo, Yo, K=3, kernels=[G.RBF(1, 1.0, 1.0 + 0.5 * i) for i in range(3)], variance=0.05)
        comp = torch.zeros(3, 2, dtype=torch.float64)
        cols = torch.zeros(40, 3, dtype=torch.float64)
        for i, kern in enumerate(mo.kern):
            if not cs.owns(i):
                continue
            B = kern.K(Xo) + np.diag(1.0 / ((mo.phi[:, i] + 1e-6) / mo.variance))
            Bi, LB, _, logdet = G.pdinv(B)
            comp[i, 0] = -0.5 * G.dpotrs(LB, mo.YYT)[0].trace() - 0.5 * logdet
            comp[i, 1] = -0.5 * mo.D * mo.phi[:, i].sum() * np.log(2 * np.pi * mo.variance)
            alpha = G.dpotrs(LB, Yo)[0]
            cols[:, i] = torch.from_numpy(-0.5 * mo.variance * (np.sum(np.square(alpha), 1) - np.diag(Bi)) / (mo.phi[:, i] ** 2 + 1e-6))
        cs.combine(comp, cols)
        bound = float(comp.sum()) + mo.mixing_prop_bound() + mo.H
        assert abs(bound - mo.bound()) < 1e-12 * abs(mo.bound())
        g_o, ng_o = mo.vb_grad_natgrad()
        g_s, ng_s = mo._natural(cols.numpy() + mo.mixing_prop_bound_grad() + mo.Hgrad)
        assert np.allclose(ng_s, ng_o, rtol=1e-12, atol=1e-12)
        assert cs.any_failed(rank == 1) is True and cs.any_failed(False) is False
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
