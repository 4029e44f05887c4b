"""Row ("series") sharding over a torch.distributed process group: one process per GPU, every rank holds a
contiguous block of rows, everything K- or D-sized is replicated.

Per bound evaluation the ranks exchange ONE all-reduce of the statistics vector
[T (KP x DP) | phi_hat (KP) | H] and, per gradient evaluation, one all-reduce of the three conjugate-gradient
dot products.  The backend is whatever the group was created with: NCCL over NVLink on the GPU box, gloo in
the CPU tests of this host logic (tests/test_sharding_gloo.py).
"""
import ctypes
import os
import warnings

import numpy as np


# (group object | "WORLD", world, stats_len, device index) -> exchange.  The group object itself is part of the key (held strongly:
# an id() could be reused after the group is garbage-collected).  Models of equal geometry share one inbox and one epoch counter, so
# only one of them may be inside its device loop at a time: `claim` / `release` enforce it.
_PEER_CACHE = {}


def claim(peers, owner):
    if peers is None:
        return
    cur = peers.get("owner")
    if cur is not None and cur is not owner:
        raise RuntimeError("gpclust_b200: the NVLink peer exchange of this geometry is in use by another model's device loop; "
                           "models that share a process group run their optimisations one at a time")
    peers["owner"] = owner


def release(peers, owner):
    if peers is not None and peers.get("owner") is owner:
        peers["owner"] = None


class RowShards(object):
    def __init__(self, group=None):
        self.group = group
        self.dist = None
        self.world, self.rank = 1, 0
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.dist = dist
                self.world = dist.get_world_size(group)
                self.rank = dist.get_rank(group)
        except ImportError:
            pass

    @classmethod
    def solo(cls):
        """A single-process instance regardless of any initialised process group (models that do not shard rows)."""
        self = cls.__new__(cls)
        self.group, self.dist, self.world, self.rank = None, None, 1, 0
        return self

    def partition(self, n_local, device=None):
        """-> (row_start, N_total) for this rank given its local row count."""
        if self.world == 1:
            return 0, int(n_local)
        import torch
        t = torch.zeros(self.world, dtype=torch.int64, device=device)
        t[self.rank] = int(n_local)
        self.dist.all_reduce(t, group=self.group)
        counts = t.cpu().numpy()
        return int(counts[:self.rank].sum()), int(counts.sum())

    def draw_rows(self, n_local, row_start, n_total, K):
        """The reference's initial condition `np.random.randn(N, K)` (collapsed_mixture.py:41) for this rank's
        rows: every rank draws the SAME global legacy stream and keeps its block, so that a sharded run starts
        from the point a single-process run would."""
        if self.world == 1:
            return np.random.randn(n_local, K)
        out = None
        chunk = max(1, (1 << 22) // max(K, 1))
        pos = 0
        keep = []
        while pos < n_total:
            n = min(chunk, n_total - pos)
            block = np.random.randn(n, K)
            lo, hi = max(row_start, pos), min(row_start + n_local, pos + n)
            if lo < hi:
                keep.append(block[lo - pos:hi - pos].copy())
            pos += n
        out = np.concatenate(keep, 0) if keep else np.zeros((0, K))
        return out

    def setup_peers(self, stats_len, device):
        """Peer-memory exchange descriptor (gpc_peers_t, include/gpclust_b200.h) for the device-resident loop: one
        symmetric-memory inbox per rank, mapped into every process of the box.  Collective.  -> dict or None (single
        process, more than 8 ranks, symmetric memory unavailable, or GPCLUST_B200_PEER=0: the NCCL path is used)."""
        if self.world == 1 or self.world > 8 or os.environ.get("GPCLUST_B200_PEER", "1") == "0":
            return None
        from . import _cabi
        key = (self.group if self.group is not None else "WORLD", self.world, int(stats_len), getattr(device, "index", None))
        if key in _PEER_CACHE:
            return _PEER_CACHE[key]
        try:
            import torch
            import torch.distributed._symmetric_memory as symm
            group = self.group if self.group is not None else self.dist.group.WORLD
            n = int(_cabi.lib.gpc_xchg_doubles_fused(self.world, int(stats_len)))  # serves both exchange layouts (fused loop kernel / launch sequence)
            buf = symm.empty(n, dtype=torch.float64, device=device)
            hdl = symm.rendezvous(buf, group)
            buf.zero_()
            torch.cuda.synchronize()
            self.dist.barrier(group=self.group)  # nobody pushes before every inbox has been cleared
            epochs = torch.zeros(2, dtype=torch.int64, device=device)
            P = _cabi.GpcPeers()
            P.world, P.rank, P.stats_len = self.world, self.rank, int(stats_len)
            ptrs = list(hdl.buffer_ptrs)
            for r in range(self.world):
                P.inbox[r] = ptrs[r]
            P.epochs = epochs.data_ptr()
            raw = torch.frombuffer(bytearray(bytes(P)), dtype=torch.uint8).to(device)
            _PEER_CACHE[key] = {"desc": raw, "buf": buf, "handle": hdl, "epochs": epochs, "owner": None}
            return _PEER_CACHE[key]
        except Exception as e:  # pragma: no cover - depends on the box
            warnings.warn("gpclust_b200: peer-memory exchange unavailable (%s: %s); using NCCL all-reduces" % (type(e).__name__, e))
            return None

    def allgather_int(self, value, device=None):
        """-> int64 numpy array of `value` on every rank (rank order)."""
        if self.world == 1:
            return np.array([int(value)], dtype=np.int64)
        import torch
        t = torch.zeros(self.world, dtype=torch.int64, device=device)
        t[self.rank] = int(value)
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def allreduce(self, t):
        """In-place sum over ranks (no-op for a single process)."""
        if self.world > 1:
            self.dist.all_reduce(t, group=self.group)
        return t

    @staticmethod
    def even_split(n_total, world, rank):
        """Contiguous near-even split used by bench.py: rows [lo, hi) of rank."""
        return rank * n_total // world, (rank + 1) * n_total // world


class ComponentShards(RowShards):
    """OMGP: the dense N x N work of one component does not shard, but the components are
    independent -- component i is owned by rank i % world.  Data and responsibilities are replicated; the owners'
    results (per-component bound scalars, the grad_Lm column, kernel gradients) are DISJOINT contributions that one
    all-reduce(sum) assembles on every rank, and a failed factorisation on any rank is raised on all of them."""

    def owns(self, i):
        return i % self.world == self.rank

    def combine(self, *tensors):
        """In-place sum over ranks of tensors that are zero outside this rank's components."""
        for t in tensors:
            self.allreduce(t)

    def any_failed(self, failed, like=None):
        """True on every rank when `failed` is true on at least one."""
        if self.world == 1:
            return bool(failed)
        import torch
        t = torch.tensor([1.0 if failed else 0.0], dtype=torch.float64, device=None if like is None else like.device)
        self.allreduce(t)
        return bool(t.item() > 0.0)
