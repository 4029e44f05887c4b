"""Row ("series") sharding over a torch.distributed process group: one process per GPU, every rank holds a
contiguous block of rows, everything K- or D-sized is replicated.

Per bound evaluation the ranks exchange ONE all-reduce of the statistics vector
[T (KP x DP) | phi_hat (KP) | H] and, per gradient evaluation, one all-reduce of the three conjugate-gradient
dot products.  The backend is whatever the group was created with: NCCL over NVLink on the GPU box, gloo in
the CPU tests of this host logic (tests/test_sharding_gloo.py).
"""
import numpy as np


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

    def allreduce(self, t):
        """In-place sum over ranks (no-op for a single process)."""
        if self.world > 1:
            self.dist.all_reduce(t, group=self.group)
        return t

    @staticmethod
    def even_split(n_total, world, rank):
        """Contiguous near-even split used by bench.py: rows [lo, hi) of rank."""
        return rank * n_total // world, (rank + 1) * n_total // world
