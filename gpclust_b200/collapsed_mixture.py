"""Host mirror of GPclust/collapsed_mixture.py (class CollapsedMixture) with the variational state
resident in HBM.

What lives where
  device (torch CUDA float64 tensors, fragment-major, padded -- include/gpclust_b200.h "Layout"):
      logits phi_ (three rotating buffers: accepted / previous accepted / trial), natural gradient
      (current / previous), search direction, the feature rows F, per-CTA partials, the reduced
      statistics [T | phi_hat | H], per-cluster posterior outputs, the control block that is read
      back once per bound evaluation.
  host: loop control, K bookkeeping, the merge/split search (collapsed_mixture.py:96-205).

Reference attributes (`phi`, `phi_`, `phi_hat`, `H`, `Hgrad`, ...) are properties that copy from the
device on access.  Rows ("series") may be sharded over ranks of a torch.distributed process group:
every rank holds a contiguous block of rows, the statistics vector and the three dot products are
all-reduced; everything K- or D-sized is replicated.
"""
import ctypes
import os
import sys

import numpy as np
from numpy.linalg import LinAlgError

from . import _cabi
from ._cabi import lib, check, ptr, ptr_off
from .collapsed_vb import CollapsedVB
from . import sharding
from .sharding import RowShards

# control block layout (doubles)
_C_BOUND, _C_MIX, _C_H, _C_SQ, _C_NUM, _C_DEN, _C_BETA, _C_INFO = 0, 1, 2, 3, 4, 5, 6, 8
_CTRL_LEN = 16


class _EventPair(object):
    def __init__(self, owner, name):
        self.owner, self.name = owner, name

    def __enter__(self):
        ev = self.owner._kernel_events
        if ev is not None:
            torch = self.owner.torch
            self.pair = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.pair[0].record()

    def __exit__(self, *exc):
        ev = self.owner._kernel_events
        if ev is not None:
            self.pair[1].record()
            ev.setdefault(self.name, []).append(self.pair)
        return False


class _PassGraph(object):
    """A captured batch of loop-body passes (gpc_graph_* of the C-ABI): one launch replays n passes."""

    def __init__(self, handle, passes, launches):
        self.handle, self.passes, self.launches = handle, passes, launches

    _retired = []  # handles of dropped graphs, destroyed at the next safe point (never from inside a capture / the GC)

    def __del__(self):
        if self.handle:
            _PassGraph._retired.append(self.handle)
        self.handle = None

    @staticmethod
    def reap():
        while _PassGraph._retired:
            lib.gpc_graph_destroy(_PassGraph._retired.pop())


class CollapsedMixture(CollapsedVB):
    """collapsed_mixture.py:15-205.  N is the number of rows held by THIS rank."""

    def __init__(self, N, K, prior_Z="symmetric", alpha=1.0, name="col_mix", phi_init=None, group=None, before_upload=None):
        CollapsedVB.__init__(self, name)
        self.torch = _cabi.require_cuda()
        torch = self.torch
        self.N, self.K = int(N), int(K)
        assert prior_Z in ["symmetric", "DP"]
        self.prior_Z = prior_Z
        self.alpha = alpha
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._grid = lib.gpc_default_grid(self.device.index)
        if self._grid <= 0:
            check(self._grid or -1, "gpc_default_grid")

        # row sharding (one process per GPU); a single process is world 1
        self._shards = RowShards(group)
        self._world, self._rank = self._shards.world, self._shards.rank
        self._row_start, self.N_total = self._shards.partition(self.N, self.device)

        self._Npad = _cabi.padded_rows(self.N)
        self._F = None          # Npad x DP features, set by the subclass
        self._DP = None
        self._kbuf_K = None
        self._ctrl = _cabi.dzeros(self.device, _CTRL_LEN)
        self._ctrl_host = torch.zeros(_CTRL_LEN, dtype=torch.float64).pin_memory()
        self._counter = torch.zeros(4, dtype=torch.int32, device=self.device)
        self._gpart = _cabi.dzeros(self.device, self._grid * 16)  # 4 doubles per CTA, up to 4 CTAs per SM
        self._state_valid = False
        self._bound_value = None
        self._peers = None      # peer-memory exchange descriptor (rows sharded over the GPUs of one box), see sharding.py
        self._loop_dev = None   # gpc_loop_t + trace rows (device-resident loop state), allocated on first optimize()
        self._loop_host = None
        self._graphs = {}       # captured pass batches, keyed by (passes, K, alpha, prior); dropped whenever a buffer is re-allocated
        self._cap_stream = None
        self._sync_words = torch.zeros(4, dtype=torch.int32, device=self.device)  # grid-barrier / exit / ticket counters of the fused loop kernel
        self._launch_events = None  # bench.py: (start, end) CUDA events recorded around the launches of run_passes / optimize
        self._loop_timing = None  # bench.py: 8 x int64 device accumulators filled by the fused loop kernel (ns per phase)
        self.kernel_launches = 0  # launches of this package's own kernels (bench.py reports it)
        self._kernel_events = None  # bench.py: {"natgrad": [(start, end), ...], "step_stats": [...]} CUDA events

        # random initial conditions for the vb parameters (collapsed_mixture.py:41): the legacy global
        # numpy stream, drawn for ALL rows so that a sharded run starts from the same point
        if before_upload is not None:
            before_upload()  # subclass hook: device work that does not depend on the rows, started before they cross PCIe
        if phi_init is None:
            phi_init = self._draw_rows(self.K)
        self._alloc_k(self.K)
        self._upload_logits(phi_init)

    # ------------------------------------------------------------------ sharding helpers
    def _is_root(self):
        return self._rank == 0

    def _draw_rows(self, K):
        return self._shards.draw_rows(self.N, self._row_start, self.N_total, K)

    def _allreduce(self, t):
        self._shards.allreduce(t)

    # ------------------------------------------------------------------ device buffers
    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def _alloc_k(self, K):
        torch = self.torch
        KP = _cabi.padded_k(K)
        if self._kbuf_K is not None and KP == self._KP:
            self._kbuf_K = K
            return
        self._KP = KP
        self._kbuf_K = K
        self._graphs = {}
        z = lambda *shape: _cabi.dzeros(self.device, *shape)
        self._xbuf = [z(self._Npad, KP) for _ in range(3)]
        self._lsebuf = [z(self._Npad) for _ in range(3)]  # row logsumexp of the matching logits buffer
        self._xi, self._xprev, self._xtrial = 0, None, 1
        self._ngbuf = [z(self._Npad, KP) for _ in range(2)]
        self._ngi = 0
        self._sd = z(self._Npad, KP)
        self._bias = z(KP)
        self._per_k = z(KP, 4)
        self._mixgrad = z(KP)
        self._chunks = (KP + 63) // 64 if KP > 64 else 1   # K > 64: clusters in chunks of 64 (gpc_*_chunk entry points)
        if self._chunks > 1:
            self._sa_vec = z(self._Npad)
            self._hpart = z(self._grid)
            self._dots_scratch = z(4)
        self._alloc_kd()

    def _alloc_kd(self):
        """Buffers whose size depends on both KP and DP (needs the features to be known)."""
        if self._DP is None:
            return
        torch = self.torch
        KP, DP = self._KP, self._DP
        self._graphs = {}
        self._stats_len = lib.gpc_stats_len(KP, DP)
        self._stats = _cabi.dzeros(self.device, self._stats_len)
        self._spart = _cabi.dzeros(self.device, self._grid * self._stats_len)
        # K > 64: one launch per pass by thread-block clusters of KP/64 CTAs (gpc_*_wide); GPCLUST_B200_WIDE=0 keeps the
        # per-chunk launches (gpc_*_chunk) for comparison
        self._wide = self._chunks > 1 and os.environ.get("GPCLUST_B200_WIDE", "1") != "0"
        self._nparts = self._grid  # per-CTA (per-cluster in wide mode) blocks of partial statistics
        if self._wide:
            self._ncl = int(lib.gpc_wide_clusters(KP, DP))
            if self._ncl < 1:
                check(self._ncl or -1, "gpc_wide_clusters")
            self._nparts = self._ncl
        self._peers = self._shards.setup_peers(self._stats_len, self.device) if (self._use_peer_exchange() and (self._chunks == 1 or self._wide)) else None
        self._Wt = _cabi.dzeros(self.device, KP, DP)

        self._alloc_model()

    def _alloc_model(self):
        """Subclass hook: per-cluster output buffers (K x D x D ...)."""

    def _use_peer_exchange(self):
        """Subclass switch: True when the model's posterior kernel can consume the peer inbox directly."""
        return False

    def _set_features(self, F, DP):
        self._F, self._DP = F, DP
        self._alloc_kd()

    def _upload_logits(self, x):
        """x: numpy (N*K,) / (N,K) or a CUDA tensor of that shape -> padded current buffer; CG history reset."""
        torch = self.torch
        K = self.K
        if K != self._kbuf_K:
            self._alloc_k(K)
        cur = self._xbuf[self._xi]
        if isinstance(x, torch.Tensor):
            src = x.to(device=self.device, dtype=torch.float64).reshape(self.N, K).contiguous()
            check(lib.gpc_pack_rows(ptr(src), self.N, K, self._KP, _cabi.LAYOUT["logits"], ptr(cur), self._stream()), "gpc_pack_rows")
            self.kernel_launches += 1
        else:
            # host array: chunked, pipelined staging -> PCIe -> layout conversion (see _cabi._Ingest)
            self.kernel_launches += _cabi.upload_pack(_cabi.f64(np.asarray(x).reshape(self.N, K)), K, self._KP, _cabi.LAYOUT["logits"], cur,
                                                      self.device, self._stream())
        self._xprev = None
        self._xtrial = (self._xi + 1) % 3
        self._state_valid = False
        self._pass_state = None
        self._cache = {}

    # ------------------------------------------------------------------ reference API: parameters
    def set_vb_param(self, phi_):
        """collapsed_mixture.py:49-61: accept the flattened variational parameters and recompute."""
        self._upload_logits(phi_)
        self.do_computations()

    def get_vb_param(self):
        return self.phi_.flatten()  # collapsed_mixture.py:63-64

    def _unpad(self, buf):
        out = self.torch.empty((self.N, self.K), dtype=self.torch.float64, device=self.device)
        check(lib.gpc_unpack_rows(ptr(buf), self.N, self.K, self._KP, _cabi.LAYOUT["logits"], ptr(out), self._stream()), "gpc_unpack_rows")
        self.kernel_launches += 1
        return out

    @property
    def phi_(self):
        return _cabi.to_host(self._unpad(self._xbuf[self._xi]))

    @phi_.setter
    def phi_(self, value):
        value = np.asarray(value)
        self.K = value.shape[1] if value.ndim == 2 else self.K
        self._upload_logits(value)

    def phi_device(self):
        """N x K responsibilities as a CUDA tensor (no host copy)."""
        return self._softmax_readout()[0]

    def _softmax_readout(self):
        if "softmax" not in self._cache:
            with _cabi.nvtx_range("gpclust.phi_readout"):
                self._cache["softmax"] = self._softmax_readout_now()
        return self._cache["softmax"]

    def _softmax_readout_now(self):
        torch = self.torch
        phi = torch.empty((self.N, self.K), dtype=torch.float64, device=self.device)
        logphi = torch.empty((self.N, self.K), dtype=torch.float64, device=self.device)
        grid = min(self._grid * 4, (self.N + 7) // 8)
        hpart = torch.zeros(grid, dtype=torch.float64, device=self.device)
        x = self._unpad(self._xbuf[self._xi])
        check(lib.gpc_softmax_rows(ptr(x), self.N, self.K, self.K, ptr(phi), ptr(logphi), self.K, ptr(hpart), grid,
                                   self._stream()), "gpc_softmax_rows")
        self.kernel_launches += 1
        return (phi, logphi, hpart)

    @property
    def phi(self):
        return _cabi.to_host(self._softmax_readout()[0])

    @property
    def Hgrad(self):
        return _cabi.to_host(-self._softmax_readout()[1])  # collapsed_mixture.py:56

    def _ensure_state(self):
        if not self._state_valid:
            self.do_computations()

    @property
    def phi_hat(self):
        self._ensure_state()
        o = self._KP * self._DP
        return self._stats[o:o + self.K].cpu().numpy()

    @property
    def H(self):
        self._ensure_state()
        return float(self._ctrl_host[_C_H])

    @property
    def phi_tilde_plus_hat(self):
        return self.phi_hat[::-1].cumsum()[::-1]  # collapsed_mixture.py:58

    @property
    def phi_tilde(self):
        return self.phi_tilde_plus_hat - self.phi_hat  # collapsed_mixture.py:59

    def mixing_prop_bound(self):
        """collapsed_mixture.py:66-78, evaluated by the finalize kernel."""
        self._ensure_state()
        return float(self._ctrl_host[_C_MIX])

    def mixing_prop_bound_grad(self):
        """collapsed_mixture.py:80-94, evaluated by the finalize kernel."""
        self._ensure_state()
        return self._mixgrad[:self.K].cpu().numpy()

    def bound(self):
        self._ensure_state()
        return self._bound_value

    # ------------------------------------------------------------------ device passes
    def _launch_cluster(self):
        raise NotImplementedError

    def _launch_finalize(self):
        raise NotImplementedError

    def _timed(self, name):
        """Context manager recording a CUDA-event pair around one launch on the current stream (bench only)."""
        return _EventPair(self, name)

    def _s_pass(self, x_base, ng, sd_old, sd_new, x_new, lse_new, beta_mode, step, sq_old):
        with self._timed("step_stats"):
            self._launch_step_stats(x_base, ng, sd_old, sd_new, x_new, lse_new, beta_mode, step, sq_old)
        with self._timed("posterior"):
            self._launch_posterior()

    def _launch_step_stats(self, x_base, ng, sd_old, sd_new, x_new, lse_new, beta_mode, step, sq_old):
        if self._wide:
            check(lib.gpc_step_stats_wide(ptr(self._F), ptr(x_base), ptr(ng), ptr(sd_old), ptr(sd_new), ptr(x_new), ptr(lse_new),
                                          ptr(self._ctrl[_C_SQ:]), float(sq_old), ptr(self._ctrl[_C_BETA:]), ptr(self._spart), float(step),
                                          _cabi.BETA[beta_mode], self.N, self.K, self._KP, self._DP, self._ncl, self._stream()),
                  "gpc_step_stats_wide")
            self.kernel_launches += 1
            return
        if self._chunks > 1:
            return self._launch_step_stats_chunked(x_base, ng, sd_old, sd_new, x_new, lse_new, beta_mode, step, sq_old)
        check(lib.gpc_step_stats(ptr(self._F), ptr(x_base), ptr(ng), ptr(sd_old), ptr(sd_new), ptr(x_new), ptr(lse_new), ptr(self._ctrl[_C_SQ:]),
                                 float(sq_old), ptr(self._ctrl[_C_BETA:]), ptr(self._spart), float(step), _cabi.BETA[beta_mode], self.N,
                                 self.K, self._KP, self._DP, self._grid, self._stream()), "gpc_step_stats")
        self.kernel_launches += 1

    def _launch_step_stats_chunked(self, x_base, ng, sd_old, sd_new, x_new, lse_new, beta_mode, step, sq_old):
        """K > 64: CG step + row logsumexp + entropy over whole rows, then the statistics of one 64-cluster chunk per launch."""
        KP, DP, st = self._KP, self._DP, self._stream()
        check(lib.gpc_step_lse(ptr(x_base), ptr(ng), ptr(sd_old), ptr(sd_new), ptr(x_new), ptr(lse_new), ptr(self._ctrl[_C_SQ:]), float(sq_old),
                               ptr(self._ctrl[_C_BETA:]), ptr(self._hpart), float(step), _cabi.BETA[beta_mode], self.N, self.K, KP, self._grid, st),
              "gpc_step_lse")
        x_at = x_new if ng is not None else x_base
        for c in range(self._chunks):
            kc = min(64, self.K - 64 * c)
            if kc <= 0:
                break
            check(lib.gpc_stats_chunk(ptr(self._F), ptr_off(x_at, 512 * c), ptr(lse_new), ptr(self._hpart) if c == 0 else None, ptr(self._spart),
                                      self._stats_len, 64 * c * DP, KP * DP + 64 * c, KP * DP + KP, 1 if c == 0 else 0, self.N, kc, DP, 8 * KP,
                                      self._grid, st), "gpc_stats_chunk")
        self.kernel_launches += 1 + self._chunks

    def _launch_natgrad_wide_or_chunked(self, x, lse, x_old, lse_old, ng_old, ng_out, grad_out, dots):
        if not self._wide:
            return self._launch_natgrad_chunked(x, lse, x_old, lse_old, ng_old, ng_out, grad_out, dots)
        check(lib.gpc_natgrad_wide(ptr(self._F), ptr(x), ptr(lse), ptr(self._Wt), ptr(self._bias), ptr(x_old), ptr(lse_old), ptr(ng_old),
                                   ptr(ng_out), ptr(grad_out), ptr(self._gpart), ptr(self._counter), dots, self.N, self.K, self._KP, self._DP,
                                   self._ncl, self._stream()), "gpc_natgrad_wide")
        self.kernel_launches += 1

    def _launch_natgrad_chunked(self, x, lse, x_old, lse_old, ng_old, ng_out, grad_out, dots):
        """K > 64: a = F.W + bias - x per chunk (row sums of phi*a accumulated), then centring, gradient and dots over whole rows."""
        KP, DP, st = self._KP, self._DP, self._stream()
        for c in range(self._chunks):
            kc = min(64, self.K - 64 * c)
            if kc <= 0:
                break
            check(lib.gpc_natgrad_chunk(ptr(self._F), ptr_off(x, 512 * c), ptr(lse), ptr_off(self._Wt, 64 * c * DP), ptr_off(self._bias, 64 * c),
                                        ptr_off(ng_out, 512 * c), ptr(self._sa_vec), 1 if c == 0 else 0, ptr(self._gpart), ptr(self._counter),
                                        ptr(self._dots_scratch), self.N, kc, DP, 8 * KP, self._grid, st), "gpc_natgrad_chunk")
        check(lib.gpc_natgrad_finish(ptr(ng_out), ptr(x), ptr(lse), ptr(self._sa_vec), ptr(x_old), ptr(lse_old), ptr(ng_old), ptr(grad_out),
                                     ptr(self._gpart), ptr(self._counter), dots, self.N, self.K, KP, self._grid * 4, st), "gpc_natgrad_finish")
        self.kernel_launches += 1 + self._chunks

    def _launch_posterior(self):
        check(lib.gpc_reduce_partials(ptr(self._spart), self._nparts, self._stats_len, ptr(self._stats), self._stream()), "gpc_reduce_partials")
        self.kernel_launches += 1
        self._allreduce(self._stats)
        self._launch_cluster()
        self._launch_finalize()

    def _readback(self):
        """One small D2H copy per bound evaluation: bound, mix, H, the CG dots, beta, the not-PD flag."""
        self._ctrl_host.copy_(self._ctrl, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        h = self._ctrl_host
        info = int(h.view(self.torch.int32)[2 * _C_INFO])
        if info & _cabi.GPC_INFO_PEER_TIMEOUT:
            self._state_valid = False
            raise _cabi.GpclustLibraryError("gpclust_b200: a rank of the NVLink peer exchange never arrived (timeout)")
        return float(h[_C_BOUND]), float(h[_C_SQ]), float(h[_C_BETA]), info != 0

    def _info_ptr(self):
        return ptr(self._ctrl[_C_INFO:])

    def do_computations(self):
        """Statistics, per-cluster posterior and bound at the current logits (MOHGP.py:70-90 / MOG.py:52-61).
        Raises numpy.linalg.LinAlgError when a per-cluster factorisation fails (GPy jitchol semantics)."""
        with _cabi.nvtx_range("gpclust.do_computations"):
            return self._do_computations()

    def _do_computations(self):
        self._s_pass(self._xbuf[self._xi], None, None, None, None, self._lsebuf[self._xi], "zero", 0.0, 0.0)
        bound, _, _, failed = self._readback()
        self._cache = {}
        if failed:
            self._state_valid = False
            raise LinAlgError("not positive definite, even with jitter.")
        self._bound_value = bound
        self._state_valid = True

    # ---- CG hooks (collapsed_vb.py:152-193 on the device)
    def _cg_gradient(self, with_old):
        self._ensure_state()
        have_old = with_old and self._xprev is not None
        with self._timed("natgrad"):
            self._launch_natgrad(have_old)
        if self._world > 1:
            self._allreduce(self._ctrl[_C_SQ:_C_SQ + 3])

    def _launch_natgrad(self, have_old):
        x, lse = self._xbuf[self._xi], self._lsebuf[self._xi]
        x_old = self._xbuf[self._xprev] if have_old else None
        lse_old = self._lsebuf[self._xprev] if have_old else None
        ng_old = self._ngbuf[1 - self._ngi] if have_old else None
        if self._chunks > 1:
            return self._launch_natgrad_wide_or_chunked(x, lse, x_old, lse_old, ng_old, self._ngbuf[self._ngi], None, ptr(self._ctrl[_C_SQ:]))
        check(lib.gpc_natgrad(ptr(self._F), ptr(x), ptr(lse), ptr(self._Wt), ptr(self._bias), ptr(x_old), ptr(lse_old), ptr(ng_old),
                              ptr(self._ngbuf[self._ngi]), None, ptr(self._gpart), ptr(self._counter), ptr(self._ctrl[_C_SQ:]), self.N,
                              self.K, self._KP, self._DP, self._grid, self._stream()), "gpc_natgrad")
        self.kernel_launches += 1

    def _cg_trial(self, beta_mode, step_length, sq_old):
        mode = "zero" if beta_mode == "steepest_retry" else beta_mode
        self._s_pass(self._xbuf[self._xi], self._ngbuf[self._ngi], self._sd, self._sd, self._xbuf[self._xtrial], self._lsebuf[self._xtrial],
                     mode, step_length, sq_old)
        bound, sq, beta, failed = self._readback()
        self._cache = {}
        self._trial_failed = failed
        self._trial_bound = bound
        return bound, sq, beta, failed

    def _cg_restore(self):
        """collapsed_vb.py:189-190: set_vb_param(phi_old); bound = self.bound()."""
        self.do_computations()
        self._xbuf[self._xtrial].copy_(self._xbuf[self._xi])
        self._lsebuf[self._xtrial].copy_(self._lsebuf[self._xi])
        self._trial_bound = self._bound_value
        self._trial_failed = False
        return self._bound_value

    def _cg_failed_pass_scalars(self):
        h = self._ctrl_host  # filled by _loop_sync_host: the dots of the G pass and the conjugate beta of the failed pass
        return float(h[_C_SQ]), float(h[_C_BETA])

    def _cg_accept(self):
        free = self._xprev if self._xprev is not None else 3 - self._xi - self._xtrial
        self._xprev, self._xi, self._xtrial = self._xi, self._xtrial, free
        self._ngi = 1 - self._ngi
        self._bound_value = self._trial_bound
        self._state_valid = not self._trial_failed
        self._cache = {}

    # ------------------------------------------------------------------ device-resident loop (collapsed_vb.py:143-303)
    _TRACE_CAP = 64   # trace rows kept on the device between two read-backs
    _BATCH = 16       # loop-body passes enqueued per read-back

    def _device_loop_ok(self):
        return self._chunks == 1 or self._wide  # the per-chunk launches (GPCLUST_B200_WIDE=0) run under the host-driven loop

    def _bufs(self):
        b = _cabi.GpcBufs()
        for i in range(3):
            b.x[i] = self._xbuf[i].data_ptr()
            b.lse[i] = self._lsebuf[i].data_ptr()
        for i in range(2):
            b.ng[i] = self._ngbuf[i].data_ptr()
        b.sd = self._sd.data_ptr()
        return b

    def _loop_upload(self, L):
        torch = self.torch
        nbytes = ctypes.sizeof(_cabi.GpcLoop) + self._TRACE_CAP * 32
        if self._loop_dev is None:
            self._graphs = {}
            self._loop_dev = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            self._loop_host = torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
        ctypes.memmove(self._loop_host.data_ptr(), ctypes.addressof(L), ctypes.sizeof(_cabi.GpcLoop))
        self._loop_dev[:ctypes.sizeof(_cabi.GpcLoop)].copy_(self._loop_host[:ctypes.sizeof(_cabi.GpcLoop)], non_blocking=True)

    def _loop_download(self):
        """-> (gpc_loop_t snapshot, trace rows [(iteration, bound, squareNorm, beta), ...]); one D2H copy + sync."""
        self._loop_host.copy_(self._loop_dev, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        raw = ctypes.string_at(self._loop_host.data_ptr(), self._loop_host.numel())
        L = _cabi.GpcLoop.from_buffer_copy(raw[:ctypes.sizeof(_cabi.GpcLoop)])
        rows = np.frombuffer(raw[ctypes.sizeof(_cabi.GpcLoop):], dtype=np.float64).reshape(-1, 4)[:L.n_trace]
        return L, [(int(r[0]), float(r[1]), float(r[2]), float(r[3])) for r in rows]

    def _loop_posterior(self, phase):
        """Subclass hook: statistics -> posterior -> bound -> device-side accept / reject for `phase`."""
        raise NotImplementedError

    def _enqueue_pass(self, bufs):
        """One bound evaluation of the loop body collapsed_vb.py:152-193 (the conjugate step, or the steepest retry of :182-193 when
        the previous one was rejected) as a fixed launch sequence; nothing is read back."""
        lp, st = ptr(self._loop_dev), self._stream()
        dots, beta = ptr(self._ctrl[_C_SQ:]), ptr(self._ctrl[_C_BETA:])
        px = ptr(self._peers["desc"]) if self._peers is not None else None
        g_pass = lib.gpc_loop_natgrad_wide if self._wide else lib.gpc_loop_natgrad_px
        s_pass = lib.gpc_loop_step_stats_wide if self._wide else lib.gpc_loop_step_stats_px
        ctas = self._ncl if self._wide else self._grid  # clusters of KP/64 CTAs in wide mode
        with self._timed("natgrad"):
            check(g_pass(ptr(self._F), ctypes.byref(bufs), ptr(self._Wt), ptr(self._bias), ptr(self._gpart), ptr(self._counter), dots,
                         lp, px, self.N, self.K, self._KP, self._DP, ctas, st), "gpc_loop_natgrad")
        if self._world > 1 and px is None:
            self._allreduce(self._ctrl[_C_SQ:_C_SQ + 3])
        # slot mode (gpc_loop_t.slot_mode): ONE bound evaluation per enqueued sequence; after a rejected conjugate step the next
        # sequence skips its G pass and runs S / posterior in the RETRY phase (the `phase` arguments are ignored)
        with self._timed("step_stats"):
            check(s_pass(ptr(self._F), ctypes.byref(bufs), dots, beta, ptr(self._spart), lp, px, 0, self.N, self.K,
                         self._KP, self._DP, ctas, st), "gpc_loop_step_stats")
        with self._timed("posterior"):
            self._loop_posterior(0)
        self.kernel_launches += 2
        if not self._slot_mode():
            # the fixed five-launch pass (slot_mode = 0): the RETRY kernels follow every conjugate evaluation and return at once
            # unless the bound fell
            with self._timed("step_stats_retry"):
                check(s_pass(ptr(self._F), ctypes.byref(bufs), dots, beta, ptr(self._spart), lp, px, 1, self.N, self.K,
                             self._KP, self._DP, ctas, st), "gpc_loop_step_stats")
            with self._timed("posterior_retry"):
                self._loop_posterior(1)
            self.kernel_launches += 1

    _GRAPH_PASSES = 16  # passes per captured graph (plus a single-pass graph for the remainder)

    def _graphs_ok(self):
        if os.environ.get("GPCLUST_B200_GRAPH", "1") == "0" or self._kernel_events is not None:
            return False
        return self._world == 1 or self._peers is not None  # the NCCL exchange stays on the eager path

    def _pass_graph(self, bufs, n):
        """The launch sequence of n passes, captured once (every argument is invariant: the kernels pick their
        operands from the device-resident loop state)."""
        key = (n, self.K, float(self.alpha), self.prior_Z, self._slot_mode())
        g = self._graphs.get(key)
        if g is None:
            torch = self.torch
            if self._cap_stream is None:
                self._cap_stream = torch.cuda.Stream(device=self.device)
            launches0 = self.kernel_launches
            _PassGraph.reap()
            with torch.cuda.stream(self._cap_stream):
                st = self._stream()
                check(lib.gpc_graph_begin(st), "gpc_graph_begin")
                handle = ctypes.c_void_p()
                try:
                    for _ in range(n):
                        self._enqueue_pass(bufs)
                finally:
                    rc = lib.gpc_graph_end(st, ctypes.byref(handle))
                check(rc, "gpc_graph_end")
            g = _PassGraph(handle, n, self.kernel_launches - launches0)
            self.kernel_launches = launches0  # nothing ran during the capture
            self._graphs[key] = g
        return g

    def _fused_ok(self):
        """Subclass switch: True when the whole loop runs as one persistent cooperative kernel (gpc_mohgp_loop)."""
        return False

    def _launch_fused(self, bufs, n):
        raise NotImplementedError

    def _enqueue_passes(self, bufs, n):
        if self._fused_ok():
            # ONE launch for the n evaluations: G rows, S rows and the posterior are phases of a persistent kernel separated by
            # grid-wide barriers (csrc/gpc_loop_kernel.cuh)
            self._launch_fused(bufs, n)
            return
        if not self._graphs_ok():
            for _ in range(n):
                self._enqueue_pass(bufs)
            return
        st = self._stream()
        while n > 0:
            g = self._pass_graph(bufs, self._GRAPH_PASSES if n >= self._GRAPH_PASSES else 1)
            check(lib.gpc_graph_launch(g.handle, st), "gpc_graph_launch")
            self.kernel_launches += g.launches
            n -= g.passes

    @staticmethod
    def _slot_mode():
        return os.environ.get("GPCLUST_B200_SLOT", "1") != "0"

    def _loop_sync_host(self, L):
        """Adopt the device loop state on the host side (buffer selectors, bound, control block)."""
        self._xi, self._xprev, self._xtrial, self._ngi = L.xi, (None if L.xprev < 0 else L.xprev), L.xtrial, L.ngi
        self._readback()
        self._bound_value = L.bound_old
        # after GPC_DONE_FAULT the statistics / weights / control block belong to the failed trial point, not to x[xi]
        self._state_valid = not (L.done & _cabi.DONE_FAULT)
        self._cache = {}

    def _run_on_device(self, method, maxiter, ftol, gtol, step_length, iteration, bound_old, sq_old, first, say=False, max_passes=None):
        """Drive the loop from the device-resident state until it terminates (or `max_passes` passes were enqueued).
        -> (iteration, bound_old, sq_old, first, done bits, trace rows)"""
        self._ensure_state()
        L = _cabi.GpcLoop()
        L.bound_old, L.sq_old, L.ftol, L.gtol, L.maxiter, L.step = bound_old, sq_old, ftol, gtol, float(maxiter), step_length
        L.iteration, L.method, L.first = iteration, _cabi.BETA["zero" if method == "steepest" else method], int(first)
        L.xi, L.xprev, L.xtrial, L.ngi = self._xi, (-1 if self._xprev is None else self._xprev), self._xtrial, self._ngi
        L.need_retry = L.done = L.n_trace = 0
        L.trace_cap = self._TRACE_CAP
        L.slot_mode = 1 if (self._slot_mode() or self._fused_ok()) else 0  # the fused loop kernel works in evaluations (slots) only
        L.pass_budget = int(max_passes) if max_passes is not None else 0  # the device raises DONE_BUDGET after that many accepted steps
        self._loop_upload(L)
        bufs = self._bufs()
        rows_all = []
        sharding.claim(self._peers, self)
        try:
            return self._drive_device_loop(L, bufs, rows_all, maxiter, max_passes, say)
        finally:
            sharding.release(self._peers, self)

    def _drive_device_loop(self, L, bufs, rows_all, maxiter, max_passes, say):
        while True:
            # every enqueued sequence is one bound evaluation, which is what `iteration` / `maxiter` count (collapsed_vb.py:177,191)
            remaining = max(1, int(np.ceil(float(maxiter) - L.iteration))) if np.isfinite(maxiter) else self._BATCH
            n = min(self._BATCH, remaining)
            if max_passes is not None:
                left = max_passes - len(rows_all)
                # accepted steps wanted + room for rejected conjugate steps (each costs a sequence of its own in slot mode)
                n = min(self._TRACE_CAP, left + max(2, left // 4)) if (self._slot_mode() or self._fused_ok()) else left
            ev = self._launch_events
            if ev is not None and not rows_all:
                ev[0].record()  # bench.py: the timed region starts when the first launch is enqueued (loop-state set-up is not a step)
            with _cabi.nvtx_range("gpclust.vb_loop[%d evaluations]" % n):
                self._enqueue_passes(bufs, n)
            if ev is not None:
                ev[1].record()  # ... and ends behind the last one (re-recorded should a second batch be needed)
            L, rows = self._loop_download()
            rows_all.extend(rows)
            if say:
                for it, b, sq, beta in rows:
                    sys.stdout.write("\riteration " + str(it) + " bound=" + str(b) + " grad=" + str(sq) + ", beta=" + str(beta) + "\n")
                sys.stdout.flush()
            if L.done & _cabi.DONE_PEER_TIMEOUT:
                self._state_valid = False
                raise _cabi.GpclustLibraryError("gpclust_b200: a rank of the NVLink peer exchange never arrived (timeout on rank %d of %d); "
                                                "the optimisation state of this model is invalid" % (self._rank, self._world))
            if L.done:
                break
            # rewind the trace cursor for the next batch (a 4-byte field; everything else stays device-owned)
            off = _cabi.GpcLoop.n_trace.offset
            self._loop_dev[off:off + 4].zero_()
        self._loop_sync_host(L)
        return L.iteration, L.bound_old, L.sq_old, bool(L.first), L.done, rows_all

    def run_passes(self, n, method="HS", step_length=1.0):
        """Exactly n passes of the loop body with the termination tests disabled (bench / profiling): one enqueue,
        one read-back.  -> trace rows."""
        self._ensure_state()
        st = getattr(self, "_pass_state", None)
        if st is None or st["xi"] != self._xi:
            st = {"iteration": 0, "bound_old": self.bound(), "sq_old": 0.0, "first": True}
        cap, self._TRACE_CAP = self._TRACE_CAP, max(self._TRACE_CAP, n + max(2, n // 4))
        if self._loop_dev is not None and self._loop_dev.numel() < ctypes.sizeof(_cabi.GpcLoop) + self._TRACE_CAP * 32:
            self._loop_dev = None
        batch, self._BATCH = self._BATCH, n
        try:
            it, b, sq, first, done, rows = self._run_on_device(method, np.inf, -1.0, -1.0, step_length, st["iteration"], st["bound_old"], st["sq_old"],
                                                               st["first"], max_passes=n)
        finally:
            self._BATCH, self._TRACE_CAP = batch, cap
        self._pass_state = {"iteration": it, "bound_old": b, "sq_old": sq, "first": first, "xi": self._xi}
        return rows

    def vb_grad_natgrad(self):
        """Flattened (grad, natgrad) as numpy arrays -- the reference's return convention
        (MOHGP.py:137-153, MOG.py:72-84).  Does not disturb the optimiser's device state."""
        self._ensure_state()
        torch = self.torch
        ng = torch.empty((self._Npad, self._KP), dtype=torch.float64, device=self.device)
        gr = torch.empty((self._Npad, self._KP), dtype=torch.float64, device=self.device)
        dots = torch.zeros(4, dtype=torch.float64, device=self.device)
        if self._chunks > 1:
            self._launch_natgrad_wide_or_chunked(self._xbuf[self._xi], self._lsebuf[self._xi], None, None, None, ng, gr, ptr(dots))
        else:
            check(lib.gpc_natgrad(ptr(self._F), ptr(self._xbuf[self._xi]), ptr(self._lsebuf[self._xi]), ptr(self._Wt), ptr(self._bias), None, None,
                                  None, ptr(ng), ptr(gr),
                                  ptr(self._gpart), ptr(self._counter), ptr(dots), self.N, self.K, self._KP, self._DP, self._grid,
                                  self._stream()), "gpc_natgrad")
            self.kernel_launches += 1
        return _cabi.to_host(self._unpad(gr)).reshape(-1), _cabi.to_host(self._unpad(ng)).reshape(-1)

    # ------------------------------------------------------------------ merge / split search (host control)
    def reorder(self):
        """collapsed_mixture.py:96-103: re-order the clusters so that the biggest one is first (DP only)."""
        if self.prior_Z == "DP":
            i = np.argsort(self.phi_hat)[::-1]
            self.set_vb_param(self.phi_[:, i])

    def remove_empty_clusters(self, threshold=1e-6):
        """collapsed_mixture.py:105-110."""
        i = self.phi_hat > threshold
        phi_ = self.phi_[:, i]
        self.K = int(i.sum())
        self.set_vb_param(phi_)

    def try_split(self, indexK=None, threshold=0.9, verbose=True, maxiter=100, optimize_params=None):
        """collapsed_mixture.py:112-189.  With rows sharded over ranks the membership count and the `special` point are
        taken over ALL rows: every rank draws the same legacy-RNG values, the owner of the special row is found from the
        all-gathered per-rank counts, everything else is local to the rank's rows."""
        verbose = verbose and self._is_root()
        if indexK is None:
            indexK = np.random.multinomial(1, self.phi_hat / self.N_total).argmax()
        if indexK > (self.K - 1):
            return False  # index exceed no. of clusters
        elif self.phi_hat[indexK] < 1:
            return False  # no data to split
        phi = self.phi
        indexN = np.nonzero(phi[:, indexK] > threshold)[0]
        counts = self._shards.allgather_int(len(indexN), self.device)
        n_members = int(counts.sum())
        if n_members < 2:
            return False  # only one data point
        if verbose:
            print("\nattempting to split cluster ", indexK)
        bound_old = self.bound()
        phi_old = self.get_vb_param().copy()
        param_old = self.optimizer_array.copy()  # collapsed_mixture.py:143-144
        old_K = self.K

        # re-initalize
        logits = phi_old.reshape(self.N, old_K)
        logits = np.hstack((logits, logits.min(1)[:, None]))
        # np.random.permutation(indexN)[0] of the reference: the first entry of a shuffle depends on the length only
        j = int(np.random.permutation(n_members)[0])
        before = int(counts[:self._rank].sum())
        logits[indexN, -1] = logits[indexN, indexK].copy()
        if before <= j < before + len(indexN):
            special = indexN[j - before]
            logits[special, -1] = np.max(logits[special]) + 10
        self.K = old_K + 1
        self.set_vb_param(logits)

        self.optimize(maxiter=maxiter, verbose=verbose)
        self.remove_empty_clusters()
        bound_new = self.bound()

        bound_increase = bound_new - bound_old
        if bound_increase < 1e-3:
            self.K = old_K
            self.set_vb_param(phi_old)
            if param_old.size:
                self.optimizer_array = param_old  # collapsed_mixture.py:174
            if verbose:
                print("split failed, bound changed by: ", bound_increase, "(K=%s)" % self.K)
            return False
        if verbose:
            print("split suceeded, bound changed by: ", bound_increase, ",", self.K - old_K, " new clusters", "(K=%s)" % self.K)
            print("optimizing new split to convergence:")
        self.last_split_increase = bound_increase
        if optimize_params:
            self.optimize(**optimize_params)
        else:
            self.optimize(maxiter=5000, verbose=verbose)
        return True

    def systematic_splits(self, verbose=True):
        """collapsed_mixture.py:191-195."""
        for kk in range(self.K):
            self.recursive_splits(kk, verbose=verbose)

    def recursive_splits(self, k=0, verbose=True, optimize_params=None):
        """collapsed_mixture.py:197-205."""
        success = self.try_split(k, verbose=verbose, optimize_params=optimize_params)
        if success:
            if not k == (self.K - 1):
                self.recursive_splits(self.K - 1, verbose=verbose, optimize_params=optimize_params)
            self.recursive_splits(k, verbose=verbose, optimize_params=optimize_params)
