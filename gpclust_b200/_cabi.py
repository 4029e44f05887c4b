"""ctypes binding of include/gpclust_b200.h -- the only way the host code reaches the GPU.

There is no CPU fallback: if the library is missing the import raises, and every compute call
requires a CUDA device (the launchers return a negative cudaError otherwise, which `check` raises).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPCLUST_B200_LIB") or os.path.join(_HERE, "_lib", "libgpclust_b200.so")  # env override: kernel-variant experiments

GPC_ROW_TILE = 64
GPC_EINVAL = 2
GPC_INFO_NOT_PD = 1
GPC_INFO_PEER_TIMEOUT = 2
PRIOR = {"symmetric": 0, "DP": 1}
BETA = {"zero": 0, "HS": 1, "PR": 2, "FR": 3}
LAYOUT = {"features": 0, "logits": 1}
KERN = {"rbf": 0, "matern32": 1, "matern52": 2, "bias": 3, "white": 4}

_p = ctypes.c_void_p
_i = ctypes.c_int
_ll = ctypes.c_longlong
_d = ctypes.c_double

# name -> argtypes; mirrors include/gpclust_b200.h one to one (tests/test_cabi.py checks both sides)
PROTOTYPES = {
    "gpc_abi_version": [],
    "gpc_padded_k": [_i],
    "gpc_padded_d": [_i],
    "gpc_padded_rows": [_ll],
    "gpc_stats_len": [_i, _i],
    "gpc_default_grid": [_i],
    "gpc_natgrad": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _p],
    "gpc_step_stats": [_p, _p, _p, _p, _p, _p, _p, _p, _d, _p, _p, _d, _i, _ll, _i, _i, _i, _i, _p],
    "gpc_loop_natgrad": [_p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _p],
    "gpc_loop_step_stats": [_p, _p, _p, _p, _p, _p, _i, _ll, _i, _i, _i, _i, _p],
    "gpc_loop_natgrad_px": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _p],
    "gpc_loop_step_stats_px": [_p, _p, _p, _p, _p, _p, _p, _i, _ll, _i, _i, _i, _i, _p],
    "gpc_loop_natgrad_wide": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _p],
    "gpc_loop_step_stats_wide": [_p, _p, _p, _p, _p, _p, _p, _i, _ll, _i, _i, _i, _i, _p],
    "gpc_reduce_push": [_p, _i, _i, _p, _p, _p, _i, _p],
    "gpc_xchg_doubles": [_i, _i],
    "gpc_xchg_doubles_fused": [_i, _i],
    "gpc_mohgp_loop": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _d, _i, _p, _p, _p, _p, _i, _ll, _i, _i, _i, _i,
                       _i, _p],
    "gpc_natgrad_chunk": [_p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _ll, _i, _i, _ll, _i, _p],
    "gpc_natgrad_finish": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _i, _i, _p],
    "gpc_step_lse": [_p, _p, _p, _p, _p, _p, _p, _d, _p, _p, _d, _i, _ll, _i, _i, _i, _p],
    "gpc_stats_chunk": [_p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _ll, _i, _i, _ll, _i, _p],
    "gpc_wide_clusters": [_i, _i],
    "gpc_natgrad_wide": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _p],
    "gpc_step_stats_wide": [_p, _p, _p, _p, _p, _p, _p, _p, _d, _p, _p, _d, _i, _ll, _i, _i, _i, _i, _p],
    "gpc_reduce_partials": [_p, _i, _i, _p, _p],
    "gpc_mohgp_setup": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _p],
    "gpc_mohgp_const": [_p, _p, _p, _d, _i, _p, _p],
    "gpc_mohgp_cluster": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "gpc_mohgp_eig_ws_len": [_i],
    "gpc_mohgp_eig": [_p, _p, _p, _p, _p, _i, _p, _p],
    "gpc_mohgp_kern_grads_work": [_i, _i],
    "gpc_mohgp_kern_grads": [_p, _p, _p, _p, _p, _p, _d, _i, _i, _i, _i, _p, _p, _p, _p, _p],
    "gpc_mohgp_post": [_p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _d, _i, _i, _i, _i, _i, _p],
    "gpc_loop_mohgp_post": [_p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _d, _i, _i, _i, _i, _i, _p, _p, _p, _i, _p],
    "gpc_loop_mohgp_post_px": [_p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _d, _i, _i, _i, _i, _i, _p, _p, _p, _i, _p],
    "gpc_mohgp_finalize": [_p, _p, _p, _p, _p, _p, _d, _i, _i, _i, _i, _p],
    "gpc_mog_features": [_p, _p, _ll, _i, _i, _p, _p],
    "gpc_mog_cluster": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _d, _d, _i, _i, _i, _i, _p],
    "gpc_mog_finalize": [_p, _p, _p, _p, _p, _d, _d, _d, _d, _d, _i, _i, _i, _i, _i, _p],
    "gpc_loop_mog_post": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _d, _d, _p, _p, _p, _d, _d, _d, _i, _i, _i, _i, _i, _p, _p, _p, _i, _p],
    "gpc_kern_fill": [_p, _p, _i, _i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_d), ctypes.POINTER(_d), _d, _p, _p],
    "gpc_omgp_padded_n": [_ll],
    "gpc_omgp_slabs": [_ll],
    "gpc_omgp_part_stride": [],
    "gpc_omgp_build": [_p, _ll, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_d), ctypes.POINTER(_d), _p, _i, _d, _d, _d, _p, _p],
    "gpc_chol_blocked": [_p, _ll, _p, _p, _p, _p],
    "gpc_tri_inverse_blocked": [_p, _ll, _p, _p, _p],
    "gpc_chol_inverse": [_p, _ll, _p, _p, _p, _p, _p, _p],
    "gpc_omgp_component": [_p, _p, _p, _p, _i, _ll, _i, _d, _p, _p, _p, _i, _p, _p],
    "gpc_omgp_kgrad_tiles": [_ll],
    "gpc_omgp_kern_grads": [_p, _p, _i, _p, _ll, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_d), ctypes.POINTER(_d), _p, _p, _p, _p, _p, _p, _p],
    "gpc_omgp_finalize": [_p, _p, _i, _p, _p, _p, _d, _d, _i, _i, _i, _p],
    "gpc_dense_natgrad": [_p, _p, _p, _p, _p, _p, _ll, _i, _p, _p, _p, _p],
    "gpc_cg_axpy": [_p, _p, _p, _d, _d, _ll, _p, _p, _p],
    "gpc_softmax_rows": [_p, _ll, _i, _i, _p, _p, _i, _p, _i, _p],
    "gpc_multiple_mahalanobis": [_p, _p, _p, _ll, _i, _i, _p, _p],
    "gpc_multiple_pdinv": [_p, _i, _i, _p, _p, _p, _p],
    "gpc_gram_partials": [_p, _ll, _i, _p, _i, _p],
    "gpc_normalise_rows": [_p, _ll, _i, _p, _p],
    "gpc_replicate_scores": [_p, _ll, _i, _p, _i, _p, _p],
    "gpc_pack_rows": [_p, _ll, _i, _i, _i, _p, _p],
    "gpc_unpack_rows": [_p, _ll, _i, _i, _i, _p, _p],
    "gpc_pad_rows": [_p, _ll, _i, _i, _p, _p],
    "gpc_unpad_rows": [_p, _ll, _i, _i, _p, _p],
    "gpc_graph_begin": [_p],
    "gpc_graph_end": [_p, ctypes.POINTER(_p)],
    "gpc_graph_launch": [_p, _p],
    "gpc_graph_destroy": [_p],
}
_RESTYPE = {"gpc_padded_rows": _ll, "gpc_omgp_padded_n": _ll, "gpc_omgp_kgrad_tiles": _ll, "gpc_xchg_doubles": _ll, "gpc_xchg_doubles_fused": _ll, "gpc_mohgp_kern_grads_work": _ll}


class GpcBufs(ctypes.Structure):
    """gpc_bufs_t of include/gpclust_b200.h"""
    _fields_ = [("x", _p * 3), ("lse", _p * 3), ("ng", _p * 2), ("sd", _p)]


class GpcLoop(ctypes.Structure):
    """gpc_loop_t of include/gpclust_b200.h (the trace rows follow it in device memory)"""
    _fields_ = [("bound_old", _d), ("sq_old", _d), ("ftol", _d), ("gtol", _d), ("maxiter", _d), ("step", _d),
                ("iteration", _i), ("method", _i), ("first", _i), ("xi", _i), ("xprev", _i), ("xtrial", _i), ("ngi", _i),
                ("need_retry", _i), ("done", _i), ("n_trace", _i), ("trace_cap", _i), ("slot_mode", _i), ("pass_budget", _i), ("pad", _i)]


class GpcPeers(ctypes.Structure):
    """gpc_peers_t of include/gpclust_b200.h"""
    _fields_ = [("world", _i), ("rank", _i), ("stats_len", _i), ("pad", _i), ("inbox", _p * 8), ("epochs", _p)]


DONE_FTOL, DONE_GTOL, DONE_MAXITER, DONE_FAULT, DONE_BUDGET, DONE_PEER_TIMEOUT = 1, 2, 4, 8, 16, 32


class GpclustLibraryError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise GpclustLibraryError(
            "gpclust_b200: %s is missing -- build it with `python -m gpclust_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise GpclustLibraryError("gpclust_b200: symbol %s missing from %s (stale build?)" % (name, LIB_PATH))
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, _i)
    return lib


lib = _load()


def check(rc, what=""):
    """0 -> ok; GPC_EINVAL -> ValueError; negative -> CUDA error (no device, launch failure, ...)."""
    if rc == 0:
        return
    if rc == GPC_EINVAL:
        raise ValueError("gpclust_b200: invalid argument in %s" % what)
    raise GpclustLibraryError("gpclust_b200: %s failed with CUDA error %d (is a B200 / CUDA device present?)" % (what, -rc))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def padded_k(K):
    kp = lib.gpc_padded_k(int(K))
    if kp < 0:
        raise ValueError("gpclust_b200: K=%d is outside the supported range 1..512" % K)
    return kp


def ptr_off(t, n_doubles):
    """Device pointer n_doubles into a float64 tensor."""
    return ctypes.c_void_p(t.data_ptr() + 8 * int(n_doubles))


def padded_d(D):
    dp = lib.gpc_padded_d(int(D))
    if dp < 0:
        raise ValueError("gpclust_b200: feature dimension %d is outside the supported range 1..64" % D)
    return dp


def padded_rows(N):
    return int(lib.gpc_padded_rows(int(N)))


def as_c_array(values, ctype):
    values = list(values)
    return (ctype * len(values))(*values)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise GpclustLibraryError("gpclust_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no CPU fallback.")
    return torch


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def to_host(t):
    """Device tensor -> numpy array through PINNED host memory (torch's caching host allocator keeps the staging
    blocks, so repeated read-outs of N x K arrays run at PCIe speed instead of the pageable-copy rate)."""
    import torch
    if t.numel() * t.element_size() < (1 << 20):
        return t.cpu().numpy()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return h.numpy()


# ---------------------------------------------------------------------------------------------- host -> device ingest
class _Ingest(object):
    """Pipelined upload of a (pageable) numpy matrix into the fragment-major device layout.

    A plain cudaMemcpy from pageable memory is staged by the driver through one bounce buffer on one thread (a few GB/s).  Here the
    rows are cut into chunks; several host threads copy chunk i+1 into a pinned staging buffer while the DMA engine moves chunk i
    over PCIe and gpc_pack_rows re-orders chunk i-1 into the fragment-major layout -- three stages in flight, all device work
    in order on the caller's stream.  Staging buffers and the thread pool are created once per process."""
    CHUNK_BYTES = int(os.environ.get("GPCLUST_B200_INGEST_CHUNK_MB", "32")) << 20
    NBUF = 3

    def __init__(self):
        self.pinned, self.dev, self.events, self.pool, self.threads = None, {}, None, None, 0

    def _setup(self, torch, device):
        if self.pinned is None:
            from concurrent.futures import ThreadPoolExecutor
            self.threads = max(1, min(int(os.environ.get("GPCLUST_B200_INGEST_THREADS", "8")), (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
            self.pool = ThreadPoolExecutor(self.threads)
            self.pinned = [torch.empty(self.CHUNK_BYTES // 8, dtype=torch.float64, pin_memory=True) for _ in range(self.NBUF)]
            self.events = [torch.cuda.Event() for _ in range(self.NBUF)]
            self.busy = [False] * self.NBUF
        key = (device.type, device.index)
        if key not in self.dev:
            self.dev[key] = [torch.empty(self.CHUNK_BYTES // 8, dtype=torch.float64, device=device) for _ in range(self.NBUF)]
        return self.dev[key]

    def upload_pack(self, src, C, CP, layout, dst, device, stream_ptr):
        """src: C-contiguous float64 numpy (N, C).  dst: device tensor, fragment-major (padded_rows(N) x CP).  Returns launches."""
        import torch
        N = src.shape[0]
        dev = self._setup(torch, device)
        with nvtx_range("gpclust.ingest[%dx%d]" % (N, C)):
            return self._upload_pack(src, N, C, CP, layout, dst, dev, stream_ptr)

    def _upload_pack(self, src, N, C, CP, layout, dst, dev, stream_ptr):
        import torch
        try:
            pinned_src = torch.from_numpy(src)
            if not pinned_src.is_pinned():
                pinned_src = None
        except Exception:
            pinned_src = None
        rows = max(GPC_ROW_TILE, (self.CHUNK_BYTES // (8 * C)) // GPC_ROW_TILE * GPC_ROW_TILE)  # whole 64-row tiles: only the last chunk pads
        launches = 0
        for i, lo in enumerate(range(0, N, rows)):
            hi = min(N, lo + rows)
            b = i % self.NBUF
            n = hi - lo
            d = dev[b][:n * C]
            if pinned_src is not None:
                # the caller's array is page-locked already: DMA straight out of it, no staging copy (device chunk buffers are
                # recycled in stream order)
                d.copy_(pinned_src[lo:hi].reshape(-1), non_blocking=True)
            else:
                if self.busy[b]:
                    self.events[b].synchronize()  # the DMA that last read this staging buffer (this upload's or an earlier one's) has finished
                stage = self.pinned[b][:n * C].view(n, C).numpy()
                if self.threads > 1 and n * C >= (1 << 18):
                    cuts = [lo + n * t // self.threads for t in range(self.threads + 1)]
                    list(self.pool.map(lambda ab: np.copyto(stage[ab[0] - lo:ab[1] - lo], src[ab[0]:ab[1]]), zip(cuts[:-1], cuts[1:])))
                else:
                    np.copyto(stage, src[lo:hi])
                d.copy_(self.pinned[b][:n * C], non_blocking=True)
                self.events[b].record()
                self.busy[b] = True
            check(lib.gpc_pack_rows(ctypes.c_void_p(d.data_ptr()), n, C, CP, layout, ctypes.c_void_p(dst.data_ptr() + 8 * lo * CP), stream_ptr),
                  "gpc_pack_rows")
            launches += 1
        return launches


_ingest = _Ingest()


def upload_pack(src, C, CP, layout, dst, device, stream_ptr):
    return _ingest.upload_pack(src, C, CP, layout, dst, device, stream_ptr)


# ---------------------------------------------------------------------------------------------- guarded device allocations
# compute-sanitizer is closed on this pool (profiles/sanitizer_r02_closed.txt); out-of-bounds WRITES of the kernels are caught instead by
# guard zones: with GUARD on (tests/test_gpu_guards.py, or GPCLUST_B200_GUARD=1) every device array of a model is carved out of a
# larger allocation whose 4 KB margins hold a canary pattern, verified by check_guards() after the kernels have run.
GUARD = os.environ.get("GPCLUST_B200_GUARD", "0") == "1"
_GUARD_DOUBLES = 512
_CANARY = -6.02214076e23
_guarded = []


def dzeros(device, *shape, dtype=None):
    """Zero-initialised device tensor (float64 unless dtype is given); guard zones around it when GUARD is on."""
    import torch
    dtype = dtype or torch.float64
    if not GUARD or dtype != torch.float64:
        return torch.zeros(shape, dtype=dtype, device=device)
    n = int(np.prod(shape)) if len(shape) else 1
    base = torch.full((n + 2 * _GUARD_DOUBLES,), _CANARY, dtype=torch.float64, device=device)
    mid = base[_GUARD_DOUBLES:_GUARD_DOUBLES + n]
    mid.zero_()
    _guarded.append((base, n, shape))
    return mid.view(shape)


def check_guards(clear=False):
    """-> list of (shape, side, number of damaged canary words) over every guarded allocation still alive."""
    import torch
    bad = []
    for base, n, shape in _guarded:
        lo = int((base[:_GUARD_DOUBLES] != _CANARY).sum().item())
        hi = int((base[_GUARD_DOUBLES + n:] != _CANARY).sum().item())
        if lo:
            bad.append((tuple(shape), "below", lo))
        if hi:
            bad.append((tuple(shape), "above", hi))
    if clear:
        del _guarded[:]
    return bad


# ---------------------------------------------------------------------------------------------- NVTX ranges (nsys / ncu --nvtx timelines)
class nvtx_range(object):
    """with nvtx_range("gpclust.loop"): ...  -- torch.cuda.nvtx when available, otherwise a no-op."""

    def __init__(self, name):
        self.name, self.on = name, False

    def __enter__(self):
        try:
            import torch
            torch.cuda.nvtx.range_push(self.name)
            self.on = True
        except Exception:
            self.on = False
        return self

    def __exit__(self, *exc):
        if self.on:
            import torch
            torch.cuda.nvtx.range_pop()
        return False
