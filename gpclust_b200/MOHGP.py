"""Host mirror of GPclust/MOHGP.py (class MOHGP): a hierarchical mixture of Gaussian processes.

Same constructor, attributes and methods as the reference; the arithmetic of do_computations /
bound / vb_grad_natgrad (MOHGP.py:70-90, :124-135, :137-153) runs in the sm_100a kernels:

    S pass (gpc_step_stats)      softmax, phi_hat, H, ybark = phi^T Y                MOHGP.py:76
    gpc_mohgp_post               per-cluster posterior in the eigenbasis of A = Ly^-1 Sf Ly^-T (no per-cluster
                                 factorisation), bound, G-pass weights                MOHGP.py:79-90, :124-135
    G pass (gpc_natgrad)         Mahalanobis terms as one GEMM, natural gradient      MOHGP.py:137-153
    gpc_mohgp_cluster            the reference's literal Cholesky / TRSM / SYRK per cluster -- only when the
                                 K x D x D attributes (Lambda_inv, _C_chols) are read MOHGP.py:79-85

Y may be a numpy array or a CUDA tensor (N x D; N = the rows this rank holds).  The dead N x D x D
`YYT` of MOHGP.py:52 is not materialised.
"""
import os

import numpy as np

from . import _cabi
from ._cabi import lib, check, ptr
from .collapsed_mixture import CollapsedMixture
from .kern import cov_to_device


class MOHGP(CollapsedMixture):
    def __init__(self, X, kernF, kernY, Y, K=2, alpha=1.0, prior_Z="symmetric", name="MOHGP", phi_init=None, group=None):
        torch = _cabi.require_cuda()
        N, self.D = Y.shape
        self.X = np.asarray(X, dtype=np.float64)
        assert self.X.shape[0] == self.D, "input data don't match observations"

        self.kernF = kernF
        self.kernY = kernY
        self._early_setup = None
        CollapsedMixture.__init__(self, N, K, prior_Z, alpha, name, phi_init=phi_init, group=group, before_upload=self._kernel_setup_early)

        # features: F = Y, fragment-major, Npad x DP
        DP = _cabi.padded_d(self.D)
        F = _cabi.dzeros(self.device, self._Npad, DP) if _cabi.GUARD else torch.empty((self._Npad, DP), dtype=torch.float64, device=self.device)  # fragment-major (include/gpclust_b200.h "Layout")
        if isinstance(Y, torch.Tensor):
            Yd = Y.to(device=self.device, dtype=torch.float64).contiguous()
            self._Y_host = None
            check(lib.gpc_pack_rows(ptr(Yd), N, self.D, DP, _cabi.LAYOUT["features"], ptr(F), self._stream()), "gpc_pack_rows")
            self.kernel_launches += 1
            del Yd
        else:
            self._Y_host = _cabi.f64(Y)
            self.kernel_launches += _cabi.upload_pack(self._Y_host, self.D, DP, _cabi.LAYOUT["features"], F, self.device, self._stream())
        self._set_features(F, DP)

        # Computations that can be done outside the optimisation loop (MOHGP.py:52-53)
        self._compute_YTY()
        # initialize kernels (MOHGP.py:47-49) and the first do_computations (MOHGP.py:55)
        self.parameters_changed()

    def _kernel_setup_early(self):
        """The covariance fills, pdinv(Sy + 1e-6 I), A and its eigendecomposition (MOHGP.py:47-49) depend on X and the kernels
        only: they are enqueued on a side stream BEFORE the rows are uploaded, so the 4 ms Jacobi kernel runs while Y and the
        initial logits are still crossing PCIe.  parameters_changed() joins the side stream."""
        torch = self.torch
        self._X_dev = torch.from_numpy(self.X).to(self.device)
        D = self.D
        z = lambda *shape: _cabi.dzeros(self.device, *shape)
        self._Sf, self._Sy = z(D, D), z(D, D)
        self._Ly, self._Ly_inv, self._Sy_inv, self._A = z(D, D), z(D, D), z(D, D), z(D, D)
        self._Sy_logdet, self._const = z(1), z(1)
        self._YTY = z(D, D)
        self._eig_ws = z(int(lib.gpc_mohgp_eig_ws_len(D)))
        if os.environ.get("GPCLUST_B200_EARLY", "1") == "0":
            return  # parameters_changed() launches the set-up in stream order
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._launch_kernel_setup()
        self._early_setup = side

    def _launch_kernel_setup(self):
        D = self.D
        self._Sf.copy_(cov_to_device(self.kernF, self.X, self._X_dev))
        self._Sy.copy_(cov_to_device(self.kernY, self.X, self._X_dev))
        check(lib.gpc_mohgp_setup(ptr(self._Sf), ptr(self._Sy), ptr(self._Ly), ptr(self._Ly_inv), ptr(self._Sy_inv), ptr(self._A),
                                  ptr(self._Sy_logdet), self._info_ptr(), D, self._stream()), "gpc_mohgp_setup")
        check(lib.gpc_mohgp_eig(ptr(self._A), ptr(self._Ly), ptr(self._Ly_inv), ptr(self._Sy_inv), ptr(self._Sf), D, ptr(self._eig_ws),
                                self._stream()), "gpc_mohgp_eig")
        self.kernel_launches += 4

    @property
    def Y(self):
        if self._Y_host is None:
            out = self.torch.empty((self.N, self.D), dtype=self.torch.float64, device=self.device)
            check(lib.gpc_unpack_rows(ptr(self._F), self.N, self.D, self._DP, _cabi.LAYOUT["features"], ptr(out), self._stream()), "gpc_unpack_rows")
            self._Y_host = _cabi.to_host(out)
        return self._Y_host

    def _alloc_model(self):
        torch = self.torch
        D, K = self.D, self._KP
        z = lambda *shape: _cabi.dzeros(self.device, *shape)
        self._muk_t = z(K, D)
        self._Syi_ybark_t = z(K, D)
        self._Lambda_inv = None  # K x D x D attributes: allocated and filled on first read (_dense_posterior)
        self._C_chols_dev = None

    def _compute_YTY(self):
        torch = self.torch
        DP = self._DP
        grid = min(self._grid, self._Npad // 32)
        part = torch.zeros(grid * DP * DP, dtype=torch.float64, device=self.device)
        full = torch.zeros(DP * DP, dtype=torch.float64, device=self.device)
        check(lib.gpc_gram_partials(ptr(self._F), self._Npad, DP, ptr(part), grid, self._stream()), "gpc_gram_partials")
        check(lib.gpc_reduce_partials(ptr(part), grid, DP * DP, ptr(full), self._stream()), "gpc_reduce_partials")
        self.kernel_launches += 2
        self._allreduce(full)
        self._YTY.copy_(full.view(DP, DP)[:self.D, :self.D])

    def parameters_changed(self):
        """MOHGP.py:58-67: refresh the kernel matrices and their factorisation, then recompute (and, when kernel parameters
        are free, their gradients: update_kern_grads, MOHGP.py:92-122)."""
        D = self.D
        if self._early_setup is not None:  # constructor: already enqueued on the side stream (see _kernel_setup_early)
            self.torch.cuda.current_stream().wait_stream(self._early_setup)
            self._early_setup = None
        else:
            self._launch_kernel_setup()
        check(lib.gpc_mohgp_const(ptr(self._YTY), ptr(self._Sy_inv), ptr(self._Sy_logdet), float(self.N_total), D, ptr(self._const),
                                  self._stream()), "gpc_mohgp_const")
        self.kernel_launches += 1
        _, _, _, failed = self._readback()
        if failed:
            raise np.linalg.LinAlgError("not positive definite, even with jitter.")
        self.do_computations()
        self._kern_grads_valid = False

    def _kernel_param_list(self):
        out = []
        for k in (self.kernF, self.kernY):  # link order of MOHGP.py:44
            if hasattr(k, "param_list"):
                out.extend(k.param_list())
        return out

    def _kernel_gradients(self):
        if not getattr(self, "_kern_grads_valid", False):
            self.update_kern_grads()
        return CollapsedMixture._kernel_gradients(self)

    def update_kern_grads(self):
        """MOHGP.py:92-122: derivative of the bound w.r.t. the kernel parameters through dL/dSf and dL/dSy.  The two D x D matrices
        come from the device (gpc_mohgp_kern_grads: eigenbasis form, O(K D^2) instead of K triangular solves and K D^3 products on
        the host); the contraction with dK/dtheta (GPy update_gradients_full) is D x D host algebra."""
        self._ensure_state()
        torch = self.torch
        D = self.D
        work = torch.empty(int(lib.gpc_mohgp_kern_grads_work(self.K, D)), dtype=torch.float64, device=self.device)
        out = torch.empty((2, D, D), dtype=torch.float64, device=self.device)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        with _cabi.nvtx_range("gpclust.kern_grads"):
            check(lib.gpc_mohgp_kern_grads(ptr(self._stats), ptr(self._eig_ws), ptr(self._Sf), ptr(self._Sy_inv), ptr(self._YTY), ptr(self._Syi_ybark_t),
                                           float(self.N_total), self.K, D, self._KP, self._DP, ptr(work), ptr(out[0]), ptr(out[1]), ptr(info),
                                           self._stream()), "gpc_mohgp_kern_grads")
        self.kernel_launches += 2
        host = out.cpu().numpy()
        if int(info.item()):
            raise np.linalg.LinAlgError("not positive definite, even with jitter.")
        dL_dSf, dL_dSy = host[0], host[1]
        self.kernF.update_gradients_full(dL_dK=dL_dSf, X=self.X)
        self.kernY.update_gradients_full(dL_dK=dL_dSy, X=self.X)
        self._dL_dSf, self._dL_dSy = dL_dSf, dL_dSy
        self._kern_grads_valid = True

    def _launch_posterior(self):
        """Statistics -> posterior -> bound in one launch (two + an all-reduce when rows are sharded)."""
        if self._world > 1:
            check(lib.gpc_reduce_partials(ptr(self._spart), self._nparts, self._stats_len, ptr(self._stats), self._stream()), "gpc_reduce_partials")
            self.kernel_launches += 1
            self._allreduce(self._stats)
            parts, nparts = self._stats, 1
        else:
            parts, nparts = self._spart, self._nparts
        check(lib.gpc_mohgp_post(ptr(parts), nparts, ptr(self._stats), ptr(self._eig_ws), ptr(self._Sy_inv), ptr(self._Sf), ptr(self._const),
                                 ptr(self._Wt), ptr(self._muk_t), ptr(self._Syi_ybark_t), ptr(self._per_k), ptr(self._bias), ptr(self._ctrl),
                                 ptr(self._mixgrad), ptr(self._counter[1:]), self._info_ptr(), float(self.alpha), _cabi.PRIOR[self.prior_Z],
                                 self.K, self.D, self._KP, self._DP, self._stream()), "gpc_mohgp_post")
        self.kernel_launches += 1

    def _use_peer_exchange(self):
        return True

    def _fused_ok(self):
        # K <= 64 (one 64-cluster chunk per CTA), every SM's CTA co-resident, the NVLink exchange (or one GPU).  GPCLUST_B200_FUSED=0
        # keeps the per-evaluation launch sequence (graph replay) for comparison.
        return (self._chunks == 1 and self._grid > self._KP and (self._world == 1 or self._peers is not None) and
                os.environ.get("GPCLUST_B200_FUSED", "1") != "0" and self._kernel_events is None)

    def _launch_fused(self, bufs, n):
        import ctypes
        px = ptr(self._peers["desc"]) if self._peers is not None else None
        check(lib.gpc_mohgp_loop(ptr(self._F), ctypes.byref(bufs), ptr(self._Wt), ptr(self._bias), ptr(self._gpart), ptr(self._spart),
                                 ptr(self._ctrl[3:]), ptr(self._ctrl[6:]), ptr(self._stats), ptr(self._eig_ws), ptr(self._Sy_inv), ptr(self._Sf),
                                 ptr(self._const), ptr(self._muk_t), ptr(self._Syi_ybark_t), ptr(self._per_k), ptr(self._ctrl), ptr(self._mixgrad),
                                 self._info_ptr(), float(self.alpha), _cabi.PRIOR[self.prior_Z], ptr(self._loop_dev), px, ptr(self._sync_words),
                                 ptr(self._loop_timing), int(n), self.N, self.K, self.D, self._KP, self._DP, self._grid, self._stream()),
              "gpc_mohgp_loop")
        self.kernel_launches += 1

    def _loop_posterior(self, phase):
        if self._peers is not None:
            # rows sharded over the GPUs of one box: the reduced statistics go straight into every peer's inbox over
            # NVLink (no NCCL call, no host involvement); the posterior kernel sums the W inbox vectors in rank order
            px = ptr(self._peers["desc"])
            check(lib.gpc_reduce_push(ptr(self._spart), self._nparts, self._stats_len, px, ptr(self._counter[2:]), ptr(self._loop_dev), phase,
                                      self._stream()), "gpc_reduce_push")
            check(lib.gpc_loop_mohgp_post_px(px, self._world, ptr(self._stats), ptr(self._eig_ws), ptr(self._Sy_inv), ptr(self._Sf), ptr(self._const),
                                             ptr(self._Wt), ptr(self._muk_t), ptr(self._Syi_ybark_t), ptr(self._per_k), ptr(self._bias),
                                             ptr(self._ctrl), ptr(self._mixgrad), ptr(self._counter[1:]), self._info_ptr(), float(self.alpha),
                                             _cabi.PRIOR[self.prior_Z], self.K, self.D, self._KP, self._DP, ptr(self._ctrl[3:]), ptr(self._ctrl[6:]),
                                             ptr(self._loop_dev), phase, self._stream()), "gpc_loop_mohgp_post_px")
            self.kernel_launches += 2
            return
        if self._world > 1:
            check(lib.gpc_reduce_partials(ptr(self._spart), self._nparts, self._stats_len, ptr(self._stats), self._stream()), "gpc_reduce_partials")
            self.kernel_launches += 1
            self._allreduce(self._stats)
            parts, nparts = self._stats, 1
        else:
            parts, nparts = self._spart, self._nparts
        check(lib.gpc_loop_mohgp_post(ptr(parts), nparts, ptr(self._stats), ptr(self._eig_ws), ptr(self._Sy_inv), ptr(self._Sf), ptr(self._const),
                                      ptr(self._Wt), ptr(self._muk_t), ptr(self._Syi_ybark_t), ptr(self._per_k), ptr(self._bias), ptr(self._ctrl),
                                      ptr(self._mixgrad), ptr(self._counter[1:]), self._info_ptr(), float(self.alpha), _cabi.PRIOR[self.prior_Z],
                                      self.K, self.D, self._KP, self._DP, ptr(self._ctrl[3:]), ptr(self._ctrl[6:]), ptr(self._loop_dev), phase,
                                      self._stream()), "gpc_loop_mohgp_post")
        self.kernel_launches += 1

    def _dense_posterior(self):
        """The reference's literal per-cluster route (MOHGP.py:79-85) for the K x D x D attributes; results of the
        hot-loop route (Wt, muk, per-cluster scalars) are left untouched."""
        self._ensure_state()
        if "dense" not in self._cache:
            torch = self.torch
            D, KP = self.D, self._KP
            z = lambda *shape: _cabi.dzeros(self.device, *shape)
            self._Lambda_inv, self._C_chols_dev = z(KP, D, D), z(KP, D, D)
            scratch = {"Wt": z(KP, self._DP), "muk_t": z(KP, D), "per_k": z(KP, 4), "s": z(KP, D), "info": torch.zeros(1, dtype=torch.int32, device=self.device)}
            check(lib.gpc_mohgp_cluster(ptr(self._stats), ptr(self._A), ptr(self._Ly), ptr(self._Ly_inv), ptr(self._Sy), ptr(self._Sy_inv), ptr(self._Sf),
                                        ptr(scratch["Wt"]), ptr(scratch["muk_t"]), ptr(scratch["per_k"]), ptr(scratch["s"]), ptr(self._Lambda_inv),
                                        ptr(self._C_chols_dev), ptr(scratch["info"]), self.K, self.D, self._KP, self._DP, self._stream()),
                  "gpc_mohgp_cluster")
            self.kernel_launches += 1
            if int(scratch["info"].item()):
                raise np.linalg.LinAlgError("not positive definite, even with jitter.")
            self._cache["dense"] = scratch
        return self._cache["dense"]

    # ---- reference attributes (copied from the device on access)
    def _np(self, t):
        return t.cpu().numpy()

    Sf = property(lambda self: self._np(self._Sf))
    Sy = property(lambda self: self._np(self._Sy))
    Sy_inv = property(lambda self: self._np(self._Sy_inv))
    Sy_chol = property(lambda self: self._np(self._Ly))
    Sy_chol_inv = property(lambda self: self._np(self._Ly_inv))
    Sy_logdet = property(lambda self: float(self._Sy_logdet.cpu()[0]))
    YTY = property(lambda self: self._np(self._YTY))

    @property
    def ybark(self):  # D x K (MOHGP.py:76)
        self._ensure_state()
        T = self._stats[:self._KP * self._DP].view(self._KP, self._DP)
        return self._np(T[:self.K, :self.D]).T.copy()

    @property
    def muk(self):  # D x K (MOHGP.py:90)
        self._ensure_state()
        return self._np(self._muk_t[:self.K]).T.copy()

    @property
    def Syi_ybark(self):  # D x K (MOHGP.py:88)
        self._ensure_state()
        return self._np(self._Syi_ybark_t[:self.K]).T.copy()

    @property
    def Lambda_inv(self):  # K x D x D (MOHGP.py:85)
        self._dense_posterior()
        return self._np(self._Lambda_inv[:self.K])

    @property
    def _C_chols(self):  # list of K lower factors (MOHGP.py:82)
        self._dense_posterior()
        return list(self._np(self._C_chols_dev[:self.K]))

    @property
    def log_det_diff(self):  # MOHGP.py:83
        self._ensure_state()
        return self._np(self._per_k[:self.K, 0])

    @property
    def Syi_ybarkybarkT_Syi(self):  # K x D x D (MOHGP.py:89)
        s = self.Syi_ybark.T
        return s[:, None, :] * s[:, :, None]

    def predict_components(self, Xnew):
        """The predictive density under each component (MOHGP.py:156-174); small host algebra on the
        device-computed posterior (K x D x D), not part of the VB iteration."""
        Xnew = np.asarray(Xnew, dtype=np.float64)
        Sy_chol_inv, Sf, Sy_inv = self.Sy_chol_inv, self.Sf, self.Sy_inv
        from scipy.linalg import solve_triangular
        tmp = [solve_triangular(L, Sy_chol_inv, lower=True) for L in self._C_chols]
        B_invs = [ph * np.dot(t.T, t) for ph, t in zip(self.phi_hat, tmp)]
        kx = self.kernF.K(self.X, Xnew)
        kxx = self.kernF.K(Xnew) + self.kernY.K(Xnew)
        tmp = [np.eye(self.D) - np.dot(Bi, Sf) for Bi in B_invs]
        mu = [kx.T.dot(t).dot(Sy_inv).dot(yb) for t, yb in zip(tmp, self.ybark.T)]
        var = [kxx - kx.T.dot(Bi).dot(kx) for Bi in B_invs]
        return mu, var
