"""Host mirror of GPclust/MOG.py (class MOG): a collapsed mixture of Gaussians with Gaussian-Wishart
priors.  Same constructor / attributes / methods as the reference (MOG.py:34-113).

The N-sized work of do_computations (MOG.py:56-57: Xsumk, Ck) and vb_grad_natgrad (MOG.py:74-77: the
N x K x D Mahalanobis terms) are the same two GEMMs as MOHGP once every datum is expanded into the
feature row  f(x) = [x_i x_j (i <= j) | x_i]  of the centred data:
    phi^T F           = [Ck (upper triangle) | Xsumk]
    F . W_k + bias_k  = -1/2 vN_k (x - mun_k)^T Sns_inv_k (x - mun_k) + per-cluster constants
so the reference's N x D x D (`XXT`, MOG.py:46) and N x D x D x K (MOG.py:75-76) temporaries never exist.
The data are centred at their mean before the expansion (the posterior is translation invariant);
`mun` is shifted back on read-out.
"""
import numpy as np
from scipy.special import gammaln

from . import _cabi
from ._cabi import lib, check, ptr
from .collapsed_mixture import CollapsedMixture


class MOG(CollapsedMixture):
    def __init__(self, X, K=2, prior_Z="symmetric", alpha=1.0, prior_m=None, prior_kappa=1e-6, prior_S=None, prior_v=None, name="MOG",
                 phi_init=None, group=None):
        torch = _cabi.require_cuda()
        self.N, self.D = X.shape
        D = self.D
        if D > 9:
            raise ValueError("MOG: D=%d needs %d feature columns; the accelerated path supports D <= 9" % (D, D * (D + 3) // 2))

        CollapsedMixture.__init__(self, self.N, K, prior_Z, alpha, name=name, phi_init=phi_init, group=group)

        if isinstance(X, torch.Tensor):
            Xd = X.to(device=self.device, dtype=torch.float64).contiguous()
            self._X_host = None
        else:
            self._X_host = _cabi.f64(X)
            Xd = torch.from_numpy(self._X_host).to(self.device)
        # data mean over all ranks
        s = Xd.sum(0)
        self._allreduce(s)
        mean = (s / float(self.N_total)).cpu().numpy()

        # store the prior cluster parameters (MOG.py:39-42)
        self.m0 = mean.copy() if prior_m is None else np.asarray(prior_m, dtype=np.float64)
        self.k0 = prior_kappa
        self.S0 = np.eye(D) * 1e-3 if prior_S is None else np.asarray(prior_S, dtype=np.float64)
        self.v0 = prior_v or D + 1.0
        self.k0m0m0T = self.k0 * self.m0[:, np.newaxis] * self.m0[np.newaxis, :]
        # MOG.py:47 -- the reference takes the log of the SQUARE ROOT of the Cholesky diagonal; kept.
        self.S0_halflogdet = np.sum(np.log(np.sqrt(np.diag(np.linalg.cholesky(self.S0)))))

        self._centre = mean
        self._centre_dev = torch.from_numpy(mean).to(self.device)
        self._m0c_dev = torch.from_numpy(self.m0 - mean).to(self.device)
        self._S0_dev = torch.from_numpy(_cabi.f64(self.S0)).to(self.device)

        DP = _cabi.padded_d(D * (D + 1) // 2 + D)
        F = torch.empty((self._Npad, DP), dtype=torch.float64, device=self.device)  # fragment-major
        check(lib.gpc_mog_features(ptr(Xd), ptr(self._centre_dev), self.N, D, DP, ptr(F), self._stream()), "gpc_mog_features")
        self.kernel_launches += 1
        self._Xd = Xd
        self._set_features(F, DP)
        self.do_computations()

    @property
    def X(self):
        if self._X_host is None:
            self._X_host = self._Xd.cpu().numpy()
        return self._X_host

    def _alloc_model(self):
        torch = self.torch
        D, K = self.D, self._KP
        z = lambda *shape: _cabi.dzeros(self.device, *shape)
        self._mun_c = z(K, D)
        self._Sns = z(K, D, D)
        self._Sns_inv = z(K, D, D)

    def _launch_cluster(self):
        check(lib.gpc_mog_cluster(ptr(self._stats), ptr(self._m0c_dev), ptr(self._S0_dev), ptr(self._Wt), ptr(self._per_k),
                                  ptr(self._mun_c), ptr(self._Sns), ptr(self._Sns_inv), self._info_ptr(), float(self.k0), float(self.v0),
                                  self.K, self.D, self._KP, self._DP, self._stream()), "gpc_mog_cluster")
        self.kernel_launches += 1

    def _launch_finalize(self):
        check(lib.gpc_mog_finalize(ptr(self._stats), ptr(self._per_k), ptr(self._bias), ptr(self._ctrl), ptr(self._mixgrad),
                                   float(self.N_total), float(self.k0), float(self.v0), float(self.S0_halflogdet), float(self.alpha),
                                   _cabi.PRIOR[self.prior_Z], self.K, self._KP, self._DP, self.D, self._stream()), "gpc_mog_finalize")
        self.kernel_launches += 1

    def _loop_posterior(self, phase):
        check(lib.gpc_reduce_partials(ptr(self._spart), self._nparts, self._stats_len, ptr(self._stats), self._stream()), "gpc_reduce_partials")
        self._allreduce(self._stats)
        check(lib.gpc_loop_mog_post(ptr(self._stats), ptr(self._m0c_dev), ptr(self._S0_dev), ptr(self._Wt), ptr(self._per_k), ptr(self._mun_c),
                                    ptr(self._Sns), ptr(self._Sns_inv), self._info_ptr(), float(self.k0), float(self.v0), ptr(self._bias),
                                    ptr(self._ctrl), ptr(self._mixgrad), float(self.N_total), float(self.S0_halflogdet), float(self.alpha),
                                    _cabi.PRIOR[self.prior_Z], self.K, self.D, self._KP, self._DP, ptr(self._ctrl[3:]), ptr(self._ctrl[6:]),
                                    ptr(self._loop_dev), phase, self._stream()), "gpc_loop_mog_post")
        self.kernel_launches += 3

    # ---- reference attributes (MOG.py:54-61), K LAST as in the reference
    @property
    def kNs(self):
        self._ensure_state()
        return self._per_k[:self.K, 0].cpu().numpy()

    @property
    def vNs(self):
        self._ensure_state()
        return self._per_k[:self.K, 1].cpu().numpy()

    @property
    def Sns_halflogdet(self):
        self._ensure_state()
        return self._per_k[:self.K, 2].cpu().numpy()

    @property
    def mun(self):  # D x K
        self._ensure_state()
        return (self._mun_c[:self.K].cpu().numpy() + self._centre[None, :]).T.copy()

    @property
    def Xsumk(self):  # D x K
        self._ensure_state()
        NQ = self.D * (self.D + 1) // 2
        T = self._stats[:self._KP * self._DP].view(self._KP, self._DP)[:self.K].cpu().numpy()
        return (T[:, NQ:NQ + self.D] + self.phi_hat[:, None] * self._centre[None, :]).T.copy()

    @property
    def Sns(self):  # D x D x K
        self._ensure_state()
        return np.ascontiguousarray(self._Sns[:self.K].cpu().numpy().transpose(1, 2, 0))

    @property
    def Sns_inv(self):  # D x D x K
        self._ensure_state()
        return np.ascontiguousarray(self._Sns_inv[:self.K].cpu().numpy().transpose(1, 2, 0))

    # ---- prediction (MOG.py:87-113): K x D x D host algebra on the device-computed posterior
    def predict_components_ln(self, Xnew):
        """The log predictive density under each component at Xnew"""
        Xnew = np.asarray(Xnew, dtype=np.float64)
        mun, Sns_inv, kNs, vNs, hld = self.mun, self.Sns_inv, self.kNs, self.vNs, self.Sns_halflogdet
        Dist = Xnew[:, :, np.newaxis] - mun[np.newaxis, :, :]
        tmp = np.sum(Dist[:, :, None, :] * Sns_inv[None, :, :, :], 1)
        mahalanobis = np.sum(tmp * Dist, 1) / (kNs + 1.0) * kNs * (vNs - self.D + 1.0)
        halflndetSigma = hld + 0.5 * self.D * np.log((kNs + 1.0) / (kNs * (vNs - self.D + 1.0)))
        v = vNs[np.newaxis, :]
        Z = (gammaln(0.5 * (v + 1.0)) - gammaln(0.5 * (v - self.D + 1.0))
             - (0.5 * self.D) * (np.log(v - self.D + 1.0) + np.log(np.pi))
             - halflndetSigma
             - (0.5 * (v + 1.0)) * np.log(1.0 + mahalanobis / (v - self.D + 1.0)))
        return Z

    def predict_components(self, Xnew):
        """The predictive density under each component at Xnew"""
        return np.exp(self.predict_components_ln(Xnew))

    def predict(self, Xnew):
        """The predictive density of the model at Xnew"""
        Z = self.predict_components(Xnew)
        pi = self.phi_hat + self.alpha
        pi /= pi.sum()
        Z *= pi[np.newaxis, :]
        return Z.sum(1)
