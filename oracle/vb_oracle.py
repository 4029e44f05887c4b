"""TEST INFRASTRUCTURE ONLY -- CPU restatement of GPclust's collapsed-VB hot path (numpy/scipy, FP64).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module; the product package (gpclust_b200/) never does and has no CPU fallback.

Every routine cites the reference lines it restates (paths relative to /root/reference/GPclust/).
The reference itself cannot be imported here: each module imports GPy (absent, un-installable) and
`scipy.weave` does not exist on Python 3, so what would actually run is the `np_utilities.py`
fallbacks -- those are what is restated, quirks included.

Parity pins:
  * MOG + CollapsedMixture + CollapsedVB.optimize + try_split/systematic_splits: PINNED by the seeded
    notebook trace `tests/golden/mog_notebook_golden.json` (tests/test_oracle_golden.py): identical
    iteration numbering over 61 + 52 + 55 printed lines, bounds to 4e-12 relative (the notebook prints 12
    significant digits), K = 4 -> 5 -> 11 through the splits, posterior mean/covariance and predictive
    densities to the printed digits.
  * MOHGP, OMGP (and MOG again): PINNED against outputs of the reference's own modules -- /root/reference/GPclust
    imported unmodified in the build container with a stub `GPy` (tests/golden/make_reference_goldens.py) --
    stored in tests/golden/reference_*.npz and checked by tests/test_reference_goldens.py: start bound, gradient,
    natural gradient, posterior means, full optimisation trace (iteration numbering, bounds, betas), final phi,
    MOHGP.predict_components, OMGP.update_kern_grads, and the merge-split search on MOHGP (reference_splits_*.npz:
    attempts, decisions, bound increments, final K / assignments, state of the legacy numpy stream).
  * Oracle-independent identities (tests/test_oracle_identities.py): gradient = derivative of the bound, product-of-
    Gaussians posterior, equality of the Mahalanobis / softmax forms incl. the C port, monotone bound.
    What remains the shim's (not the reference's) is the GPy layer itself (thin LAPACK wrappers, closed-form
    covariances): GPy is a third-party dependency absent from /root/reference (`GPy>=0.6`, setup.py:14).
"""
import sys

import numpy as np
from scipy.special import digamma, gammaln

from . import gpy_shim as G

LinAlgError = np.linalg.LinAlgError


# --------------------------------------------------------------------------- L1 utilities
def softmax_rows(X):
    """np_utilities.py:25-40 (`softmax_numpy`): row softmax, log-softmax and total entropy."""
    shifted = X - X.max(1)[:, None]
    phi = np.exp(shifted)
    norm = phi.sum(1)
    phi /= norm[:, None]
    logphi = shifted - np.log(norm)[:, None]
    H = (-(phi * logphi).sum(1)).sum()
    return phi, logphi, H


def mahalanobis_pairs_solve(X1, X2, L):
    """np_utilities.py:43-59 (`multiple_mahalanobis_numpy_loops`): per pair, solve with L L^T.
    Python double loop -- small inputs only."""
    S = L.dot(L.T)
    out = np.zeros((X1.shape[0], X2.shape[0]))
    for n in range(X1.shape[0]):
        for m in range(X2.shape[0]):
            d = X1[n] - X2[m]
            out[n, m] = d.dot(np.linalg.solve(S, d))
    return out


def mahalanobis_pairs_substitution(X1, X2, L):
    """utilities.py:147-170 (weave `multiple_mahalanobis`): forward-substitute L t = x1_n - x2_m for
    every pair, result = t.t.  Vectorised over pairs, same per-pair operation order in i, j."""
    N1, D = X1.shape
    diff = X1[:, None, :] - X2[None, :, :]
    t = np.zeros_like(diff)
    for i in range(D):
        acc = diff[:, :, i].copy()
        for j in range(i):
            acc -= L[i, j] * t[:, :, j]
        t[:, :, i] = acc / L[i, i]
    out = np.zeros(diff.shape[:2])
    for i in range(D):
        out += t[:, :, i] * t[:, :, i]
    return out


def mahalanobis_whitened_gemm(X1, X2, L):
    """Same quantity as a whitened GEMM (|z1|^2 - 2 z1.z2 + |z2|^2); NOT the reference's operation
    order -- used for large-N CPU baselines ("CPU-vectorised" line of BASELINE.md section 3)."""
    Z1 = G.dtrtrs(L, X1.T, lower=1)[0].T
    Z2 = G.dtrtrs(L, X2.T, lower=1)[0].T
    return (Z1 * Z1).sum(1)[:, None] - 2.0 * Z1.dot(Z2.T) + (Z2 * Z2).sum(1)[None, :]


def multiple_pdinv(A):
    """np_utilities.py:6-22: inverses and half log-dets of the D x D x K stack A[:, :, k]."""
    chols = [G.jitchol(A[:, :, k]) for k in range(A.shape[-1])]
    hld = np.array([np.sum(np.log(np.diag(L))) for L in chols])
    invs = [G.dpotri(L, True)[0] for L in chols]
    invs = [np.triu(I) + np.triu(I, 1).T for I in invs]
    return np.dstack(invs), hld


def lngammad(v, D):
    """np_utilities.py:62-64."""
    return np.sum([gammaln((v + 1.0 - d) / 2.0) for d in range(1, D + 1)], 0)


def ln_dirichlet_C(a):
    """np_utilities.py:67-69."""
    return gammaln(a.sum()) - np.sum(gammaln(a))


# --------------------------------------------------------------------------- L2 + L3
class MixtureOracle(object):
    """collapsed_vb.py:12-323 (CollapsedVB) + collapsed_mixture.py:15-205 (CollapsedMixture).

    Kernel hyper-parameter optimisation (collapsed_vb.py:313-323 -> GPy Model.optimize) is outside
    the v0 scope: `optimize_parameters()` returns 0. as the reference does for a model with no free
    parameters (collapsed_vb.py:322-323); MOG has none, so its trace is unaffected.
    """

    mahalanobis = staticmethod(mahalanobis_pairs_substitution)
    softmax = staticmethod(softmax_rows)  # swap in oracle.native.softmax_weave for the "reference native" CPU baseline

    def __init__(self, N, K, prior_Z="symmetric", alpha=1.0):
        # collapsed_mixture.py:24-47
        assert prior_Z in ("symmetric", "DP")
        self.N, self.K, self.prior_Z, self.alpha = N, K, prior_Z, alpha
        self.hyperparam_interval = 50  # collapsed_vb.py:29
        self.trace = []
        self.out = sys.stdout
        self.phi_ = np.random.randn(self.N, self.K)  # legacy global stream, collapsed_mixture.py:41
        self._softmax_state()

    def _softmax_state(self):
        # collapsed_mixture.py:42-47 and :54-59
        self.phi, logphi, self.H = self.softmax(self.phi_)
        self.phi_hat = self.phi.sum(0)
        self.Hgrad = -logphi
        if self.prior_Z == "DP":
            self.phi_tilde_plus_hat = self.phi_hat[::-1].cumsum()[::-1]
            self.phi_tilde = self.phi_tilde_plus_hat - self.phi_hat

    def set_vb_param(self, x):
        # collapsed_mixture.py:49-61
        self.phi_ = np.asarray(x, dtype=np.float64).reshape(self.N, self.K)
        self._softmax_state()
        self.do_computations()

    def get_vb_param(self):
        return self.phi_.flatten()  # collapsed_mixture.py:63-64

    def randomize(self):
        self.set_vb_param(np.random.randn(self.N * self.K))  # collapsed_vb.py:37-39

    def log_likelihood(self):
        return self.bound()  # collapsed_vb.py:56-61

    def optimize_parameters(self):
        return 0.0  # collapsed_vb.py:318-323 with an empty optimizer_array

    # -- mixing proportions: collapsed_mixture.py:66-94
    def mixing_prop_bound(self):
        if self.prior_Z == "symmetric":
            return ln_dirichlet_C(np.ones(self.K) * self.alpha) - ln_dirichlet_C(self.alpha + self.phi_hat)
        A = gammaln(1.0 + self.phi_hat)
        B = gammaln(self.alpha + self.phi_tilde)
        C = gammaln(self.alpha + 1.0 + self.phi_tilde_plus_hat)
        D = self.K * (gammaln(1.0 + self.alpha) - gammaln(self.alpha))
        return A.sum() + B.sum() - C.sum() + D

    def mixing_prop_bound_grad(self):
        if self.prior_Z == "symmetric":
            return digamma(self.alpha + self.phi_hat)
        A = digamma(self.phi_hat + 1.0)
        B = np.hstack((0, digamma(self.phi_tilde + self.alpha)[:-1].cumsum()))
        C = digamma(self.phi_tilde_plus_hat + self.alpha + 1.0).cumsum()
        return A + B - C

    def _natural(self, grad_phi):
        # MOHGP.py:150-153 / MOG.py:81-84 / OMGP.py:144-147
        natgrad = grad_phi - np.sum(self.phi * grad_phi, 1)[:, None]
        grad = natgrad * self.phi
        return grad.flatten(), natgrad.flatten()

    def _loop_pass(self, st, method, step_length):
        """One pass of the loop body collapsed_vb.py:152-193.  `st` carries bound_old, iteration and the
        previous iteration's natgrad / grad / searchDir / squareNorm."""
        grad, natgrad = self.vb_grad_natgrad()
        grad, natgrad = -grad, -natgrad
        squareNorm = np.dot(natgrad, grad)
        if method == "steepest" or not st["iteration"]:
            beta = 0
        elif method == "PR":
            beta = np.dot((natgrad - st["natgrad_old"]), grad) / st["squareNorm_old"]
        elif method == "FR":
            beta = squareNorm / st["squareNorm_old"]
        else:
            beta = np.dot((natgrad - st["natgrad_old"]), grad) / np.dot((natgrad - st["natgrad_old"]), st["grad_old"])
        if np.isnan(beta) or (beta < 0.0):
            beta = 0.0
        searchDir = -natgrad + beta * st["searchDir_old"]

        phi_old = self.get_vb_param().copy()
        try:
            self.set_vb_param(phi_old + step_length * searchDir)
            bound = self.bound()
        except LinAlgError:
            self.set_vb_param(phi_old)
            bound = st["bound_old"] - 1
        st["iteration"] += 1

        if bound < st["bound_old"]:  # revert to steepest (== VBEM), collapsed_vb.py:182-193
            searchDir = -natgrad
            try:
                self.set_vb_param(phi_old + step_length * searchDir)
                bound = self.bound()
            except LinAlgError:
                self.set_vb_param(phi_old)
                bound = self.bound()
            st["iteration"] += 1
            st["rejected"] = st.get("rejected", 0) + 1
        return bound, squareNorm, beta, grad, natgrad, searchDir

    @staticmethod
    def _carry(st, bound, squareNorm, grad, natgrad, searchDir):
        # collapsed_vb.py:294-303
        st["natgrad_old"] = natgrad.copy()
        st["grad_old"] = grad.copy()
        st["searchDir_old"] = searchDir.copy()
        st["squareNorm_old"] = squareNorm
        st["bound_old"] = bound

    def iterate(self, n_passes, method="HS", step_length=1.0):
        """Exactly n_passes passes of the loop body, no termination tests (bench.py's CPU baseline)."""
        st = {"iteration": 0, "bound_old": self.bound(), "searchDir_old": 0.0}
        for _ in range(n_passes):
            bound, squareNorm, beta, grad, natgrad, searchDir = self._loop_pass(st, method, step_length)
            self._carry(st, bound, squareNorm, grad, natgrad, searchDir)
        return st

    # -- the conjugate natural-gradient loop: collapsed_vb.py:63-310 (numerics :143-193, :228-303)
    def optimize(self, method="HS", maxiter=500, ftol=1e-6, gtol=1e-6, step_length=1.0, callback=None, verbose=True):
        assert method in ("FR", "PR", "HS", "steepest"), "invalid conjugate gradient method specified."
        st = {"iteration": 0, "bound_old": self.bound(), "searchDir_old": 0.0}
        while True:
            if callback is not None:
                callback()
            bound, squareNorm, beta, grad, natgrad, searchDir = self._loop_pass(st, method, step_length)
            iteration, bound_old = st["iteration"], st["bound_old"]

            self.trace.append((iteration, float(bound), float(squareNorm), float(beta)))
            if verbose:
                self.out.write("iteration %d bound=%.12g grad=%.12g, beta=%.12g\n" % (iteration, bound, squareNorm, beta))

            if np.abs(bound - bound_old) <= ftol:
                if verbose:
                    self.out.write("vb converged (ftol)\n")
                if self.optimize_parameters() < 1e-1:
                    break
            if squareNorm <= gtol:
                if verbose:
                    self.out.write("vb converged (gtol)\n")
                if self.optimize_parameters() < 1e-1:
                    break
            if iteration >= maxiter:
                if verbose:
                    self.out.write("maxiter exceeded\n")
                break

            if (iteration > 1) and not (iteration % self.hyperparam_interval):
                self.optimize_parameters()
            self._carry(st, bound, squareNorm, grad, natgrad, searchDir)

    # -- merge/split search: collapsed_mixture.py:96-205
    def reorder(self):
        if self.prior_Z == "DP":
            i = np.argsort(self.phi_hat)[::-1]
            self.set_vb_param(self.phi_[:, i])

    def remove_empty_clusters(self, threshold=1e-6):
        i = self.phi_hat > threshold
        phi_ = self.phi_[:, i]
        self.K = int(i.sum())
        self.set_vb_param(phi_)

    def try_split(self, indexK=None, threshold=0.9, verbose=True, maxiter=100, optimize_params=None):
        if indexK is None:
            indexK = np.random.multinomial(1, self.phi_hat / self.N).argmax()
        if indexK > (self.K - 1):
            return False
        elif self.phi_hat[indexK] < 1:
            return False
        if np.sum(self.phi[:, indexK] > threshold) < 2:
            return False
        if verbose:
            self.out.write("\nattempting to split cluster  %d\n" % indexK)

        bound_old = self.bound()
        phi_old = self.get_vb_param().copy()
        hyper = getattr(self, "hyper_free", False)  # collapsed_mixture.py:143-144: param_old = self.optimizer_array.copy()
        param_old = [getattr(o, n) for o, n in self.kernel_param_list()] if hyper else []
        old_K = self.K

        self.K += 1
        self.phi_ = np.hstack((self.phi_, self.phi_.min(1)[:, None]))
        indexN = np.nonzero(self.phi[:, indexK] > threshold)[0]  # pre-split phi (appendix A.8)
        special = np.random.permutation(indexN)[0]
        self.phi_[indexN, -1] = self.phi_[indexN, indexK].copy()
        self.phi_[special, -1] = np.max(self.phi_[special]) + 10
        self.set_vb_param(self.get_vb_param())

        self.optimize(maxiter=maxiter, verbose=verbose)
        self.remove_empty_clusters()
        bound_new = self.bound()
        bound_increase = bound_new - bound_old
        if bound_increase < 1e-3:
            self.K = old_K
            self.set_vb_param(phi_old)
            if hyper:  # collapsed_mixture.py:174: self.optimizer_array = param_old
                for (o, n), v in zip(self.kernel_param_list(), param_old):
                    setattr(o, n, v)
                self.parameters_changed()
            if verbose:
                self.out.write("split failed, bound changed by:  %.12g (K=%s)\n" % (bound_increase, self.K))
            return False
        if verbose:
            self.out.write("split suceeded, bound changed by:  %.12g , %d  new clusters (K=%s)\n" % (bound_increase, self.K - old_K, self.K))
            self.out.write("optimizing new split to convergence:\n")
        self.last_split_increase = bound_increase
        if optimize_params:
            self.optimize(**optimize_params)
        else:
            self.optimize(maxiter=5000, verbose=verbose)
        return True

    def systematic_splits(self, verbose=True):
        for kk in range(self.K):
            self.recursive_splits(kk, verbose=verbose)

    def recursive_splits(self, k=0, verbose=True, optimize_params=None):
        success = self.try_split(k, verbose=verbose, optimize_params=optimize_params)
        if success:
            if not k == (self.K - 1):
                self.recursive_splits(self.K - 1, verbose=verbose, optimize_params=optimize_params)
            self.recursive_splits(k, verbose=verbose, optimize_params=optimize_params)


# --------------------------------------------------------------------------- L4: MOG
class MOGOracle(MixtureOracle):
    """MOG.py:34-113."""

    def __init__(self, X, K=2, prior_Z="symmetric", alpha=1.0, prior_m=None, prior_kappa=1e-6, prior_S=None, prior_v=None):
        self.X = X
        self.N, self.D = X.shape
        self.m0 = self.X.mean(0) if prior_m is None else prior_m
        self.k0 = prior_kappa
        self.S0 = np.eye(self.D) * 1e-3 if prior_S is None else prior_S
        self.v0 = prior_v or self.D + 1.0
        self.k0m0m0T = self.k0 * self.m0[:, None] * self.m0[None, :]
        # MOG.py:47 -- quirk kept: log of the SQUARE ROOT of the Cholesky diagonal (a quarter log-det)
        self.S0_halflogdet = np.sum(np.log(np.sqrt(np.diag(np.linalg.cholesky(self.S0)))))
        MixtureOracle.__init__(self, self.N, K, prior_Z, alpha)
        self.do_computations()

    def do_computations(self):
        # MOG.py:52-61 (Ck formed as phi^T (x x^T) without materialising N x D x D)
        self.kNs = self.phi_hat + self.k0
        self.vNs = self.phi_hat + self.v0
        self.Xsumk = np.tensordot(self.X, self.phi, ((0), (0)))
        Ck = np.einsum("nk,ni,nj->ijk", self.phi, self.X, self.X)
        self.mun = (self.k0 * self.m0[:, None] + self.Xsumk) / self.kNs[None, :]
        munmunT = self.mun[:, None, :] * self.mun[None, :, :]
        self.Sns = self.S0[:, :, None] + Ck + self.k0m0m0T[:, :, None] - self.kNs[None, None, :] * munmunT
        self.Sns_inv, self.Sns_halflogdet = multiple_pdinv(self.Sns)

    def bound(self):
        # MOG.py:63-70
        return (-0.5 * self.D * np.sum(np.log(self.kNs / self.k0))
                + self.K * self.v0 * self.S0_halflogdet - np.sum(self.vNs * self.Sns_halflogdet)
                + np.sum(lngammad(self.vNs, self.D)) - self.K * lngammad(self.v0, self.D)
                + self.mixing_prop_bound()
                + self.H
                - 0.5 * self.N * self.D * np.log(np.pi))

    def vb_grad_natgrad(self):
        # MOG.py:72-84; the N x D x D x K temporary is replaced by an einsum of the same sum
        x_m = self.X[:, :, None] - self.mun[None, :, :]
        dlndtS_dphi = np.einsum("nik,ijk,njk->nk", x_m, self.Sns_inv, x_m)
        grad_phi = ((-0.5 * self.D / self.kNs + 0.5 * digamma((self.vNs - np.arange(self.D)[:, None]) / 2.0).sum(0)
                     + self.mixing_prop_bound_grad() - self.Sns_halflogdet - 1.0)
                    + (self.Hgrad - 0.5 * dlndtS_dphi * self.vNs))
        return self._natural(grad_phi)

    def predict_components_ln(self, Xnew):
        # MOG.py:87-98
        Dist = Xnew[:, :, None] - self.mun[None, :, :]
        tmp = np.sum(Dist[:, :, None, :] * self.Sns_inv[None, :, :, :], 1)
        mahalanobis = np.sum(tmp * Dist, 1) / (self.kNs + 1.0) * self.kNs * (self.vNs - self.D + 1.0)
        halflndetSigma = self.Sns_halflogdet + 0.5 * self.D * np.log((self.kNs + 1.0) / (self.kNs * (self.vNs - self.D + 1.0)))
        v = self.vNs[None, :]
        return (gammaln(0.5 * (v + 1.0)) - gammaln(0.5 * (v - self.D + 1.0))
                - (0.5 * self.D) * (np.log(v - self.D + 1.0) + np.log(np.pi))
                - halflndetSigma
                - (0.5 * (v + 1.0)) * np.log(1.0 + mahalanobis / (v - self.D + 1.0)))

    def predict_components(self, Xnew):
        return np.exp(self.predict_components_ln(Xnew))  # MOG.py:100-102

    def predict(self, Xnew):
        # MOG.py:104-113
        Z = self.predict_components(Xnew)
        pi = self.phi.sum(0) + self.alpha
        pi /= pi.sum()
        return (Z * pi[None, :]).sum(1)


# --------------------------------------------------------------------------- L4: MOHGP
class HyperOracle(object):
    """The slice of GPy / paramz the reference leans on for the non-variational parameters (collapsed_vb.py:313-323):
    Logexp-transformed optimiser vector and Model.optimize = L-BFGS-B on -log_likelihood.  Mixed into MOHGPOracle / OMGPOracle,
    which provide kernel_param_list(), kernel_gradients() and parameters_changed()."""

    hyper_free = False  # True: the kernel parameters are free, optimize() interleaves their optimisation (collapsed_vb.py:300-301)
    hyperparam_max_iters = 20  # collapsed_vb.py:32-35

    # paramz Logexp transform + GPy Model.optimize (default optimiser: scipy L-BFGS-B on -log_likelihood), collapsed_vb.py:313-323
    @staticmethod
    def _finv(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f > 36.0, f, np.log(np.expm1(np.minimum(f, 36.0))))

    @staticmethod
    def _f(x):
        x = np.asarray(x, dtype=np.float64)
        return np.where(x > 36.0, x, np.log1p(np.exp(np.clip(x, -745.0, 36.0))))

    def _set_x(self, x):
        for (k, name), v in zip(self.kernel_param_list(), self._f(x)):
            setattr(k, name, float(v))
        self.parameters_changed()

    def _objective_grads(self, x):
        try:
            self._set_x(x)
            f = np.array([getattr(k, name) for k, name in self.kernel_param_list()])
            g = self.kernel_gradients() * np.where(f > 36.0, 1.0, -np.expm1(-f))
            return -float(self.bound()), -g
        except LinAlgError:
            return np.inf, np.zeros(len(x))

    def optimize_parameters(self):
        if not self.hyper_free:
            return 0.0
        from scipy.optimize import fmin_l_bfgs_b
        start = self.bound()
        x0 = self._finv([getattr(k, name) for k, name in self.kernel_param_list()])
        x_opt, _, _ = fmin_l_bfgs_b(self._objective_grads, x0, maxfun=self.hyperparam_max_iters, maxiter=self.hyperparam_max_iters)
        self._set_x(x_opt)
        return self.bound() - start



class MOHGPOracle(HyperOracle, MixtureOracle):
    """MOHGP.py:33-90, :124-153.  `mahalanobis` selects the pairwise form: the weave substitution
    (default, utilities.py:147-170), the Python-3 fallback solve, or the whitened GEMM."""

    def __init__(self, X, kernF, kernY, Y, K=2, alpha=1.0, prior_Z="symmetric", mahalanobis=None):
        N, self.D = Y.shape
        self.Y, self.X = Y, X
        assert X.shape[0] == self.D, "input data don't match observations"
        if mahalanobis is not None:
            self.mahalanobis = mahalanobis
        MixtureOracle.__init__(self, N, K, prior_Z, alpha)
        self.kernF, self.kernY = kernF, kernY
        self.parameters_changed()

    def parameters_changed(self):
        # MOHGP.py:47-55 / :58-67 ; the dead N x D x D `YYT` (:52) is deliberately not built
        self.Sf = self.kernF.K(self.X)
        self.Sy = self.kernY.K(self.X)
        self.Sy_inv, self.Sy_chol, self.Sy_chol_inv, self.Sy_logdet = G.pdinv(self.Sy + np.eye(self.D) * 1e-6)
        self.YTY = np.dot(self.Y.T, self.Y)
        self.do_computations()

    def do_computations(self):
        # MOHGP.py:70-90
        self.ybark = np.dot(self.phi.T, self.Y).T
        A = G.backsub_both_sides(self.Sy_chol, self.Sf, transpose="right")
        self.Cs = [np.eye(self.D) + A * ph for ph in self.phi_hat]
        self._C_chols = [G.jitchol(C) for C in self.Cs]
        self.log_det_diff = np.array([2.0 * np.sum(np.log(np.diag(L))) for L in self._C_chols])
        T = [G.dtrtrs(L, self.Sy_chol.T, lower=1)[0] for L in self._C_chols]
        self.Lambda_inv = np.array([(self.Sy - np.dot(t.T, t)) / ph if (ph > 1e-6) else self.Sf
                                    for ph, t in zip(self.phi_hat, T)])
        self.Syi_ybark, _ = G.dpotrs(self.Sy_chol, self.ybark, lower=1)
        self.Syi_ybarkybarkT_Syi = self.Syi_ybark.T[:, None, :] * self.Syi_ybark.T[:, :, None]
        self.muk = (self.Lambda_inv * self.Syi_ybark.T[:, :, None]).sum(1).T

    def update_kern_grads(self):
        # MOHGP.py:92-122 -- derivative of the bound w.r.t. the kernel parameters, through dL/dSf and dL/dSy
        tmp = [G.dtrtrs(L, self.Sy_chol_inv, lower=1)[0] for L in self._C_chols]
        B_invs = [ph * np.dot(t.T, t) for ph, t in zip(self.phi_hat, tmp)]
        LiSfi = [np.eye(self.D) - np.dot(self.Sf, Bi) for Bi in B_invs]
        tmp1 = [np.dot(L.T, s) for L, s in zip(LiSfi, self.Syi_ybark.T)]
        tmp = 0.5 * sum([np.dot(t[:, None], t[None, :]) for t in tmp1])
        tmp += -0.5 * sum(B_invs)
        self.dL_dSf = tmp
        self.kernF.update_gradients_full(dL_dK=tmp, X=self.X)
        Byks = [np.dot(Bi, yk) for Bi, yk in zip(B_invs, self.ybark.T)]
        tmp = sum([np.dot(Byk[:, None], Byk[None, :]) / np.power(ph, 3) - SyyS / ph - Bi / ph
                   for Bi, Byk, ph, SyyS in zip(B_invs, Byks, self.phi_hat, self.Syi_ybarkybarkT_Syi) if ph > 1e-6])
        tmp = tmp + (self.K - self.N) * self.Sy_inv
        tmp = tmp + G.mdot(self.Sy_inv, self.YTY, self.Sy_inv)
        tmp = tmp / 2.0
        self.dL_dSy = tmp
        self.kernY.update_gradients_full(dL_dK=tmp, X=self.X)

    def kernel_param_list(self):
        return self.kernF.param_list() + self.kernY.param_list()

    def kernel_gradients(self):
        """d bound / d (kernel parameter), in link order kernF then kernY (what GPy hands to its optimiser, untransformed)."""
        self.update_kern_grads()
        return np.array([getattr(k, name + "_gradient") for k, name in self.kernel_param_list()])

    def bound(self):
        # MOHGP.py:124-135
        return (-0.5 * (self.N * self.D * np.log(2.0 * np.pi)
                        + self.log_det_diff.sum()
                        + self.N * self.Sy_logdet
                        + np.sum(self.YTY * self.Sy_inv))
                + 0.5 * np.sum(self.Syi_ybarkybarkT_Syi * self.Lambda_inv)
                + self.mixing_prop_bound()
                + self.H)

    def vb_grad_natgrad(self):
        # MOHGP.py:137-153
        ynmk2 = self.mahalanobis(self.Y, self.muk.T, self.Sy_chol)
        grad_phi = ((self.mixing_prop_bound_grad()
                     - 0.5 * np.sum(np.sum(self.Lambda_inv * self.Sy_inv[None, :, :], 1), 1))
                    + (self.Hgrad - 0.5 * ynmk2))
        return self._natural(grad_phi)


    def predict_components(self, Xnew):
        # MOHGP.py:156-174 (the TypeError branch for hierarchical kernY is GPy-specific and out of scope)
        tmp = [G.dtrtrs(L, self.Sy_chol_inv, lower=1)[0] for L in self._C_chols]
        B_invs = [ph * np.dot(t.T, t) for ph, t in zip(self.phi_hat, tmp)]
        kx = self.kernF.K(self.X, Xnew)
        kxx = self.kernF.K(Xnew) + self.kernY.K(Xnew)
        tmp = [np.eye(self.D) - np.dot(Bi, self.Sf) for Bi in B_invs]
        mu = [G.mdot(kx.T, t, self.Sy_inv, yb) for t, yb in zip(tmp, self.ybark.T)]
        var = [kxx - G.mdot(kx.T, Bi, kx) for Bi in B_invs]
        return mu, var


# --------------------------------------------------------------------------- L4: OMGP
class OMGPOracle(HyperOracle, MixtureOracle):
    """OMGP.py:14-53, :96-147 (bound and approximate natural gradient; appendix A.11 quirks kept)."""

    def __init__(self, X, Y, K=2, kernels=None, variance=1.0, alpha=1.0, prior_Z="symmetric"):
        N, self.D = Y.shape
        self.Y, self.X = Y, X
        self.YYT = G.tdot(self.Y)
        self.kern = [G.RBF(input_dim=1) for _ in range(K)] if kernels is None else kernels
        self.variance = float(variance)
        MixtureOracle.__init__(self, N, K, prior_Z, alpha)

    def do_computations(self):
        # OMGP.py:40-53: grow the kernel list by ONE per call, shrink to K
        if len(self.kern) < self.K:
            self.kern.append(self.kern[-1].copy())
        if len(self.kern) > self.K:
            self.kern = self.kern[:self.K]

    def bound(self):
        # OMGP.py:96-123
        GP_bound = 0.0
        for i, kern in enumerate(self.kern):
            Kmat = kern.K(self.X)
            B_inv = np.diag(1.0 / ((self.phi[:, i] + 1e-6) / self.variance))
            Bi, LB, LBi, Blogdet = G.pdinv(Kmat + B_inv)
            GP_bound -= 0.5 * G.dpotrs(LB, self.YYT)[0].trace()
            GP_bound -= 0.5 * Blogdet
            GP_bound -= 0.5 * self.D * np.einsum("j,j->", self.phi[:, i], np.log(2 * np.pi * self.variance) * np.ones(self.N))
        return GP_bound + self.mixing_prop_bound() + self.H

    def vb_grad_natgrad(self):
        # OMGP.py:125-147
        grad_Lm = np.zeros_like(self.phi)
        for i, kern in enumerate(self.kern):
            Kmat = kern.K(self.X)
            B_inv = np.diag(1.0 / ((self.phi[:, i] + 1e-6) / self.variance))
            K_B_inv, L_B, _, _ = G.pdinv(Kmat + B_inv)
            alpha, _ = G.dpotrs(L_B, self.Y)
            dL_dB_diag = np.sum(np.square(alpha), 1) - np.diag(K_B_inv)
            grad_Lm[:, i] = -0.5 * self.variance * dL_dB_diag / (self.phi[:, i] ** 2 + 1e-6)
        grad_phi = grad_Lm + self.mixing_prop_bound_grad() + self.Hgrad
        return self._natural(grad_phi)

    def parameters_changed(self):
        pass  # OMGP.py:35-38 only refreshes the gradients; bound() / vb_grad_natgrad() recompute everything from the kernels

    def update_kern_grads(self):
        """OMGP.py:55-94: gradients of the bound w.r.t. the kernel parameters of every component and the noise variance.
        Quirk kept (appendix A.11): the kernel gradients use B = K + diag(variance / phi) with NO 1e-6, and the same
        dL_dB then enters the variance gradient together with 1 / (phi + 1e-6)."""
        grad_Lm_variance = 0.0
        for i, kern in enumerate(self.kern):
            Kmat = kern.K(self.X)
            B_inv = np.diag(1.0 / (self.phi[:, i] / self.variance))
            Bi, LB, LBi, Blogdet = G.pdinv(Kmat + B_inv)
            tmp = G.dpotrs(LB, self.YYT)[0]
            tmp[np.diag_indices(self.N)] -= 1.0  # GPy.util.diag.subtract(tmp, 1)
            dL_dB = G.dpotrs(LB, tmp.T)[0]
            kern.update_gradients_full(dL_dK=0.5 * dL_dB, X=self.X)
            grad_B_inv = np.diag(1.0 / (self.phi[:, i] + 1e-6))
            grad_Lm_variance += 0.5 * np.trace(np.dot(dL_dB, grad_B_inv))
            grad_Lm_variance -= 0.5 * self.D * np.einsum("j,j->", self.phi[:, i], (1.0 / self.variance) * np.ones(self.N))
        self.variance_gradient = float(grad_Lm_variance)

    def kernel_param_list(self):
        """GPy link order (OMGP.py:31-32): the noise variance first, then the kernels of the components."""
        out = [(self, "variance")]
        for k in self.kern:
            out.extend(k.param_list())
        return out

    def kernel_gradients(self):
        self.update_kern_grads()
        return np.array([getattr(o, n + "_gradient") for o, n in self.kernel_param_list()])
