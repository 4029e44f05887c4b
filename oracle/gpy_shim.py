"""TEST INFRASTRUCTURE ONLY -- CPU oracle support code, never imported by the product package.

GPy is a third-party dependency of the reference (setup.py:14 `install_requires=['GPy>=0.6']`,
README.md:86 "tested the demos with GPy v0.8") that is absent from /root/reference and cannot be
installed here (no network).  This module restates, from GPy's published behaviour, exactly the thin
LAPACK wrappers and covariance functions the hot path calls:

  GPy.util.linalg.{jitchol, pdinv, dtrtrs, dpotrs, dpotri, backsub_both_sides, mdot, tdot}
      call sites: MOHGP.py:7,49,63,79,82,84,88  OMGP.py:7-8,18,108,113,136-137  np_utilities.py:18,20
  GPy.kern.{RBF, Matern32, Matern52, Bias, White, Add}.K(X, X2=None)
      call sites: MOHGP.py:47-48,61-62  OMGP.py:25,104,132

Parity status: the MOG path (which touches only jitchol/dpotri) is pinned by the reference's seeded
notebook trace (tests/golden/mog_notebook_golden.json).  The kernel functions are pinned only by
their closed forms -- "parity unpinned" for them (no GPy output is available to compare against).
"""
import numpy as np
from scipy import linalg as sla
from scipy.linalg import lapack


class LinAlgError(np.linalg.LinAlgError):
    pass


def jitchol(A, maxtries=5):
    """Lower Cholesky with GPy's jitter retry: dpotrf; on failure add mean(diag)*1e-6*10^i, 5 tries."""
    A = np.ascontiguousarray(A)
    L, info = lapack.dpotrf(A, lower=1)
    if info == 0:
        return np.tril(L)
    d = np.diag(A)
    if np.any(d <= 0.0):
        raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = d.mean() * 1e-6
    for _ in range(maxtries):
        if not np.isfinite(jitter):
            break
        L, info = lapack.dpotrf(A + np.eye(A.shape[0]) * jitter, lower=1)
        if info == 0:
            return np.tril(L)
        jitter *= 10.0
    raise np.linalg.LinAlgError("not positive definite, even with jitter.")


def dtrtrs(A, B, lower=1, trans=0, unitdiag=0):
    """Triangular solve; returns (X, info) like the LAPACK wrapper."""
    X, info = lapack.dtrtrs(np.asfortranarray(A), np.asfortranarray(B), lower=lower, trans=trans, unitdiag=unitdiag)
    return X, info


def dpotrs(L, B, lower=1):
    """Solve (L L^T) X = B given the Cholesky factor; returns (X, info)."""
    X, info = lapack.dpotrs(np.asfortranarray(L), np.asfortranarray(B), lower=lower)
    return X, info


def dpotri(L, lower=1):
    """Inverse from the Cholesky factor; only the `lower`/upper triangle LAPACK wrote is meaningful
    before symmetrisation (np_utilities.py:20-21 symmetrises from the upper triangle with lower=True
    in Fortran order = upper in C order)."""
    Ai, info = lapack.dpotri(np.asfortranarray(L), lower=lower)
    # GPy.util.linalg.dpotri symmetrises the result
    Ai = np.asarray(Ai)
    if lower:
        Ai = np.tril(Ai) + np.tril(Ai, -1).T
    else:
        Ai = np.triu(Ai) + np.triu(Ai, 1).T
    return Ai, info


def pdinv(A):
    """-> (A^-1, L, L^-1, logdet A) with L = jitchol(A)."""
    L = jitchol(A)
    logdet = 2.0 * np.sum(np.log(np.diag(L)))
    Li = dtrtrs(L, np.eye(L.shape[0]), lower=1)[0]
    Ai, _ = dpotri(L, lower=1)
    return Ai, L, Li, logdet


def backsub_both_sides(L, X, transpose="left"):
    """'right': L^-1 X L^-T ; 'left': L^-T X L^-1."""
    if transpose == "left":
        tmp, _ = dtrtrs(L, X, lower=1, trans=1)
        return dtrtrs(L, tmp.T, lower=1, trans=1)[0].T
    tmp, _ = dtrtrs(L, X, lower=1, trans=0)
    return dtrtrs(L, tmp.T, lower=1, trans=0)[0].T


def mdot(*args):
    out = args[0]
    for a in args[1:]:
        out = np.dot(out, a)
    return out


def tdot(M):
    return np.dot(M, M.T)


# ----------------------------------------------------------------------------- covariance functions
class Kern(object):
    """Minimal stand-in for GPy.kern.Kern: `.K(X, X2=None)`, `+` builds an Add, `.copy()`."""

    input_dim = 1

    def K(self, X, X2=None):
        raise NotImplementedError

    def __add__(self, other):
        return Add([self, other])

    def copy(self):
        import copy
        return copy.deepcopy(self)

    def spec(self):
        """Flat description consumed by the device kernel-fill (kind, variance, lengthscale)."""
        raise NotImplementedError


class _Stationary(Kern):
    def __init__(self, input_dim=1, variance=1.0, lengthscale=1.0, name=None):
        self.input_dim = input_dim
        self.variance = float(variance)
        self.lengthscale = float(lengthscale)

    def _scaled_dist(self, X, X2):
        # GPy kern/src/stationary.py: r^2 = -2 X X2^T + |x|^2 + |x2|^2, diagonal zeroed when X2 is None,
        # clipped at 0, r = sqrt(r2)/lengthscale
        X = np.atleast_2d(X)
        if X2 is None:
            Xsq = np.sum(np.square(X), 1)
            r2 = -2.0 * np.dot(X, X.T) + (Xsq[:, None] + Xsq[None, :])
            r2[np.diag_indices(X.shape[0])] = 0.0
        else:
            X2 = np.atleast_2d(X2)
            r2 = -2.0 * np.dot(X, X2.T) + (np.sum(np.square(X), 1)[:, None] + np.sum(np.square(X2), 1)[None, :])
        r2 = np.clip(r2, 0.0, np.inf)
        return np.sqrt(r2) / self.lengthscale

    def K(self, X, X2=None):
        return self.K_of_r(self._scaled_dist(X, X2))

    def update_gradients_full(self, dL_dK, X, X2=None):
        # GPy Stationary.update_gradients_full (non-ARD): variance.gradient = sum(K o dL_dK)/variance,
        # lengthscale.gradient = -sum(dK_dr o dL_dK o r)/lengthscale
        r = self._scaled_dist(X, X2)
        self.variance_gradient = float(np.sum(self.K_of_r(r) * dL_dK) / self.variance)
        self.lengthscale_gradient = float(-np.sum(self.dK_dr(r) * dL_dK * r) / self.lengthscale)

    def param_list(self):
        return [(self, "variance"), (self, "lengthscale")]

    def spec(self):
        return [(self.kind, self.variance, self.lengthscale)]


class RBF(_Stationary):
    kind = "rbf"

    def K_of_r(self, r):
        return self.variance * np.exp(-0.5 * r ** 2)

    def dK_dr(self, r):
        return -r * self.K_of_r(r)


class Matern32(_Stationary):
    kind = "matern32"

    def K_of_r(self, r):
        return self.variance * (1.0 + np.sqrt(3.0) * r) * np.exp(-np.sqrt(3.0) * r)

    def dK_dr(self, r):
        return -3.0 * self.variance * r * np.exp(-np.sqrt(3.0) * r)


class Matern52(_Stationary):
    kind = "matern52"

    def K_of_r(self, r):
        return self.variance * (1.0 + np.sqrt(5.0) * r + 5.0 / 3.0 * r ** 2) * np.exp(-np.sqrt(5.0) * r)

    def dK_dr(self, r):
        return self.variance * (10.0 / 3.0 * r - 5.0 * r - 5.0 * np.sqrt(5.0) / 3.0 * r ** 2) * np.exp(-np.sqrt(5.0) * r)


class Bias(Kern):
    kind = "bias"

    def __init__(self, input_dim=1, variance=1.0, name=None):
        self.input_dim = input_dim
        self.variance = float(variance)

    def K(self, X, X2=None):
        n2 = X.shape[0] if X2 is None else X2.shape[0]
        return np.full((X.shape[0], n2), self.variance)

    def update_gradients_full(self, dL_dK, X, X2=None):
        self.variance_gradient = float(np.sum(dL_dK))

    def param_list(self):
        return [(self, "variance")]

    def spec(self):
        return [(self.kind, self.variance, 1.0)]


class White(Kern):
    kind = "white"

    def __init__(self, input_dim=1, variance=1.0, name=None):
        self.input_dim = input_dim
        self.variance = float(variance)

    def K(self, X, X2=None):
        if X2 is None:
            return np.eye(X.shape[0]) * self.variance
        return np.zeros((X.shape[0], X2.shape[0]))

    def update_gradients_full(self, dL_dK, X, X2=None):
        self.variance_gradient = float(np.trace(dL_dK)) if X2 is None else 0.0

    def param_list(self):
        return [(self, "variance")]

    def spec(self):
        return [(self.kind, self.variance, 1.0)]


class Add(Kern):
    def __init__(self, parts):
        self.parts = []
        for p in parts:
            self.parts.extend(p.parts if isinstance(p, Add) else [p])
        self.input_dim = self.parts[0].input_dim

    def K(self, X, X2=None):
        out = self.parts[0].K(X, X2)
        for p in self.parts[1:]:
            out = out + p.K(X, X2)
        return out

    def update_gradients_full(self, dL_dK, X, X2=None):
        for p in self.parts:  # every part of a sum receives the same dL_dK
            p.update_gradients_full(dL_dK, X, X2)

    def param_list(self):
        out = []
        for p in self.parts:
            out.extend(p.param_list())
        return out

    def spec(self):
        s = []
        for p in self.parts:
            s.extend(p.spec())
        return s
