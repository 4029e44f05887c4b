"""Covariance functions with GPy's constructor signatures and `.K(X, X2=None)` meaning, evaluated by
the device fill `gpc_kern_fill` (call sites in the reference: MOHGP.py:47-48,61-62; OMGP.py:25,104,132).

GPy itself is not a dependency of this package.  A foreign kernel object (e.g. a real GPy kernel) may
still be passed to MOHGP / OMGP: anything without `.parts_spec()` is evaluated through its own `.K`
on the host and the D x D result uploaded (see `cov_to_device`).
"""
import copy as _copy
import ctypes

import numpy as np

from . import _cabi


class Kern(object):
    input_dim = 1
    name = "kern"

    def parts_spec(self):
        """-> list of (kind, variance, lengthscale) consumed by gpc_kern_fill."""
        raise NotImplementedError

    def K_device(self, X_dev, X2_dev=None, diag_add=0.0, out=None):
        """Fill the covariance on the device. X_dev: (n, q) float64 CUDA tensor. Returns an (n, m) CUDA tensor."""
        torch = _cabi.require_cuda()
        spec = self.parts_spec()
        n, q = X_dev.shape
        m = n if X2_dev is None else X2_dev.shape[0]
        if out is None:
            out = torch.empty((n, m), dtype=torch.float64, device=X_dev.device)
        kinds = _cabi.as_c_array([_cabi.KERN[s[0]] for s in spec], ctypes.c_int)
        var = _cabi.as_c_array([float(s[1]) for s in spec], ctypes.c_double)
        ls = _cabi.as_c_array([float(s[2]) for s in spec], ctypes.c_double)
        _cabi.check(_cabi.lib.gpc_kern_fill(_cabi.ptr(X_dev), _cabi.ptr(X2_dev), n, m, q, len(spec), kinds, var, ls, float(diag_add),
                                            _cabi.ptr(out), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "gpc_kern_fill")
        return out

    def K(self, X, X2=None):
        """numpy in, numpy out -- the GPy calling convention."""
        torch = _cabi.require_cuda()
        Xd = torch.from_numpy(_cabi.f64(np.atleast_2d(X))).cuda()
        X2d = None if X2 is None else torch.from_numpy(_cabi.f64(np.atleast_2d(X2))).cuda()
        return self.K_device(Xd, X2d).cpu().numpy()

    def __add__(self, other):
        return Add([self, other])

    def copy(self):
        return _copy.deepcopy(self)

    # ---- hyper-parameter interface (GPy: every kernel parameter is a positive-constrained Param; `fix()` removes it
    # from the optimiser's vector).  param_list() -> [(object, attribute name), ...] in GPy's link order.
    fixed = False

    def fix(self):
        self.fixed = True
        for p in getattr(self, "parts", []):
            p.fix()
        return self

    def unfix(self):
        self.fixed = False
        for p in getattr(self, "parts", []):
            p.unfix()
        return self

    constrain_fixed = fix
    unconstrain_fixed = unfix

    def param_list(self):
        return []

    def update_gradients_full(self, dL_dK, X, X2=None):
        """Set `<name>_gradient` for every parameter from dL/dK (GPy Kern.update_gradients_full; D x D host algebra)."""
        raise NotImplementedError

    def parameters_changed(self):
        pass

    @property
    def param_array(self):
        out = []
        for kind, var, ls in self.parts_spec():
            out.append(var)
            if kind in ("rbf", "matern32", "matern52"):
                out.append(ls)
        return np.array(out)


class _Stationary(Kern):
    kind = None

    def __init__(self, input_dim=1, variance=1.0, lengthscale=1.0, ARD=False, active_dims=None, name=None):
        if ARD:
            raise NotImplementedError("ARD lengthscales are not part of the accelerated path")
        self.input_dim = int(input_dim)
        self.variance = float(variance)
        self.lengthscale = float(np.asarray(lengthscale).reshape(-1)[0]) if lengthscale is not None else 1.0
        self.name = name or self.kind

    def parts_spec(self):
        return [(self.kind, self.variance, self.lengthscale)]

    def param_list(self):
        return [] if self.fixed else [(self, "variance"), (self, "lengthscale")]

    def _scaled_dist(self, X, X2=None):
        # GPy kern/src/stationary.py: r^2 = -2 X X2^T + |x|^2 + |x2|^2, diagonal zeroed when X2 is None, clipped at 0
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        if X2 is None:
            Xsq = np.sum(np.square(X), 1)
            r2 = -2.0 * np.dot(X, X.T) + (Xsq[:, None] + Xsq[None, :])
            r2[np.diag_indices(X.shape[0])] = 0.0
        else:
            X2 = np.atleast_2d(np.asarray(X2, dtype=np.float64))
            r2 = -2.0 * np.dot(X, X2.T) + (np.sum(np.square(X), 1)[:, None] + np.sum(np.square(X2), 1)[None, :])
        return np.sqrt(np.clip(r2, 0.0, np.inf)) / self.lengthscale

    def update_gradients_full(self, dL_dK, X, X2=None):
        # GPy Stationary.update_gradients_full (non-ARD)
        r = self._scaled_dist(X, X2)
        self.variance_gradient = float(np.sum(self._K_of_r(r) * dL_dK) / self.variance)
        self.lengthscale_gradient = float(-np.sum(self._dK_dr(r) * dL_dK * r) / self.lengthscale)


class RBF(_Stationary):
    kind = "rbf"

    def _K_of_r(self, r):
        return self.variance * np.exp(-0.5 * r ** 2)

    def _dK_dr(self, r):
        return -r * self._K_of_r(r)


class Matern32(_Stationary):
    kind = "matern32"

    def _K_of_r(self, r):
        return self.variance * (1.0 + np.sqrt(3.0) * r) * np.exp(-np.sqrt(3.0) * r)

    def _dK_dr(self, r):
        return -3.0 * self.variance * r * np.exp(-np.sqrt(3.0) * r)


class Matern52(_Stationary):
    kind = "matern52"

    def _K_of_r(self, r):
        return self.variance * (1.0 + np.sqrt(5.0) * r + 5.0 / 3.0 * r ** 2) * np.exp(-np.sqrt(5.0) * r)

    def _dK_dr(self, r):
        return self.variance * (10.0 / 3.0 * r - 5.0 * r - 5.0 * np.sqrt(5.0) / 3.0 * r ** 2) * np.exp(-np.sqrt(5.0) * r)


class Bias(Kern):
    def __init__(self, input_dim=1, variance=1.0, active_dims=None, name="bias"):
        self.input_dim, self.variance, self.name = int(input_dim), float(variance), name

    def parts_spec(self):
        return [("bias", self.variance, 1.0)]

    def param_list(self):
        return [] if self.fixed else [(self, "variance")]

    def update_gradients_full(self, dL_dK, X, X2=None):
        self.variance_gradient = float(np.sum(dL_dK))


class White(Kern):
    def __init__(self, input_dim=1, variance=1.0, active_dims=None, name="white"):
        self.input_dim, self.variance, self.name = int(input_dim), float(variance), name

    def parts_spec(self):
        return [("white", self.variance, 1.0)]

    def param_list(self):
        return [] if self.fixed else [(self, "variance")]

    def update_gradients_full(self, dL_dK, X, X2=None):
        self.variance_gradient = float(np.trace(dL_dK)) if X2 is None else 0.0


class Add(Kern):
    def __init__(self, parts, name="add"):
        self.parts = []
        for p in parts:
            self.parts.extend(p.parts if isinstance(p, Add) else [p])
        if len(self.parts) > 8:
            raise ValueError("at most 8 summed covariance parts are supported")
        self.input_dim = self.parts[0].input_dim
        self.name = name

    def parts_spec(self):
        s = []
        for p in self.parts:
            s.extend(p.parts_spec())
        return s

    def param_list(self):
        out = []
        for p in self.parts:
            out.extend(p.param_list())
        return out

    def update_gradients_full(self, dL_dK, X, X2=None):
        for p in self.parts:  # every part of a sum receives the same dL_dK
            p.update_gradients_full(dL_dK, X, X2)


def cov_to_device(kern, X_host, X_dev):
    """D x D covariance of `kern` on the device: our kernels fill on the device, foreign objects (GPy)
    are evaluated by their own `.K` on the host and uploaded."""
    torch = _cabi.require_cuda()
    if hasattr(kern, "parts_spec"):
        return kern.K_device(X_dev)
    return torch.from_numpy(_cabi.f64(kern.K(X_host))).to(X_dev.device)
