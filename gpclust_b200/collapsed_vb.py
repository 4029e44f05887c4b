"""Host mirror of GPclust/collapsed_vb.py (class CollapsedVB): the Riemannian conjugate natural-gradient
loop.  Control flow, iteration accounting, messages and termination follow collapsed_vb.py:143-193 and
:228-303 line for line; the N*K-long vector algebra of :152-172 is NOT done on the host -- it is
delegated to four device hooks that the mixture layer implements with the fused sm_100a passes:

    _cg_gradient(with_old)                G pass at the accepted point (natgrad + the three CG dots)
    _cg_trial(beta_mode, step, sq_old)    S pass: search direction, trial point, stats, bound
                                          -> (bound, squareNorm, beta, failed)
    _cg_accept() / _cg_restore()          buffer rotation / fall back to the accepted point

GPy is not a dependency: the model base class here is a plain object that carries the slice of GPy / paramz
the reference leans on -- the Logexp-transformed `optimizer_array` of the free (kernel, noise) parameters and
`optimize_parameters()` (collapsed_vb.py:313-323 -> GPy Model.optimize = L-BFGS-B on -bound).  With no free
parameters `optimize_parameters()` returns 0. exactly like the reference (:322-323) and `optimize()` runs the
loop from the device-resident state (collapsed_mixture._run_on_device) instead of the host-driven one below.
"""
import sys
import warnings

import numpy as np
from numpy.linalg import LinAlgError


class LinAlgWarning(Warning):
    pass


class CollapsedVB(object):
    def __init__(self, name):
        self.name = name
        self.hyperparam_interval = 50  # collapsed_vb.py:29
        self.default_method = "HS"
        self.hyperparam_opt_args = {"max_iters": 20, "messages": 1, "clear_after_finish": True}

    # ---- template methods (collapsed_vb.py:41-54)
    def randomize(self):
        self.set_vb_param(np.random.randn(self.get_vb_param().size))  # collapsed_vb.py:37-39

    def get_vb_param(self):
        raise NotImplementedError

    def set_vb_param(self, x):
        raise NotImplementedError

    def bound(self):
        raise NotImplementedError

    def vb_grad_natgrad(self):
        raise NotImplementedError

    def log_likelihood(self):
        return self.bound()  # collapsed_vb.py:56-61

    # ---- non-variational (kernel) parameters: the slice of GPy / paramz the reference leans on (collapsed_vb.py:313-323).
    # Every kernel parameter is positive-constrained through paramz's Logexp transform f(x) = log(1 + e^x); the
    # optimiser works on x.  `_kernel_param_list()` -> [(object, attribute), ...] of the FREE parameters.
    _LIM = 36.0

    def _kernel_param_list(self):
        return []

    @staticmethod
    def _logexp_finv(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f > CollapsedVB._LIM, f, np.log(np.expm1(np.minimum(f, CollapsedVB._LIM))))

    @staticmethod
    def _logexp_f(x):
        x = np.asarray(x, dtype=np.float64)
        return np.where(x > CollapsedVB._LIM, x, np.log1p(np.exp(np.clip(x, -745.0, CollapsedVB._LIM))))

    @property
    def optimizer_array(self):
        """Untransformed free parameters (GPy Model.optimizer_array)."""
        plist = self._kernel_param_list()
        if not plist:
            return np.zeros(0)
        return self._logexp_finv([getattr(o, n) for o, n in plist])

    @optimizer_array.setter
    def optimizer_array(self, x):
        plist = self._kernel_param_list()
        vals = self._logexp_f(x)
        changed = False
        for (o, n), v in zip(plist, vals):
            if getattr(o, n) != float(v):
                setattr(o, n, float(v))
                changed = True
        if changed:
            self.parameters_changed()

    def parameters_changed(self):
        pass

    def _kernel_gradients(self):
        """d bound / d (free kernel parameter), constrained space, same order as _kernel_param_list()."""
        return np.array([getattr(o, n + "_gradient") for o, n in self._kernel_param_list()])

    def _objective_grads(self, x):
        """GPy Model._objective_grads: (-log_likelihood, -transformed gradients); a failed factorisation counts as +inf."""
        try:
            self.optimizer_array = x
            f = np.asarray([getattr(o, n) for o, n in self._kernel_param_list()])
            g = self._kernel_gradients() * np.where(f > self._LIM, 1.0, -np.expm1(-f))  # Logexp.gradfactor
            return -float(self.log_likelihood()), -g
        except LinAlgError:
            return np.inf, np.zeros(len(x))

    def optimize_parameters(self):
        """collapsed_vb.py:313-323: optimise the non-variational parameters with GPy's default optimiser (L-BFGS-B,
        `hyperparam_opt_args['max_iters']` function evaluations); returns the increment of the bound."""
        if self.optimizer_array.size > 0:
            from scipy.optimize import fmin_l_bfgs_b
            start = self.bound()
            x0 = self.optimizer_array.copy()
            max_iters = int(self.hyperparam_opt_args.get("max_iters", 20))
            x_opt, f_opt, info = fmin_l_bfgs_b(self._objective_grads, x0, maxfun=max_iters, maxiter=max_iters)
            self.optimizer_array = x_opt
            return self.bound() - start
        else:
            return 0.0

    # ---- device hooks
    def _cg_gradient(self, with_old):
        raise NotImplementedError

    def _cg_trial(self, beta_mode, step_length, sq_old):
        raise NotImplementedError

    def _cg_accept(self):
        raise NotImplementedError

    def _cg_restore(self):
        raise NotImplementedError

    def _cg_failed_pass_scalars(self):
        """(squareNorm, beta) of the pass whose two trial points both failed (device loop, GPC_DONE_FAULT)."""
        raise NotImplementedError

    def _is_root(self):
        return True

    def _device_loop_ok(self):
        return False

    def _run_on_device(self, *args, **kwargs):
        raise NotImplementedError

    def _cg_iteration(self, first, method, step_length, bound_old, squareNorm_old):
        """One pass of the loop body collapsed_vb.py:152-193 -> (bound, squareNorm, beta, bound evaluations)."""
        # :152-167 -- gradient at the accepted point, beta, search direction (HS / PR need the previous point)
        self._cg_gradient(with_old=(not first) and method in ("HS", "PR"))
        beta_mode = "zero" if (method == "steepest" or first) else method

        # :170-176 -- try a conjugate step
        bound, squareNorm, beta, failed = self._cg_trial(beta_mode, step_length, squareNorm_old)
        if failed:
            bound = bound_old - 1
        evals = 1

        # :182-193 -- revert to steepest (== VBEM) if the bound fell
        if bound < bound_old:
            bound, _, _, failed = self._cg_trial("steepest_retry", step_length, squareNorm_old)
            if failed:
                warnings.warn("Caught LinalgError in setting variational parameters, trying to continue with old parameter settings",
                              LinAlgWarning)
                bound = self._cg_restore()
            evals = 2
        self._cg_accept()
        return bound, squareNorm, beta, evals

    def optimize(self, method="HS", maxiter=500, ftol=1e-6, gtol=1e-6, step_length=1.0, callback=None, verbose=True):
        """Conjugate natural-gradient ascent on the variational parameters; collapsed_vb.py:63-310."""
        assert method in ["FR", "PR", "HS", "steepest"], "invalid conjugate gradient method specified."
        say = verbose and self._is_root()
        iteration = 0
        bound_old = self.bound()
        squareNorm_old = 0.0
        first = True
        self.optimize_trace = []
        pending = None
        if callback is None and self.optimizer_array.size == 0 and self._device_loop_ok():
            # no per-iteration host work is required: accept / reject and the termination tests run on the device,
            # the host enqueues batches of passes and reads the trace back once per batch
            iteration, bound_old, squareNorm_old, first, done, rows = self._run_on_device(method, maxiter, ftol, gtol, step_length, iteration,
                                                                                       bound_old, squareNorm_old, first, say=say)
            self.optimize_trace.extend(rows)
            if not (done & 8):
                if say:  # the reference tests ftol, gtol, maxiter in this order and breaks at the first hit (collapsed_vb.py:229-291)
                    print("vb converged (ftol)" if done & 1 else ("vb converged (gtol)" if done & 2 else "maxiter exceeded"))
                return iteration
            # GPC_DONE_FAULT: the conjugate step AND the steepest retry failed to factorise (collapsed_vb.py:187-190).  The device
            # state belongs to the failed trial point: warn, recompute at the accepted point (set_vb_param(phi_old); bound =
            # self.bound()), count the two evaluations and fall through to the termination tests below, as the reference does.
            warnings.warn("Caught LinalgError in setting variational parameters, trying to continue with old parameter settings",
                          LinAlgWarning)
            sq, beta = self._cg_failed_pass_scalars()
            pending = (self._cg_restore(), sq, beta, 2)
        while True:
            if pending is not None:
                (bound, squareNorm, beta, evals), pending = pending, None
            else:
                if callback is not None:
                    callback()
                bound, squareNorm, beta, evals = self._cg_iteration(first, method, step_length, bound_old, squareNorm_old)
            first = False
            iteration += evals

            self.optimize_trace.append((iteration, bound, squareNorm, beta))
            if say:
                sys.stdout.write("\riteration " + str(iteration) + " bound=" + str(bound) + " grad=" + str(squareNorm) + ", beta=" + str(beta) + "\n")
                sys.stdout.flush()

            # collapsed_vb.py:229-291
            if np.abs(bound - bound_old) <= ftol:
                if say:
                    print("vb converged (ftol)")
                if self.optimize_parameters() < 1e-1:
                    break
            if squareNorm <= gtol:
                if say:
                    print("vb converged (gtol)")
                if self.optimize_parameters() < 1e-1:
                    break
            if iteration >= maxiter:
                if say:
                    print("maxiter exceeded")
                break

            squareNorm_old = squareNorm  # collapsed_vb.py:294-297 (the vectors stay on the device)
            if (iteration > 1) and not (iteration % self.hyperparam_interval):
                self.optimize_parameters()  # collapsed_vb.py:300-301
            bound_old = bound
        return iteration


__all__ = ["CollapsedVB", "LinAlgError", "LinAlgWarning"]
