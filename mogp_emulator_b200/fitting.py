"""``fit_GP_MAP`` for the GPU emulators -- the dispatch of mogp_emulator/fitting.py:16-186 restricted to
``GaussianProcessGPU`` / ``MultiOutputGP_GPU`` (fitting.py:152-155, 189-217).

The reference GPU path hands the search to a C++ dlib BFGS loop (mogp_gpu/src/fitting.hpp:61-120).  Here the
optimiser is the CPU reference's own -- ``scipy.optimize.minimize`` (L-BFGS-B) over ``logposterior`` /
``logpost_deriv`` (fitting.py:219-266) -- with every objective and gradient evaluation running on the GPU,
so results are comparable with the CPU ``GaussianProcess`` fit by fit.
"""
import numpy as np
from scipy.optimize import minimize

from .GaussianProcessGPU import GaussianProcessGPU
from .MultiOutputGP_GPU import MultiOutputGP_GPU


def _minimise_one(logpost, grad, sample, n_params, n_tries, theta0, method, options):
    """n_tries restarts (first from theta0 when given, the rest from prior samples); a restart whose
    factorisation fails or whose arithmetic overflows is skipped (fitting.py:237-258)."""
    best_val, best_theta = None, None
    for i in range(n_tries):
        if i == 0 and theta0 is not None:
            start = np.array(theta0, dtype=np.float64)
            assert start.shape == (n_params,), "theta0 must be a 1D array with length n_params"
        else:
            start = sample()
        try:
            with np.errstate(divide="raise", over="raise", invalid="raise"):
                res = minimize(logpost, start, method=method, jac=grad, options=options)
        except RuntimeError:
            print("Matrix not positive definite, skipping this iteration")
            continue
        except FloatingPointError:
            print("Floating point error in optimization routine, skipping this iteration")
            continue
        if best_val is None or res["fun"] < best_val:
            best_val, best_theta = res["fun"], res["x"]
    return best_theta


def _check_method(method):
    if method not in ["L-BFGS", "L-BFGS-B"]:
        raise NotImplementedError("Unknown method for optimizer - only L-BFGS implemented for GPU")
    return "L-BFGS-B"


def _fit_single_GPGPU_MAP(gp, n_tries=15, theta0=None, method="L-BFGS-B", **kwargs):
    method = _check_method(method)
    n_tries = int(n_tries)
    assert n_tries > 0, "number of attempts must be positive"
    if theta0 is not None and len(theta0) == 0:
        theta0 = None
    best = _minimise_one(gp.logposterior, gp.logpost_deriv, gp.priors.sample, gp.n_params, n_tries, theta0, method,
                         kwargs)
    if best is None:
        print("Minimization routine failed to return a value")
        gp.theta = None
    else:
        gp.fit(best)
    if not gp.theta.data_has_been_set():
        raise RuntimeError("Fitting did not converge")
    return gp


def _fit_MOGPGPU_MAP(gp, n_tries=15, theta0=None, method="L-BFGS-B", refit=False, **kwargs):
    method = _check_method(method)
    kwargs.pop("processes", None)
    n_tries = int(n_tries)
    assert n_tries > 0, "n_tries must be a positive integer"
    E = gp.n_emulators
    if theta0 is None or (hasattr(theta0, "__len__") and len(theta0) == 0):
        starts = [None] * E
    elif isinstance(theta0, np.ndarray):
        if theta0.ndim == 1:
            starts = [theta0] * E
        else:
            assert theta0.ndim == 2, "theta0 must be a 1D or 2D array"
            assert theta0.shape[0] == E, "bad shape for fitting starting points"
            starts = list(theta0)
    else:
        assert len(theta0) == E, "theta0 must be a list of length n_emulators"
        starts = list(theta0)
    lo, hi = gp.local_range
    todo = range(lo, hi) if refit else [i for i in gp.get_indices_not_fit() if lo <= i < hi]
    for i in todo:
        best = _minimise_one(lambda t, i=i: gp.logposterior(i, t), lambda t, i=i: gp.logpost_deriv(i, t),
                             gp.priors[i].sample, gp.n_params[i], n_tries, starts[i], method, kwargs)
        if best is not None:
            gp.fit_emulator(i, best)
    return gp


def fit_GP_MAP(*args, n_tries=15, theta0=None, method="L-BFGS-B", skip_failures=True, refit=False, **kwargs):
    """MAP hyperparameter fit.  ``fit_GP_MAP(gp)`` with a GPU emulator, or ``fit_GP_MAP(inputs, targets, ...)``
    which builds a ``GaussianProcessGPU`` (1-D targets) or ``MultiOutputGP_GPU`` (2-D targets)."""
    if len(args) == 1:
        gp = args[0]
        if isinstance(gp, MultiOutputGP_GPU):
            gp = _fit_MOGPGPU_MAP(gp, n_tries, theta0, method, refit, **kwargs)
        elif isinstance(gp, GaussianProcessGPU):
            gp = _fit_single_GPGPU_MAP(gp, n_tries, theta0, method, **kwargs)
        else:
            raise TypeError("single arg to fit_GP_MAP must be a GaussianProcessGPU or MultiOutputGP_GPU instance")
    elif len(args) < 2:
        raise TypeError("missing required inputs/targets arrays to GaussianProcess")
    else:
        gp_kwargs = {}
        for key in ["mean", "kernel", "priors", "nugget", "inputdict", "use_patsy"]:
            if key in kwargs:
                gp_kwargs[key] = kwargs.pop(key)
        targets = np.asarray(args[1])
        if targets.ndim == 1:
            gp = _fit_single_GPGPU_MAP(GaussianProcessGPU(*args, **gp_kwargs), n_tries, theta0, method, **kwargs)
        elif targets.ndim == 2:
            gp = _fit_MOGPGPU_MAP(MultiOutputGP_GPU(*args, **gp_kwargs), n_tries, theta0, method, **kwargs)
        else:
            raise ValueError("Bad values for *args in fit_GP_MAP")
    if isinstance(gp, GaussianProcessGPU):
        if not gp.theta.data_has_been_set():
            raise RuntimeError("GP fitting failed")
    elif len(gp.get_indices_not_fit()) > 0:
        failure_string = "Fitting failed for emulators {}".format(gp.get_indices_not_fit())
        if skip_failures:
            print(failure_string)
        else:
            raise RuntimeError(failure_string)
    return gp
