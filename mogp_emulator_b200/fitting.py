"""``fit_GP_MAP`` for the GPU emulators -- the dispatch of mogp_emulator/fitting.py:16-186 restricted to
``GaussianProcessGPU`` / ``MultiOutputGP_GPU`` (fitting.py:152-155, 189-217).

The reference GPU path hands the search to a C++ dlib BFGS loop (mogp_gpu/src/fitting.hpp:61-120).  Here the
optimiser is the CPU reference's own -- ``scipy.optimize.minimize`` (L-BFGS-B) over ``logposterior`` /
``logpost_deriv`` (fitting.py:219-266) -- with every objective and gradient evaluation running on the GPU,
so results are comparable with the CPU ``GaussianProcess`` fit by fit.
"""
import numpy as np
from scipy.optimize import minimize

from . import libmogp
from .GaussianProcessGPU import GaussianProcessGPU
from .MultiOutputGP_GPU import MultiOutputGP_GPU


def _start_points(sample, n_params, n_tries, theta0):
    """The n_tries start points of one emulator: theta0 first when given, the rest drawn from the prior
    (fitting.py:237-243).  Drawn up front on the calling thread, so a multi-output search consumes the global numpy
    random stream in emulator order -- reproducible under a seed like the reference's serial loop."""
    starts = []
    for i in range(n_tries):
        if i == 0 and theta0 is not None:
            start = np.array(theta0, dtype=np.float64)
            assert start.shape == (n_params,), "theta0 must be a 1D array with length n_params"
        else:
            start = np.array(sample(), dtype=np.float64)
        starts.append(start)
    return starts


def _minimise_one(logpost, grad, starts, method, options):
    """One L-BFGS-B run per start point; a restart whose factorisation fails (kernel matrix or H^T K^-1 H not positive
    definite: ``NotPositiveDefiniteError`` / ``LinAlgError``) or whose arithmetic overflows is skipped
    (fitting.py:237-258).  Any other error -- CUDA, argument, NCCL -- propagates."""
    best_val, best_theta = None, None
    for start in starts:
        try:
            with np.errstate(divide="raise", over="raise", invalid="raise"):
                res = minimize(logpost, start, method=method, jac=grad, options=options)
        except (libmogp.NotPositiveDefiniteError, np.linalg.LinAlgError):
            print("Matrix not positive definite, skipping this iteration")
            continue
        except FloatingPointError:
            print("Floating point error in optimization routine, skipping this iteration")
            continue
        if best_val is None or res["fun"] < best_val:
            best_val, best_theta = res["fun"], res["x"]
    return best_theta


def _check_dims(d):
    if d > libmogp.GRAD_MAX_DIMS:
        raise ValueError("fit_GP_MAP on the GPU needs log-posterior gradients, which libmogp_b200 provides for at most %d "
                         "input dimensions (this emulator has %d)" % (libmogp.GRAD_MAX_DIMS, d))


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
    _check_dims(gp.D)
    best = _minimise_one(gp.logposterior, gp.logpost_deriv, _start_points(gp.priors.sample, gp.n_params, n_tries, theta0),
                         method, kwargs)
    if best is None:
        print("Minimization routine failed to return a value")
        gp.theta = None
    else:
        gp.fit(best)
    if not gp.theta.data_has_been_set():
        raise RuntimeError("Fitting did not converge")
    return gp


class _LockstepEvaluator(object):
    """Lets several ``scipy.optimize.minimize`` runs (one thread per emulator) share batched GPU evaluations.

    Every worker asks for (value, gradient) of its emulator at its current theta and blocks; when all workers that
    are still optimising have asked, the last one to arrive evaluates the whole batch with ONE call
    (``batch_fn(indices, thetas) -> {index: (value, gradient) or None}``) and wakes the others.  Each optimiser sees
    exactly the values it would see alone, so results equal the one-emulator-at-a-time loop of the reference
    (fitting.py:189-217), only faster: E concurrent factorisations fill the GPU where a single one is bound by its
    dependency chain."""

    def __init__(self, batch_fn, indices):
        import threading
        self._batch_fn = batch_fn
        self._cond = threading.Condition()
        self._active = set(indices)
        self._pending = {}
        self._results = {}
        self._error = None
        self.n_batches = 0
        self.batch_sizes = []

    def _run_batch_locked(self):
        idx = sorted(self._pending)
        thetas = [self._pending[i] for i in idx]
        self._pending = {}
        try:
            # the batch runs on whichever worker arrived last, under that thread's np.errstate(raise): give it the default
            # error state, so one emulator's host arithmetic cannot fail the whole batch (per-emulator failures come back
            # as None / an exception object per index)
            with np.errstate(divide="warn", over="warn", invalid="warn"):
                res = self._batch_fn(idx, thetas)
        except Exception as exc:            # a failure of the call itself (CUDA, memory ...): hand it to every waiter
            res = {i: exc for i in idx}
        self.n_batches += 1
        self.batch_sizes.append(len(idx))
        self._results.update(res)
        self._cond.notify_all()

    def evaluate(self, index, theta):
        with self._cond:
            self._pending[index] = np.array(theta, dtype=np.float64)
            if len(self._pending) == len(self._active):
                self._run_batch_locked()
            while index not in self._results:
                self._cond.wait()
            res = self._results.pop(index)
        if isinstance(res, Exception):
            raise res
        if res is None:
            raise libmogp.NotPositiveDefiniteError("Unable to fit the Gaussian process: matrix not positive definite")
        return res

    def done(self, index):
        """The worker of ``index`` has no more requests: the remaining ones must not wait for it."""
        with self._cond:
            self._active.discard(index)
            if self._pending and len(self._pending) == len(self._active):
                self._run_batch_locked()


def _fit_MOGPGPU_MAP(gp, n_tries=15, theta0=None, method="L-BFGS-B", refit=False, **kwargs):
    import threading
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
    todo = list(range(lo, hi)) if refit else [i for i in gp.get_indices_not_fit() if lo <= i < hi]
    if not todo:
        return gp
    _check_dims(gp.D)
    priors = {i: gp.priors[i] for i in todo}       # built on the calling thread
    # start points drawn here, in emulator order: reproducible under np.random.seed (the workers run concurrently)
    start_points = {i: _start_points(priors[i].sample, gp.n_params[i], n_tries, starts[i]) for i in todo}
    evaluator = _LockstepEvaluator(gp.logpost_and_deriv_batch, todo)
    best = {}
    errors = []

    def worker(i):
        try:
            def fun(theta):
                val, grad = evaluator.evaluate(i, theta)
                return val, grad
            best[i] = _minimise_one(fun, True, start_points[i], method, kwargs)
        except BaseException as exc:     # noqa: BLE001 - reported on the calling thread
            errors.append((i, exc))
        finally:
            evaluator.done(i)

    threads = [threading.Thread(target=worker, args=(i,), name="mogp-map-%d" % i) for i in todo]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0][1]
    # final state: every emulator that converged sits at its best theta (one more batched fit)
    good = [i for i in todo if best.get(i) is not None]
    if good:
        res = gp.logpost_and_deriv_batch(good, [best[i] for i in good])
        del res
    for i in todo:
        if best.get(i) is None:
            gp.reset_emulator(i)
    gp.sync_fit_status()         # sharded emulators: every rank learns which outputs of the other ranks converged
    gp.map_fit_stats = {"batches": evaluator.n_batches, "mean_batch": float(np.mean(evaluator.batch_sizes or [0]))}
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
