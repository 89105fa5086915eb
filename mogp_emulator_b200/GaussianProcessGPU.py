"""``GaussianProcessGPU`` -- drop-in for the reference class of the same name
(mogp_emulator/GaussianProcessGPU.py:208-667), backed by libmogp_b200 through ctypes.

Values follow the reference's CPU ``GaussianProcess`` (GaussianProcess.py:629-927): adaptive jitter
schedule of linalg/cholesky.py:234-281, variance clipped at zero, priors added to the log-posterior on the
host.  Mean functions: zero, constant and formula means (``formula.py``), coefficients integrated out analytically as in
the CPU reference.
"""
import warnings

import numpy as np

from . import libmogp
from .hyper import GPParams, GPPriors, make_priors
from .kernels import SquaredExponential, Matern52, interpret_kernel
from .meanfunc import interpret_mean, design_matrix, design_matrix_inputderiv, MeanFit


class GPUUnavailableError(RuntimeError):
    """Raised for features the GPU emulator does not provide (reference: GaussianProcess.py GPUUnavailableError)."""


class PredictResult(dict):
    """Prediction container with keys/attributes ``mean``, ``unc``, ``deriv`` that also unpacks and
    indexes as the tuple ``(mean, unc, deriv)`` (reference: GaussianProcess.py:948-1026)."""

    _ORDER = ("mean", "unc", "deriv")

    def __getattr__(self, name):
        try:
            return dict.__getitem__(self, name)
        except KeyError:
            raise AttributeError(name)

    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__

    def __getitem__(self, key):
        if isinstance(key, int) and not isinstance(key, bool):
            if not 0 <= key < 3:
                raise KeyError(key)
            key = self._ORDER[key]
        elif not isinstance(key, str):
            raise KeyError(key)
        return dict.__getitem__(self, key)

    def __iter__(self):
        return iter([dict.__getitem__(self, k) for k in self._ORDER])

    def __len__(self):
        return 3

    def __repr__(self):
        return "\n".join("%6s: %r" % (k, dict.get(self, k)) for k in self._ORDER)


def interpret_nugget(nugget):
    """``"adaptive"`` / ``"fit"`` / non-negative number -> (nugget_type, size); TypeError / ValueError as the
    reference (GaussianProcessGPU.py:112-141)."""
    if isinstance(nugget, str):
        if nugget == "adaptive":
            return libmogp.nugget_type.adaptive, 0.0
        if nugget == "fit":
            return libmogp.nugget_type.fit, 0.0
        raise ValueError("nugget must be a string set to 'adaptive', 'fit', or a float")
    try:
        value = float(nugget)
    except (TypeError, ValueError):
        raise TypeError("nugget parameter must be a string or a non-negative float")
    if value < 0.0:
        raise ValueError("nugget parameter must be non-negative")
    return libmogp.nugget_type.fixed, value


def _check_mean(mean):
    """-> None (zero mean) or a MeanFormula; ValueError for anything that is not a valid formula (see meanfunc.py)."""
    return interpret_mean(mean)


class GaussianProcessGPU(object):
    """Single-output GP emulator on one B200."""

    def __init__(self, inputs, targets, mean=None, kernel=SquaredExponential(), priors=None, nugget="adaptive",
                 inputdict={}, use_patsy=True, max_batch_size=2000, device=0):
        if not libmogp.HAVE_LIBMOGP:
            raise RuntimeError("Cannot construct GaussianProcessGPU: the GPU library (libmogp_b200) could not be "
                               "loaded: " + libmogp.last_error())
        if not libmogp.gpu_usable():
            raise RuntimeError("Cannot construct GaussianProcessGPU: a compatible GPU could not be found")
        inputs = libmogp.as_f64(inputs)
        if inputs.ndim == 1:
            inputs = np.reshape(inputs, (-1, 1))
        assert inputs.ndim == 2
        targets = libmogp.as_f64(targets)
        assert targets.ndim == 1
        assert targets.shape[0] == inputs.shape[0]
        self._mean_spec = _check_mean(mean)
        self._inputs = inputs
        self._targets = targets
        self._dm = design_matrix(self._mean_spec, inputs)
        self._meanfit = None
        self._max_batch_size = max_batch_size    # kept for signature compatibility; predict is chunked natively
        self._device = int(device)
        self.mean = None if self._mean_spec is None else str(self._mean_spec)
        if inputdict:
            warnings.warn("The inputdict interface for mean functions has been deprecated. You must input your mean "
                          "formulae using the x[0] format directly in the formula.", DeprecationWarning)
        self.kernel_type, self.kernel = interpret_kernel(kernel)
        self._nugget_type, self._init_nugget_size = interpret_nugget(nugget)
        self._priors_arg = priors
        self._handle = None
        self._init_gpu()
        self._set_priors(priors)

    # -- device object ------------------------------------------------------------------------------
    def _init_gpu(self):
        if self._handle is None:
            self._handle = libmogp.Handle(self._inputs, self._targets.reshape(1, -1), self.kernel_type,
                                          self._nugget_type, self._init_nugget_size, device=self._device)
            self._theta = GPParams(self.D, self.nugget_type,
                                   self._init_nugget_size if self.nugget_type == "fixed" else None)
            self._logpost_data = None

    def _set_priors(self, newpriors=None):
        """Priors are host-side scalars that only enter logposterior / logpost_deriv; the default ones need a
        root find per input dimension (Priors.py:698-760), so they are built on first use."""
        self._priors_arg = newpriors
        self._priors = None if newpriors is None else make_priors(newpriors, self._inputs, self.n_corr, self.nugget_type)

    # -- properties (GaussianProcessGPU.py:335-502) ---------------------------------------------------
    @property
    def priors(self):
        if self._priors is None:
            self._priors = make_priors(None, self._inputs, self.n_corr, self.nugget_type)
        return self._priors

    @property
    def inputs(self):
        return self._inputs

    @property
    def targets(self):
        return self._targets

    @property
    def n(self):
        return self._inputs.shape[0]

    @property
    def D(self):
        return self._inputs.shape[1]

    @property
    def n_corr(self):
        return self.D

    @property
    def n_params(self):
        return self._theta.get_n_data() + self._theta.get_n_mean()

    @property
    def nugget_type(self):
        return self._nugget_type.name

    @property
    def nugget(self):
        if self.nugget_type == "fixed":
            return self._init_nugget_size
        nug = self._theta.nugget
        return 0.0 if nug is None else nug

    @nugget.setter
    def nugget(self, nugget):
        new_type, new_size = interpret_nugget(nugget)
        theta_was = self._theta.get_data() if self._theta.data_has_been_set() else None
        changed_shape = (new_type == libmogp.nugget_type.fit) != (self._nugget_type == libmogp.nugget_type.fit)
        self._nugget_type, self._init_nugget_size = new_type, new_size
        # the device object is tied to the nugget treatment: rebuild it (and the default priors)
        self._handle.close()
        self._handle = None
        self._init_gpu()
        self._set_priors(self._priors_arg)
        if theta_was is not None and not changed_shape:
            self.fit(theta_was)

    @property
    def theta(self):
        return self._theta

    @theta.setter
    def theta(self, theta):
        if theta is None:
            self._handle.reset(0)
            self._theta.unset_data()
            self._logpost_data = None
            self._meanfit = None
        else:
            self.fit(theta)

    @property
    def L(self):
        if not self._theta.data_has_been_set():
            return None
        return self._handle.get(0, libmogp.GET_L)

    @property
    def n_mean(self):
        return self._dm.shape[1]

    def get_design_matrix(self, inputs):
        return design_matrix(self._mean_spec, inputs)

    @property
    def Kinv_t(self):
        """K^-1 (y - m), with m = 0 for the weak mean priors (GaussianProcess.py:666)."""
        if not self._theta.data_has_been_set():
            return None
        return self._handle.get(0, libmogp.GET_ALPHA) if self._meanfit is None else self._Kinv_t_host

    @property
    def Kinv_t_mean(self):
        """K^-1 (y - H beta) (GaussianProcess.py:672)."""
        if not self._theta.data_has_been_set():
            return None
        return self._handle.get(0, libmogp.GET_ALPHA)

    @property
    def current_logpost(self):
        if not self._theta.data_has_been_set():
            return None
        return self._logpost_data - self.priors.logp(self._theta)

    @property
    def invQ(self):
        """(K + nugget I)^-1 as a dense (n, n) array -- ``DenseGP_GPU::get_invQ`` of the reference's native class
        (mogp_gpu/src/densegp_gpu.hpp:629).  Computed on demand from the Cholesky factor; nothing else in this library forms it."""
        if not self._theta.data_has_been_set():
            return None
        return self._handle.get(0, libmogp.GET_KINV)

    def get_K_matrix(self):
        """sigma^2 * k(X, X) without the nugget (GaussianProcess.get_K_matrix, GaussianProcess.py:545-558)."""
        if not self._theta.data_has_been_set():
            raise ValueError("hyperparameters have not been fit for this Gaussian Process")
        return self._handle.get(0, libmogp.GET_K)

    # -- fitting --------------------------------------------------------------------------------------
    def fit(self, theta):
        """Set the hyperparameters and factorise (GaussianProcess.fit, GaussianProcess.py:629-685).  A wrong
        length raises RuntimeError like the reference GPU class (densegp_gpu.hpp:495-496); a matrix that is
        not positive definite raises RuntimeError (densegp_gpu.hpp:556-570)."""
        if isinstance(theta, GPParams):
            theta = theta.get_data()
        theta = libmogp.as_f64(theta).reshape(-1)
        if theta.shape != (self.n_params,):
            raise RuntimeError("bad shape for hyperparameters: expected %d values, got %d" % (self.n_params, theta.size))
        try:
            quad, logdet, nug, status = self._handle.fit(0, theta)
        except FloatingPointError:
            # infinite squared distance (calc_r2, Kernel.py:482-483): the emulator is left "not fit"
            self._theta.unset_data()
            self._logpost_data = None
            self._meanfit = None
            raise
        self.n_fit_calls = getattr(self, "n_fit_calls", 0) + 1
        self._meanfit = None
        if status[0] != libmogp.OK:
            self._theta.unset_data()
            self._logpost_data = None
            raise libmogp.NotPositiveDefiniteError("Unable to fit the Gaussian process: matrix not positive definite" +
                                                   (", even with jitter." if self.nugget_type == "adaptive" else ""))
        mf = None
        if self.n_mean > 0:
            # analytic mean (GaussianProcess.py:657-685): K^-1 H on the device, the n_mean x n_mean algebra here, then the
            # device's alpha becomes K^-1 (y - H beta) and it learns the rank-n_mean correction of the gradient
            self._Kinv_t_host = self._handle.get(0, libmogp.GET_ALPHA)
            W = np.column_stack([self._handle.solve_list([0], self._dm[:, q])[0] for q in range(self.n_mean)])
            try:
                mf = MeanFit(self._dm, self._targets, self._Kinv_t_host, W, self.n)
            except np.linalg.LinAlgError:
                # H^T K^-1 H is numerically not positive definite (the reference's calc_Ainv raises LinAlgError, which its MAP
                # loop skips, fitting.py:244-249): the emulator stays "not fit" on the host and on the device
                self._theta.unset_data()
                self._logpost_data = None
                self._handle.reset(0)
                raise libmogp.NotPositiveDefiniteError("Unable to fit the Gaussian process: mean-function matrix H^T K^-1 H "
                                                       "not positive definite")
        self._theta.set_data(theta)
        self._theta.nugget = float(nug[0])
        if mf is None:
            self._logpost_data = 0.5 * (float(quad[0]) + float(logdet[0]) + self.n * np.log(2.0 * np.pi))
        else:
            self._handle.set_alpha_list([0], mf.alpha_mean)
            self._handle.set_mean_vectors_list([0], mf.U.T[np.newaxis])
            self._meanfit = mf
            self._theta.mean = mf.beta.copy()
            self._logpost_data = mf.data_logpost(float(quad[0]), float(logdet[0]), self.n)

    def _refit(self, theta):
        return (not self._theta.data_has_been_set()
                or not np.allclose(theta, self._theta.get_data(), rtol=1.0e-10, atol=1.0e-15))

    def logposterior(self, theta):
        """Negative log-posterior (GaussianProcess.logposterior, GaussianProcess.py:688-709)."""
        theta = libmogp.as_f64(theta).reshape(-1)
        if self._refit(theta):
            self.fit(theta)
        return self.current_logpost

    def logpost_deriv(self, theta):
        """Gradient of the negative log-posterior (GaussianProcess.logpost_deriv, GaussianProcess.py:711-782)."""
        theta = libmogp.as_f64(theta).reshape(-1)
        assert theta.shape == (self.n_params,), "bad shape for new parameters"
        if self._refit(theta):
            self.fit(theta)
        grad = self._handle.logpost_grad(0, self.n_params)
        self.n_grad_calls = getattr(self, "n_grad_calls", 0) + 1
        return grad - self.priors.dlogpdtheta(self._theta)

    def logpost_hessian(self, theta):
        raise GPUUnavailableError("The Hessian calculation is not currently implemented in the GPU version of MOGP.")

    # -- prediction -------------------------------------------------------------------------------------
    def predict(self, testing, unc=True, deriv=True, include_nugget=True, full_cov=False):
        """Posterior mean / variance at ``testing`` (GaussianProcess.predict, GaussianProcess.py:818-927) and, with
        ``deriv=True`` (the reference GPU class's default, GaussianProcessGPU.py:582), the derivative of the mean
        with respect to the test inputs, shape ``(m, D)`` (DenseGP_GPU::predict_deriv, densegp_gpu.hpp:411-448)."""
        if not self._theta.data_has_been_set():
            raise ValueError("hyperparameters have not been fit for this Gaussian Process")
        testing = libmogp.as_f64(testing)
        if self.D == 1 and testing.ndim == 1:
            testing = np.reshape(testing, (-1, 1))
        elif testing.ndim == 1:
            testing = np.reshape(testing, (1, len(testing)))
        assert testing.ndim == 2
        assert testing.shape[1] == self.D
        dmean = self._handle.predict_deriv(testing)[0][0] if deriv else None
        mf = self._meanfit
        if deriv and mf is not None and self._mean_spec.terms:
            # the mean function's share: d(H* beta)/dx* (densegp_gpu.hpp:411-448 adds mean_inputderiv the same way)
            dmean = dmean + np.dot(design_matrix_inputderiv(self._mean_spec, testing), mf.beta)
        mshift = 0.0 if mf is None else np.dot(self.get_design_matrix(testing), mf.beta)
        extra = None
        if mf is not None and unc:
            # R^T A^-1 R, R = H*^T - H^T K^-1 K*  (GaussianProcess.py:897-920)
            HtKinvKs = self._handle.kstar_dot(testing, mf.W.T[np.newaxis])[0]
            extra = mf.variance_term(self.get_design_matrix(testing), HtKinvKs, full_cov=full_cov)
        if unc and full_cov:
            # the CPU class's full_cov=True (GaussianProcess.py:899-911): (m, m) covariance, not clipped
            mean1, cov = self._handle.predict_cov(0, testing, include_nugget=include_nugget)
            return PredictResult(mean=mean1 + mshift, unc=cov if extra is None else cov + extra, deriv=dmean)
        if extra is None:
            mean, var, _ = self._handle.predict(testing, want_var=unc, include_nugget=include_nugget)
            return PredictResult(mean=mean[0] + mshift, unc=(var[0] if unc else None), deriv=dmean)
        mean, var, _ = self._handle.predict(testing, want_var=2, include_nugget=include_nugget)   # variance before the clip
        return PredictResult(mean=mean[0] + mshift, unc=np.maximum(var[0] + extra, 0.0), deriv=dmean)

    def __call__(self, testing):
        return self.predict(testing, unc=False, deriv=False)[0]

    def close(self):
        """Release the device object now (its buffers go back to the library's cache) instead of at garbage collection."""
        if self._handle is not None:
            self._handle.close()
            self._handle = None

    def __str__(self):
        return ("Gaussian Process with " + str(self.n) + " training examples and " + str(self.D) + " input variables")

    # -- pickling: drop the device object, rebuild and refit on load (GaussianProcessGPU.py:656-667) -----
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_handle"] = None
        state["_saved_theta"] = self._theta.get_data() if self._theta.data_has_been_set() else None
        return state

    def __setstate__(self, state):
        saved = state.pop("_saved_theta", None)
        self.__dict__ = state
        self._handle = None
        self._init_gpu()
        if saved is not None:
            self.fit(saved)
