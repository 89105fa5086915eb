"""Host-side hyperparameter container and priors for the GPU emulators.

Scalar support code for the hot path (never on the GPU): the raw <-> scaled transforms of the reference's
``GPParams`` (mogp_emulator/GPParams.py:3-161, 215-555; C++ twin mogp_gpu/src/gpparams.hpp) and the prior
terms that enter ``current_logpost`` and its gradient (mogp_emulator/Priors.py:86-152, 291-418, 583-1188;
C++ twin mogp_gpu/src/gppriors.hpp).  Raw layout: ``[theta_corr (D), theta_cov, (theta_nugget)]`` with
correlation length ``l = exp(-theta/2)``, ``sigma^2 = exp(theta_cov)``, ``nugget = exp(theta_nugget)``.
Mean parameters are integrated out analytically (meanfunc.py), never part of the hyperparameter vector, so ``n_mean``
(the reference GPU binding's count of mean entries inside theta) is always 0; ``mean`` holds the fitted coefficients.
"""
import numpy as np
import scipy.stats
from scipy.optimize import root
from scipy.special import gammaln


# --------------------------------------------------------------------------------------------------
# transforms (GPParams.py:3-161)
# --------------------------------------------------------------------------------------------------
class CorrTransform(object):
    """raw r <-> correlation length l = exp(-r/2)."""

    @staticmethod
    def transform(r):
        return np.exp(-0.5 * np.asarray(r, dtype=np.float64))

    @staticmethod
    def inv_transform(l):
        return -2.0 * np.log(l)

    @staticmethod
    def dscaled_draw(l):
        return -0.5 * l

    @staticmethod
    def d2scaled_draw2(l):
        return 0.25 * l


class CovTransform(object):
    """raw r <-> scaled s = exp(r) (covariance scale and fitted nugget)."""

    @staticmethod
    def transform(r):
        return np.exp(np.asarray(r, dtype=np.float64))

    @staticmethod
    def inv_transform(s):
        return np.log(s)

    @staticmethod
    def dscaled_draw(s):
        return s

    @staticmethod
    def d2scaled_draw2(s):
        return s


class GPParams(object):
    """What ``gp.theta`` returns: answers the queries the reference front-end and its tests make of the
    C++ ``GPParameters`` binding (``get_data``, ``data_has_been_set``, ``get_n_data``, ``get_n_mean``;
    GaussianProcessGPU.py:406-407, 557, 589; tests/test_GaussianProcess.py:616-619) plus the scaled views
    of the CPU class (``corr``, ``cov``, ``nugget``)."""

    def __init__(self, n_corr, nugget_type, nugget=None):
        assert nugget_type in ("adaptive", "fit", "fixed")
        self.n_mean = 0
        self.mean = np.zeros(0)          # analytic mean coefficients of the last fit (GaussianProcess.py:670-671)
        self.n_corr = int(n_corr)
        self.nugget_type = nugget_type
        self._nugget = None if nugget_type != "fixed" else float(nugget)
        self._data = np.zeros(self.n_data)
        self._set = False

    @property
    def n_data(self):
        return self.n_corr + 1 + int(self.nugget_type == "fit")

    n_params = n_data

    def get_n_data(self):
        return self.n_data

    def get_n_mean(self):
        return 0

    def data_has_been_set(self):
        return self._set

    def get_data(self):
        return self._data.copy()

    def set_data(self, data):
        data = np.array(data, dtype=np.float64).reshape(-1)
        assert data.shape == (self.n_data,), "bad shape for hyperparameters"
        self._data = data
        self._set = True
        if self.nugget_type == "fit":
            self._nugget = float(np.exp(data[-1]))
        elif self.nugget_type == "adaptive":
            self._nugget = None

    def unset_data(self):
        """Reference GPU semantics: data zeroed, flag cleared (gpparams.hpp:217-221)."""
        self._data = np.zeros(self.n_data)
        self._set = False
        self.mean = np.zeros(0)
        if self.nugget_type != "fixed":
            self._nugget = None

    @property
    def corr_raw(self):
        return self._data[:self.n_corr]

    @property
    def corr(self):
        return CorrTransform.transform(self.corr_raw)

    @property
    def cov(self):
        return float(np.exp(self._data[self.n_corr]))

    @property
    def nugget(self):
        return self._nugget

    @nugget.setter
    def nugget(self, value):
        self._nugget = None if value is None else float(value)

    def __array__(self, dtype=None, copy=None):
        return np.array(self._data, dtype=dtype)

    def __len__(self):
        return self.n_data

    def __str__(self):
        return "GPParams(corr_raw=%s, cov_raw=%s, nugget=%s [%s])" % (
            self.corr_raw, self._data[self.n_corr], self._nugget, self.nugget_type)


# --------------------------------------------------------------------------------------------------
# prior distributions (Priors.py:583-1149)
# --------------------------------------------------------------------------------------------------
class WeakPrior(object):
    """Improper flat prior: contributes nothing; samples uniformly in [-2.5, 2.5) in raw space."""

    def logp(self, x):
        return 0.0

    def dlogpdx(self, x):
        return 0.0

    def dlogpdtheta(self, x, transform):
        return float(self.dlogpdx(x) * transform.dscaled_draw(x))

    def sample(self, transform=None):
        return float(5.0 * (np.random.rand() - 0.5))


class PriorDist(WeakPrior):
    """Proper prior on the scaled parameter; two-parameter families built from a (lo, hi) 99% interval."""

    def __init__(self, shape, scale):
        assert shape > 0.0 and scale > 0.0, "shape and scale must be positive"
        self.shape = float(shape)
        self.scale = float(scale)

    def _frozen(self):
        raise NotImplementedError

    @classmethod
    def _make(cls, shape, scale):
        return cls(shape, scale)

    @classmethod
    def default_prior(cls, min_val, max_val):
        """Distribution with 0.5% of its mass below min_val and 0.5% above max_val (root find in log
        space from zeros, Priors.py:698-760); a WeakPrior when the solver fails."""
        if min_val <= 0.0 or max_val <= 0.0 or not max_val > min_val:
            return WeakPrior()

        def f(x):
            cdf = cls._make(np.exp(x[0]), np.exp(x[1]))._frozen().cdf
            return np.array([cdf(min_val) - 0.005, cdf(max_val) - 0.995])

        with np.errstate(all="ignore"):
            res = root(f, np.zeros(2))
        if not res["success"]:
            return WeakPrior()
        return cls._make(float(np.exp(res["x"][0])), float(np.exp(res["x"][1])))

    @classmethod
    def default_prior_corr(cls, inputs):
        lo, hi = min_spacing(inputs), max_spacing(inputs)
        if lo == 0.0 or hi == 0.0:
            return WeakPrior()
        return cls.default_prior(lo, hi)

    def sample_x(self):
        return float(np.atleast_1d(self._frozen().rvs(size=1))[0])

    def sample(self, transform):
        return float(transform.inv_transform(self.sample_x()))


class InvGammaPrior(PriorDist):
    def _frozen(self):
        return scipy.stats.invgamma(self.shape, scale=self.scale)

    def logp(self, x):
        return float(self.shape * np.log(self.scale) - gammaln(self.shape) - (self.shape + 1.0) * np.log(x)
                     - self.scale / x)

    def dlogpdx(self, x):
        return float(-(self.shape + 1.0) / x + self.scale / x ** 2)

    @classmethod
    def default_prior_mode(cls, min_val, max_val):
        """Mode at the geometric mean of the interval, 99.5% of the mass below max_val
        (Priors.py:1013-1060)."""
        mode = np.sqrt(min_val * max_val)

        def f(x):
            a = np.exp(x)
            return scipy.stats.invgamma(a, scale=(1.0 + a) * mode).cdf(max_val) - 0.995

        with np.errstate(all="ignore"):
            res = root(f, 0.0)
        if not res["success"]:
            return WeakPrior()
        a = float(np.exp(np.atleast_1d(res["x"])[0]))
        return cls(a, (1.0 + a) * mode)

    @classmethod
    def default_prior_corr_mode(cls, inputs):
        lo, hi = min_spacing(inputs), max_spacing(inputs)
        if lo == 0.0 or hi == 0.0:
            return WeakPrior()
        return cls.default_prior_mode(lo, hi)

    @classmethod
    def default_prior_nugget(cls, min_val=1.0e-8, max_val=1.0e-6):
        return cls.default_prior_mode(min_val, max_val)


class GammaPrior(PriorDist):
    def _frozen(self):
        return scipy.stats.gamma(self.shape, scale=self.scale)

    def logp(self, x):
        return float(-self.shape * np.log(self.scale) - gammaln(self.shape) + (self.shape - 1.0) * np.log(x)
                     - x / self.scale)

    def dlogpdx(self, x):
        return float((self.shape - 1.0) / x - 1.0 / self.scale)


class LogNormalPrior(PriorDist):
    def _frozen(self):
        return scipy.stats.lognorm(self.shape, scale=self.scale)

    def logp(self, x):
        return float(-0.5 * (np.log(x / self.scale) / self.shape) ** 2 - 0.5 * np.log(2.0 * np.pi) - np.log(x)
                     - np.log(self.shape))

    def dlogpdx(self, x):
        return float(-np.log(x / self.scale) / self.shape ** 2 / x - 1.0 / x)


def min_spacing(v):
    """Median gap between distinct sorted values (Priors.py:1170-1188)."""
    v = np.unique(np.array(v).flatten())
    if len(v) <= 2:
        return 0.0
    return float(np.median(np.diff(np.sort(v))))


def max_spacing(v):
    """Range of the distinct values (Priors.py:1151-1168)."""
    v = np.unique(np.array(v).flatten())
    if len(v) <= 1:
        return 0.0
    return float(v.max() - v.min())


class GPPriors(object):
    """Priors on [corr (D), cov, (nugget)]; weak wherever None is given (Priors.py:9-84)."""

    def __init__(self, corr=None, cov=None, nugget=None, n_corr=None, nugget_type="fit", mean=None):
        if mean is not None:
            raise ValueError("informative mean-function priors are not supported by the GPU emulator (weak priors only)")
        assert nugget_type in ("adaptive", "fit", "fixed"), "Bad value for nugget type in GPPriors"
        if corr is None:
            assert n_corr is not None and n_corr > 0, "need n_corr when no correlation priors are given"
            corr = [None] * int(n_corr)
        self.corr = [p if p is not None else WeakPrior() for p in corr]
        for p in self.corr:
            if not isinstance(p, WeakPrior):
                raise TypeError("correlation priors must be prior distribution objects or None")
        self.cov = cov if cov is not None else WeakPrior()
        self.nugget_type = nugget_type
        if nugget_type == "fit":
            self.nugget = nugget if nugget is not None else WeakPrior()
        else:
            self.nugget = None

    @property
    def n_corr(self):
        return len(self.corr)

    @classmethod
    def default_priors(cls, inputs, n_corr, nugget_type="fit", dist="invgamma"):
        """Priors.py:86-152: per-dimension priors from the input spacing, InvGamma mode-based fallback,
        default small-nugget prior when the nugget is fitted, weak covariance prior."""
        families = {"invgamma": InvGammaPrior, "gamma": GammaPrior, "lognormal": LogNormalPrior}
        if not isinstance(dist, str) or dist.lower() not in families:
            raise TypeError("dist must be 'invgamma', 'gamma' or 'lognormal'")
        fam = families[dist.lower()]
        inputs = np.asarray(inputs, dtype=np.float64)
        if inputs.shape[1] == n_corr:
            columns = np.transpose(inputs)
        elif n_corr == 1:
            columns = np.reshape(inputs, (1, -1))
        else:
            raise ValueError("Number of correlation lengths not compatible with input array")
        corr = []
        for col in columns:
            p = fam.default_prior_corr(col)
            if not isinstance(p, fam):
                p = InvGammaPrior.default_prior_corr_mode(col)
            corr.append(p)
        nug = InvGammaPrior.default_prior_nugget() if nugget_type == "fit" else None
        return cls(corr=corr, cov=None, nugget=nug, nugget_type=nugget_type)

    def logp(self, theta):
        total = 0.0
        for p, l in zip(self.corr, theta.corr):
            total += p.logp(l)
        total += self.cov.logp(theta.cov)
        if self.nugget_type == "fit":
            total += self.nugget.logp(theta.nugget)
        return float(total)

    def dlogpdtheta(self, theta):
        out = [p.dlogpdtheta(l, CorrTransform) for p, l in zip(self.corr, theta.corr)]
        out.append(self.cov.dlogpdtheta(theta.cov, CovTransform))
        if self.nugget_type == "fit":
            out.append(self.nugget.dlogpdtheta(theta.nugget, CovTransform))
        return np.array(out)

    def sample(self):
        pt = [p.sample(CorrTransform) for p in self.corr]
        pt.append(self.cov.sample(CovTransform))
        if self.nugget_type == "fit":
            pt.append(self.nugget.sample(CovTransform))
        return np.array(pt)


def make_priors(priors, inputs, n_corr, nugget_type):
    """None -> default priors; GPPriors -> as is; dict -> GPPriors(**dict)
    (create_prior_params, GaussianProcessGPU.py:143-205)."""
    if priors is None:
        return GPPriors.default_priors(inputs, n_corr, nugget_type)
    if isinstance(priors, GPPriors):
        return priors
    if isinstance(priors, dict):
        try:
            kw = dict(priors)
            kw.setdefault("n_corr", n_corr)
            kw.setdefault("nugget_type", nugget_type)
            return GPPriors(**kw)
        except TypeError:
            raise TypeError("Provided arguments for priors are not valid inputs for a GPPriors object.")
    raise TypeError("priors must be a GPPriors object, a dict of GPPriors arguments, or None")
