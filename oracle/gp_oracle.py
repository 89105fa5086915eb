"""CPU oracle for the GP fit+predict hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy/scipy restatement of the reference's CPU algorithm (alan-turing-institute/mogp-emulator
v0.7.2) for the one path this repository accelerates: kernel-matrix assembly, nugget-regularised
Cholesky, triangular solves, log-posterior (+ gradient) and the fan-out over outputs.  Every
function cites the reference file:line it follows (paths relative to /root/reference).

Who may import this module: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg -- only as the checker or the timed CPU baseline.  The product package
(mogp_emulator_b200) never imports it and has no CPU fallback.

Pinning status: PINNED.  tests/test_oracle.py checks this restatement against
  (a) the known-answer vectors the reference's own tests hold for this path
      (mogp_emulator/tests/test_Kernel.py:8-38, 721-754, 975-987; tests/test_linalg.py:103-154;
      tests/test_GaussianProcess.py:1144-1161), re-expressed in tests/golden/known_answers.py, and
  (b) outputs of the unmodified reference itself, generated in the build container by
      tests/golden/make_golden.py (committed as tests/golden/*.npz).

Third-party arithmetic: the reference delegates the factorisation and solves to LAPACK through
SciPy (scipy.linalg.cholesky / lapack.dpotrf / cho_solve, reference requirement ``scipy>=1.4``,
un-vendored).  The oracle calls the same SciPy entry points, so it inherits the same LAPACK.
"""

import numpy as np
import scipy.linalg
from scipy.linalg import lapack
from scipy.optimize import root
from scipy.special import gammaln
import scipy.stats

SQEXP = "SquaredExponential"
MAT52 = "Matern52"


# ------------------------------------------------------------------------------------------------
# kernel matrix assembly
# ------------------------------------------------------------------------------------------------

def calc_r2(x1, x2, theta_corr):
    """Scaled squared distances r2[i,j] = sum_d exp(theta_d) (x1[i,d]-x2[j,d])**2.

    Follows StationaryKernel.calc_r2, mogp_emulator/Kernel.py:476-485 verbatim in arithmetic,
    including the (n1, n2, D) broadcast temporary at :480 (this is the reference's CPU hot spot,
    so the timed baseline keeps it) and the FloatingPointError on inf at :482-483.
    """
    x1 = np.atleast_2d(np.asarray(x1, dtype=np.float64))
    x2 = np.atleast_2d(np.asarray(x2, dtype=np.float64))
    exp_theta = np.exp(np.asarray(theta_corr, dtype=np.float64))
    r2 = np.sum(exp_theta * (x1[:, np.newaxis, :] - x2[np.newaxis, :, :]) ** 2, axis=-1)
    if np.any(np.isinf(r2)):
        raise FloatingPointError("Inf enountered in kernel distance computation")
    return r2


def calc_r2_chunked(x1, x2, theta_corr, rows=512):
    """Same arithmetic as calc_r2 evaluated ``rows`` rows of x1 at a time.

    The reference cannot allocate its (n1,n2,D) temporary at n=16384,d=20 (42.9 GB,
    SURVEY.md section 6.2); this row-chunked form performs the identical per-element operations
    in the identical order (the reduction over d is per element), so results are bit-identical
    to calc_r2 wherever both run.
    """
    x1 = np.atleast_2d(np.asarray(x1, dtype=np.float64))
    x2 = np.atleast_2d(np.asarray(x2, dtype=np.float64))
    out = np.empty((x1.shape[0], x2.shape[0]))
    for s in range(0, x1.shape[0], rows):
        out[s:s + rows] = calc_r2(x1[s:s + rows], x2, theta_corr)
    return out


def calc_K(r2, kernel=SQEXP):
    """K(r2): SqExpBase.calc_K Kernel.py:787-791, Mat52Base.calc_K Kernel.py:878-882."""
    assert np.all(r2 >= 0.0), "kernel distances must be positive"
    if kernel == SQEXP:
        return np.exp(-0.5 * r2)
    if kernel == MAT52:
        return (1.0 + np.sqrt(5.0 * r2) + 5.0 / 3.0 * r2) * np.exp(-np.sqrt(5.0 * r2))
    raise ValueError("kernel must be SquaredExponential or Matern52")


def calc_dKdr2(r2, kernel=SQEXP):
    """dK/dr2: SqExpBase.calc_dKdr2 Kernel.py:793-814, Mat52Base.calc_dKdr2 Kernel.py:884-906."""
    if kernel == SQEXP:
        return -0.5 * np.exp(-0.5 * r2)
    if kernel == MAT52:
        return -5.0 / 6.0 * (1.0 + np.sqrt(5.0 * r2)) * np.exp(-np.sqrt(5 * r2))
    raise ValueError("kernel must be SquaredExponential or Matern52")


def kernel_f(x1, x2, theta_corr, kernel=SQEXP, chunked=False):
    """KernelBase.kernel_f, Kernel.py:129-131 (correlation only; sigma^2 applied by the caller)."""
    r2 = calc_r2_chunked(x1, x2, theta_corr) if chunked else calc_r2(x1, x2, theta_corr)
    return calc_K(r2, kernel)


# ------------------------------------------------------------------------------------------------
# nugget-regularised Cholesky
# ------------------------------------------------------------------------------------------------

def _check_cholesky_inputs(A):
    """linalg/cholesky.py:196-222: 2-D square, symmetric (assert_allclose), positive diagonal."""
    A = np.array(A)
    assert A.ndim == 2 and A.shape[0] == A.shape[1], "A must have shape (n,n)"
    np.testing.assert_allclose(A.T, A)
    if np.any(np.diag(A) <= 0.0):
        raise scipy.linalg.LinAlgError("not pd: non-positive diagonal elements")
    return A


def fixed_cholesky(A):
    """linalg/cholesky.py:225-231."""
    A = _check_cholesky_inputs(A)
    return scipy.linalg.cholesky(A, lower=True)


def jit_cholesky(A, maxtries=5):
    """Adaptive-jitter Cholesky, linalg/cholesky.py:234-281.

    Plain dpotrf first; on info != 0 the jitter starts at mean(diag)*1e-6 and is multiplied by
    10 after each failure, at most ``maxtries`` jittered attempts.  Returns (L, jitter) with
    jitter == 0.0 when the plain factorisation succeeded.
    """
    A = _check_cholesky_inputs(A)
    A = np.ascontiguousarray(A)
    L, info = lapack.dpotrf(A, lower=1)
    if info == 0:
        return L, 0.0
    jitter = np.diag(A).mean() * 1e-6
    num_tries = 1
    while num_tries <= maxtries and np.isfinite(jitter):
        try:
            L = scipy.linalg.cholesky(A + np.eye(A.shape[0]) * jitter, lower=True)
            return L, jitter
        except Exception:
            jitter *= 10
        finally:
            num_tries += 1
    raise scipy.linalg.LinAlgError("not positive definite, even with jitter.")


def cholesky_factor(A, nugget, nugget_type):
    """linalg/cholesky.py:168-193 for the nugget types the GPU API accepts (no "pivot")."""
    if nugget_type == "adaptive":
        L, nugget = jit_cholesky(A)
    elif nugget_type in ("fit", "fixed"):
        A += nugget * np.eye(A.shape[0])
        L = fixed_cholesky(A)
    else:
        raise ValueError("Bad value for nugget_type in cholesky_factor")
    return L, nugget


def cho_solve(L, b):
    """ChoInv.solve, linalg/cholesky.py:22-42."""
    if L.shape == (0, 0):
        return np.zeros(np.shape(b))
    return scipy.linalg.cho_solve((L, True), b)


def logdet(L):
    """ChoInv.logdet, linalg/cholesky.py:67-79."""
    return 2.0 * np.sum(np.log(np.diag(L)))


# ------------------------------------------------------------------------------------------------
# priors (scalar host math that enters current_logpost and its gradient)
# ------------------------------------------------------------------------------------------------

def _min_spacing(v):
    """Priors.py:1170-1188."""
    v = np.unique(np.array(v).flatten())
    if len(v) <= 2:
        return 0.0
    return np.median(np.diff(np.sort(v)))


def _max_spacing(v):
    """Priors.py:1151-1168."""
    v = np.unique(np.array(v).flatten())
    if len(v) <= 1:
        return 0.0
    s = np.sort(v)
    return s[-1] - s[0]


def invgamma_default(min_val, max_val):
    """InvGamma (shape, scale) with 99% of its mass in [min_val, max_val].

    PriorDist.default_prior, Priors.py:698-760 (root of the two CDF conditions in log space,
    started from zeros).  Returns None where the reference falls back to a weak prior.
    """
    def f(x):
        cdf = scipy.stats.invgamma(np.exp(x[0]), scale=np.exp(x[1])).cdf
        return np.array([cdf(min_val) - 0.005, cdf(max_val) - 0.995])

    res = root(f, np.zeros(2))
    if not res["success"]:
        return None
    return float(np.exp(res["x"][0])), float(np.exp(res["x"][1]))


def invgamma_default_mode(min_val, max_val):
    """InvGammaPrior.default_prior_mode, Priors.py:1013-1060."""
    mode = np.sqrt(min_val * max_val)

    def f(x):
        a = np.exp(x)
        return scipy.stats.invgamma(a, scale=(1.0 + a) * mode).cdf(max_val) - 0.995

    res = root(f, 0.0)
    if not res["success"]:
        return None
    a = float(np.exp(res["x"])[0]) if np.ndim(res["x"]) else float(np.exp(res["x"]))
    return a, (1.0 + a) * mode


def default_priors(inputs, nugget_type):
    """GPPriors.default_priors, Priors.py:86-152, for n_corr == D, dist="invgamma".

    Returns dict(corr=[(shape, scale) or None]*D, nugget=(shape, scale) or None); the covariance
    prior is weak (contributes 0).
    """
    corr = []
    for column in np.transpose(inputs):
        lo, hi = _min_spacing(column), _max_spacing(column)
        if lo == 0.0 or hi == 0.0:
            corr.append(None)
            continue
        p = invgamma_default(lo, hi)
        if p is None:
            p = invgamma_default_mode(lo, hi)
        corr.append(p)
    nug = invgamma_default_mode(1e-8, 1e-6) if nugget_type == "fit" else None
    return dict(corr=corr, nugget=nug)


def _invgamma_logp(x, shape, scale):
    """InvGammaPrior.logp, Priors.py:1105-1118."""
    return shape * np.log(scale) - gammaln(shape) - (shape + 1.0) * np.log(x) - scale / x


def _invgamma_dlogpdx(x, shape, scale):
    """InvGammaPrior.dlogpdx, Priors.py:1120-1131."""
    return -(shape + 1.0) / x + scale / x ** 2


def priors_logp(priors, theta_corr_raw, nugget):
    """GPPriors.logp, Priors.py:291-319 (correlation lengths l = exp(-theta/2), GPParams.py:3-80)."""
    total = 0.0
    for p, raw in zip(priors["corr"], theta_corr_raw):
        if p is not None:
            total += _invgamma_logp(np.exp(-0.5 * raw), *p)
    if priors.get("nugget") is not None:
        total += _invgamma_logp(nugget, *priors["nugget"])
    return total


def priors_dlogpdtheta(priors, theta_corr_raw, nugget, n_params):
    """GPPriors.dlogpdtheta, Priors.py:321-354; d(scaled)/d(raw): -l/2 (corr), s (cov/nugget)."""
    out = np.zeros(n_params)
    for i, (p, raw) in enumerate(zip(priors["corr"], theta_corr_raw)):
        if p is not None:
            ell = np.exp(-0.5 * raw)
            out[i] = _invgamma_dlogpdx(ell, *p) * (-0.5 * ell)
    if priors.get("nugget") is not None:
        out[-1] = _invgamma_dlogpdx(nugget, *priors["nugget"]) * nugget
    return out


# ------------------------------------------------------------------------------------------------
# single-output GP
# ------------------------------------------------------------------------------------------------

# Design matrices of the formula mean functions used by the fixtures and tests, written out column by column the way
# patsy.dmatrix(formula, data={"x": inputs.T}) builds them (GaussianProcess.py:503-512): intercept first unless removed,
# then the terms by interaction order; a left-hand side ("y ~") is ignored.  Hand-written on purpose: this table is the
# independent statement the product's formula evaluator (mogp_emulator_b200/formula.py) is checked against.
DESIGN_FUNCTIONS = {
    "x[0]": lambda X: [np.ones(len(X)), X[:, 0]],
    "y ~ x[0] + x[1]": lambda X: [np.ones(len(X)), X[:, 0], X[:, 1]],
    "x[0] + x[1]:x[2] + I(x[0]**2) + np.sin(x[1]) + x[2]": lambda X: [np.ones(len(X)), X[:, 0], X[:, 0] ** 2, np.sin(X[:, 1]),
                                                                       X[:, 2], X[:, 1] * X[:, 2]],
    "-1 + x[0]*x[1]": lambda X: [X[:, 0], X[:, 1], X[:, 0] * X[:, 1]],
}


def design_from_table(formula, inputs):
    inputs = np.atleast_2d(np.asarray(inputs, dtype=np.float64))
    return np.column_stack(DESIGN_FUNCTIONS[formula](inputs))


class OracleGP(object):
    """Zero-mean GaussianProcess restatement (GaussianProcess.py:86-927 with mean=None).

    theta layout is the reference's raw vector [theta_corr (D), theta_cov, (theta_nugget)]
    (GPParams.py:293-301): sigma^2 = exp(theta_cov), nugget = exp(theta_nugget) when fitted.
    ``priors``: None -> the reference's default priors; "weak" -> no prior terms; or the dict
    produced by default_priors().
    """

    def __init__(self, inputs, targets, kernel=SQEXP, nugget="adaptive", priors=None, chunked=False, mean=None):
        self.inputs = np.array(inputs, dtype=np.float64)
        if self.inputs.ndim == 1:
            self.inputs = self.inputs.reshape(-1, 1)
        self.mean = mean
        self.targets = np.array(targets, dtype=np.float64)
        assert self.targets.ndim == 1 and self.targets.shape[0] == self.inputs.shape[0]
        self.n, self.D = self.inputs.shape
        self.kernel = kernel
        if isinstance(nugget, str):
            assert nugget in ("adaptive", "fit")
            self.nugget_type, self.nugget = nugget, None
        else:
            assert float(nugget) >= 0.0
            self.nugget_type, self.nugget = "fixed", float(nugget)
        self.n_params = self.D + 1 + int(self.nugget_type == "fit")
        if priors is None:
            self.priors = default_priors(self.inputs, self.nugget_type)
        elif isinstance(priors, str) and priors == "weak":
            self.priors = dict(corr=[None] * self.D, nugget=None)
        else:
            self.priors = priors
        self.chunked = chunked
        self.theta = None
        self.L = None
        self.Kinv_t = None
        self.current_logpost = None

    # -- GaussianProcess.get_design_matrix, GaussianProcess.py:485-514: zero mean (no columns), the constant mean, a
    #    callable inputs -> design matrix, or one of the formulas of DESIGN_FUNCTIONS (what patsy's dmatrix returns for
    #    them, written out by hand -- patsy itself is not installed).  Mean priors are the reference's default (weak) ones.
    def get_design_matrix(self, inputs):
        inputs = np.atleast_2d(np.asarray(inputs, dtype=np.float64))
        if self.mean is None or (isinstance(self.mean, str) and self.mean in ("0", "-1")):
            return np.zeros((inputs.shape[0], 0))
        if callable(self.mean):
            return np.asarray(self.mean(inputs), dtype=np.float64)
        if self.mean in ("1", "-0"):
            return np.ones((inputs.shape[0], 1))
        if self.mean in DESIGN_FUNCTIONS:
            return design_from_table(self.mean, inputs)
        raise ValueError("mean function %r is not restated here (zero, constant, callable or DESIGN_FUNCTIONS)" % (self.mean,))

    @property
    def n_mean(self):
        return self.get_design_matrix(self.inputs[:1]).shape[1]

    # -- GaussianProcess.get_cov_matrix / get_K_matrix, GaussianProcess.py:517-558 ---------------
    def get_cov_matrix(self, other):
        other = np.atleast_2d(np.asarray(other, dtype=np.float64))
        cov = np.exp(self.theta[self.D])
        return cov * kernel_f(self.inputs, other, self.theta[:self.D], self.kernel, self.chunked)

    def get_K_matrix(self):
        return self.get_cov_matrix(self.inputs)

    # -- GaussianProcess.fit, GaussianProcess.py:629-685 (mean=None: m=0, Ainv is 0x0) -----------
    def fit(self, theta):
        theta = np.array(theta, dtype=np.float64)
        assert theta.shape == (self.n_params,), "bad shape for hyperparameters"
        self.theta = theta
        if self.nugget_type == "fit":
            self.nugget = float(np.exp(theta[-1]))
        elif self.nugget_type == "adaptive":
            self.nugget = None
        K = self.get_K_matrix()
        self.L, newnugget = cholesky_factor(K, self.nugget, self.nugget_type)
        if self.nugget_type == "adaptive":
            self.nugget = float(newnugget)
        self.Kinv_t = cho_solve(self.L, self.targets)
        dm = self.get_design_matrix(self.inputs)
        self.theta_mean = np.zeros(0)
        self.Kinv_t_mean = self.Kinv_t
        self.LA = None
        if dm.shape[1] == 0:
            self.current_logpost = 0.5 * (np.dot(self.targets, self.Kinv_t) + logdet(self.L)
                                          + self.n * np.log(2.0 * np.pi))
        else:
            # analytic mean with weak priors (GaussianProcess.py:657-685; linalg_utils.py:5-168): m = 0, B^-1 = 0
            self.Kinv_H = cho_solve(self.L, dm)                                   # (n, M)
            A = np.dot(dm.T, self.Kinv_H)                                         # calc_Ainv
            self.LA = scipy.linalg.cholesky(A, lower=True)
            H_Kinv_t = np.dot(dm.T, self.Kinv_t)
            self.theta_mean = cho_solve(self.LA, H_Kinv_t)                        # calc_mean_params
            self.Kinv_t_mean = cho_solve(self.L, self.targets - np.dot(dm, self.theta_mean))
            self.current_logpost = 0.5 * (np.dot(self.targets, self.Kinv_t)
                                          - np.dot(H_Kinv_t, cho_solve(self.LA, H_Kinv_t))
                                          + logdet(self.L) + logdet(self.LA)
                                          + (self.n - dm.shape[1]) * np.log(2.0 * np.pi))
        self.current_logpost -= priors_logp(self.priors, theta[:self.D], self.nugget)
        return self

    def _refit(self, theta):
        """GaussianProcess._refit, GaussianProcess.py:606-627."""
        return self.theta is None or not np.allclose(theta, self.theta, rtol=1e-10, atol=1e-15)

    def logposterior(self, theta):
        """GaussianProcess.logposterior, GaussianProcess.py:688-709."""
        if self._refit(theta):
            self.fit(theta)
        return self.current_logpost

    def logpost_deriv(self, theta):
        """GaussianProcess.logpost_deriv, GaussianProcess.py:711-782 with n_mean == 0.

        d/dtheta_i = 0.5*(tr(K^-1 dK_i) - alpha^T dK_i alpha) - dlogp/dtheta_i.  The reference
        evaluates tr(K^-1 dK_i) through logdet_deriv (linalg_utils.py:170-198) as
        trace(cho_solve(dK_i)); here the same trace is taken as sum(K^-1 * dK_i) with one explicit
        inverse, which is the identical quantity at O(n^3) instead of O(D n^3) cost.
        """
        if self._refit(theta):
            self.fit(theta)
        D, n = self.D, self.n
        partials = np.zeros(self.n_params)
        cov = np.exp(self.theta[D])
        exp_theta = np.exp(self.theta[:D])
        r2 = calc_r2_chunked(self.inputs, self.inputs, self.theta[:D])
        dKdr2 = cov * calc_dKdr2(r2, self.kernel)                       # Kernel.py:133-173
        Kinv = cho_solve(self.L, np.eye(n))
        if self.LA is None:
            G = Kinv - np.outer(self.Kinv_t, self.Kinv_t)
        else:
            # GaussianProcess.py:743-778 collected: -(t-q)^T dK (t-q) + tr(K^-1 dK) + tr(A^-1 dA), dA = -H^T K^-1 dK K^-1 H,
            # t - q = Kinv_t_mean for weak mean priors
            U = scipy.linalg.solve_triangular(self.LA, self.Kinv_H.T, lower=True).T      # W L_A^-T, (n, M)
            G = Kinv - np.dot(U, U.T) - np.outer(self.Kinv_t_mean, self.Kinv_t_mean)
        for i in range(D):
            diff2 = (self.inputs[:, i][:, None] - self.inputs[:, i][None, :]) ** 2
            partials[i] = 0.5 * np.sum(G * dKdr2 * (exp_theta[i] * diff2))   # Kernel.py:487-530
        Kmat = cov * calc_K(r2, self.kernel)
        partials[D] = 0.5 * np.sum(G * Kmat)                             # GaussianProcess.py:759-767
        if self.nugget_type == "fit":                                    # GaussianProcess.py:769-778
            partials[-1] = 0.5 * self.nugget * np.trace(G)
        partials -= priors_dlogpdtheta(self.priors, self.theta[:D], self.nugget, self.n_params)
        return partials

    # -- GaussianProcess.predict, GaussianProcess.py:818-927 (full_cov=False, zero mean) ---------
    def predict(self, testing, unc=True, include_nugget=True, full_cov=False):
        if self.theta is None:
            raise ValueError("hyperparameters have not been fit for this Gaussian Process")
        testing = np.array(testing, dtype=np.float64)
        if testing.ndim == 1:
            testing = testing.reshape(-1, 1) if self.D == 1 else testing.reshape(1, -1)
        assert testing.ndim == 2 and testing.shape[1] == self.D
        Ktest = self.get_cov_matrix(testing)                     # (n, m)
        dmtest = self.get_design_matrix(testing)
        mu = np.dot(dmtest, self.theta_mean) + np.dot(Ktest.T, self.Kinv_t_mean)
        var = None
        if unc:
            Kinv_Ktest = cho_solve(self.L, Ktest)
            extra = 0.0
            if self.LA is not None:                              # calc_R, linalg_utils.py:132-168; GaussianProcess.py:897-920
                R = dmtest.T - np.dot(self.get_design_matrix(self.inputs).T, Kinv_Ktest)
                LAinv_R = scipy.linalg.solve_triangular(self.LA, R, lower=True)
                extra = np.dot(LAinv_R.T, LAinv_R) if full_cov else np.sum(LAinv_R ** 2, axis=0)
            sigma_2 = np.exp(self.theta[self.D])
            if full_cov:                                     # GaussianProcess.py:899-911 (zero mean: no R term)
                sigma_2 = sigma_2 * kernel_f(testing, testing, self.theta[:self.D], self.kernel, self.chunked)
                if include_nugget:
                    sigma_2 = sigma_2 + np.eye(testing.shape[0]) * self.nugget
                Linv_Ktest = scipy.linalg.solve_triangular(self.L, Ktest, lower=True)
                return mu, sigma_2 - np.dot(Linv_Ktest.T, Linv_Ktest) + extra
            if include_nugget:
                sigma_2 = sigma_2 + self.nugget
            var = np.maximum(sigma_2 - np.sum(Ktest * Kinv_Ktest, axis=0) + extra, 0.0)
        return mu, var

    def predict_deriv(self, testing):
        """d mean / d x* at the test points, shape (m, D).  The CPU reference no longer provides this
        (GaussianProcess.py:922-926 warns and returns None); the definition is the reference GPU path's
        (mogp_gpu/src/densegp_gpu.hpp:411-448 with kernel.cu:69-100 / 264-302): sum_i alpha_i dk(x*, x_i)/dx*, where
        dk/dx*_q = sigma2 * dK/dr2 * 2 exp(theta_q) (x*_q - x_iq)."""
        testing = np.atleast_2d(np.asarray(testing, dtype=np.float64))
        w = np.exp(self.theta[:self.D])
        cov = np.exp(self.theta[self.D])
        out = np.empty((testing.shape[0], self.D))
        for c, x in enumerate(testing):                      # small cases only
            diff = x[np.newaxis, :] - self.inputs            # (n, D)
            r2 = np.sum(w * diff ** 2, axis=1)
            coef = cov * calc_dKdr2(r2, self.kernel) * self.Kinv_t_mean
            out[c] = 2.0 * w * np.dot(coef, diff)
        if self.LA is not None:
            # the mean function's share d(H* beta)/dx* (meanfunc.hpp mean_inputderiv in the GPU reference), by central
            # differences of the design matrix (exact for the constant mean)
            h = 1.0e-6
            for q in range(self.D):
                step = np.zeros(self.D)
                step[q] = h
                dH = (self.get_design_matrix(testing + step) - self.get_design_matrix(testing - step)) / (2.0 * h)
                out[:, q] += np.dot(dH, self.theta_mean)
        return out


# ------------------------------------------------------------------------------------------------
# multi-output fan-out
# ------------------------------------------------------------------------------------------------

class OracleMultiOutputGP(object):
    """MultiOutputGP restatement (MultiOutputGP.py:40-104, 182-319, 331-459): a list of independent
    single-output GPs over shared inputs; targets (E, n); predict returns (E, m) arrays; unfit
    emulators raise ValueError or give NaN rows with allow_not_fit."""

    def __init__(self, inputs, targets, kernel=SQEXP, nugget="adaptive", priors=None, chunked=False):
        targets = np.array(targets, dtype=np.float64)
        if targets.ndim == 1:
            targets = targets.reshape(1, -1)
        self.emulators = [OracleGP(inputs, t, kernel=kernel, nugget=nugget, priors=priors, chunked=chunked)
                          for t in targets]
        self.n_emulators = len(self.emulators)
        self.n, self.D = self.emulators[0].n, self.emulators[0].D

    def fit(self, thetas):
        thetas = np.array(thetas, dtype=np.float64)
        assert thetas.shape[0] == self.n_emulators
        for gp, theta in zip(self.emulators, thetas):     # serial loop, MultiOutputGP.py:348-349
            gp.fit(theta)

    def fit_emulator(self, index, theta):
        self.emulators[index].fit(theta)

    def get_indices_fit(self):
        return [i for i, gp in enumerate(self.emulators) if gp.theta is not None]

    def get_indices_not_fit(self):
        return [i for i, gp in enumerate(self.emulators) if gp.theta is None]

    def predict(self, testing, unc=True, include_nugget=True, allow_not_fit=False):
        testing = np.atleast_2d(np.asarray(testing, dtype=np.float64))
        if not allow_not_fit and len(self.get_indices_not_fit()) > 0:
            raise ValueError("hyperparameters have not been fit for this Gaussian Process")
        m = testing.shape[0]
        mean = np.full((self.n_emulators, m), np.nan)
        var = np.full((self.n_emulators, m), np.nan) if unc else None
        for i, gp in enumerate(self.emulators):
            if gp.theta is None:
                continue
            mu, v = gp.predict(testing, unc=unc, include_nugget=include_nugget)
            mean[i] = mu
            if unc:
                var[i] = v
        return mean, var


# ------------------------------------------------------------------------------------------------
# synthetic workloads shared by tests and bench
# ------------------------------------------------------------------------------------------------

# ------------------------------------------------------------------------------------------------
# validation diagnostics (mogp_emulator/validation.py) on a given prediction
# ------------------------------------------------------------------------------------------------

def pivot_cholesky(A):
    """linalg/cholesky.py:284-330: dpstrf; rows beyond the numerical rank get a decreasing fake diagonal."""
    A = np.ascontiguousarray(_check_cholesky_inputs(A))
    L, P, rank, info = lapack.dpstrf(A, lower=1)
    L = np.tril(L)
    if info < 0:
        raise scipy.linalg.LinAlgError("Illegal value in covariance matrix")
    n = A.shape[0]
    idx = np.arange(rank, n)
    L[idx, idx] = L[rank - 1, rank - 1] / np.cumprod(np.arange(rank + 1, n + 1, dtype=np.float64))
    return L, P - 1


def standard_errors(target, mean, var):
    """StandardErrors.__call__, validation.py:376-400."""
    P = np.argsort(var)[::-1]
    return ((mean - target) / np.sqrt(var))[P], P


def pivoted_errors(target, mean, cov):
    """PivotErrors.__call__, validation.py:416-441 (ChoInvPivot.solve_L, linalg/cholesky.py:130-166)."""
    L, P = pivot_cholesky(cov)
    if L.shape == (1, 1):
        return (mean - target) / L[0, 0], P
    return scipy.linalg.solve_triangular(L, (mean - target)[P], lower=True), P


def mahalanobis(target, mean, cov, n_train, n_mean=0, scaled=False):
    """validation.py:8-95 for one emulator: sum of squared pivoted errors, optionally standardised with the
    Fisher-Snedecor distribution F(n_valid, n - n_mean - 2) scaled by n_valid (validation.py:98-135)."""
    import scipy.stats
    M = np.sum(pivoted_errors(target, mean, cov)[0] ** 2)
    if scaled:
        n_valid = len(target)
        mu, var = scipy.stats.f(dfn=n_valid, dfd=n_train - n_mean - 2, scale=n_valid).stats()
        M = (M - mu) / np.sqrt(var)
    return M


def make_workload(n, d, n_out, m, seed):
    """X~U[0,1)^d, Y[k] = sin(2*sum(x)+k) + 0.01*N(0,1), Xs~U[0,1)^d, all from default_rng(seed)."""
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    Y = np.stack([np.sin(2.0 * X.sum(axis=1) + k) + 0.01 * rng.standard_normal(n) for k in range(n_out)])
    Xs = rng.random((m, d))
    return X, Y, Xs


# ------------------------------------------------------------------------------------------------
# history matching (mogp_emulator/HistoryMatching.py) on a given prediction
# ------------------------------------------------------------------------------------------------

def implausibility(mean, var, obs_val, obs_var, discrepancy=0.0, rank=1):
    """HistoryMatching.get_implausibility, HistoryMatching.py:255-289: mean / var (n_obs, m) (or (m,) for one output);
    the (rank+1)-th largest of |z - mean| / sqrt(var + discrepancy + obs_var) over the outputs, rank forced to 0 for one."""
    mean, var = np.atleast_2d(mean), np.atleast_2d(var)
    obs_val, obs_var = np.atleast_1d(obs_val), np.atleast_1d(obs_var)
    n_obs = len(obs_val)
    if n_obs == 1:
        rank = 0
    Vs = var + np.atleast_1d(discrepancy)[:, np.newaxis] + obs_var[:, np.newaxis]
    I = np.abs(obs_val[:, np.newaxis] - mean) / np.sqrt(Vs)
    return np.partition(I, n_obs - rank - 1, axis=0)[n_obs - rank - 1]


# ------------------------------------------------------------------------------------------------
# MICEFastGP.fast_predict (mogp_emulator/SequentialDesign.py:705-748)
# ------------------------------------------------------------------------------------------------

def mice_fast_predict(gp, index):
    """Variance at training point ``index`` of the fitted OracleGP ``gp`` after removing that point, by the reference's
    route: explicit (L L^T)^-1, Woodbury down-date, sigma2 + nugget - k^T Q k, clipped at zero."""
    n = gp.n
    keep = np.arange(n) != index
    cov = np.exp(gp.theta[gp.D])
    Ktest = cov * kernel_f(gp.inputs[keep], gp.inputs[index:index + 1], gp.theta[:gp.D], gp.kernel)
    invQ = np.linalg.solve(gp.L.T, np.linalg.solve(gp.L, np.eye(n)))
    invQ_mod = invQ[keep][:, keep] - np.outer(invQ[keep, index], invQ[keep, index]) / invQ[index, index]
    return np.maximum(cov + gp.nugget - np.sum(Ktest * np.dot(invQ_mod, Ktest), axis=0), 0.0)
