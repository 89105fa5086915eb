"""Host side of the analytic mean function (CPU semantics: GaussianProcess.get_design_matrix, GaussianProcess.py:485-514;
fit / predict algebra GaussianProcess.py:657-685, 887-920; linalg_utils.py:5-168 calc_Ainv / calc_mean_params / calc_R).

The mean parameters are integrated out analytically with the reference's default (weak) mean priors.  The design
matrix comes from ``formula.MeanFormula`` (a patsy-free evaluator of the formula subset that makes sense for numeric
inputs: ``"x[0]"``, ``"x[0] + x[1]:x[2]"``, ``"I(x[0]**2)"``, ``"np.sin(x[1])"``, ``"y ~ x[0]"``, ``"-1 + x[0]"`` ...); the zero
mean is ``None``, ``"0"`` or ``"-1"``, the constant mean ``"1"`` or ``"-0"``.  The algebra below is written for a general
design matrix with up to ``MAX_MEAN`` columns (the device keeps that many rank-1 vectors per output for the gradient).
The n x n work -- the solves K^-1 H, the products H^T K^-1 K* -- runs on the GPU (mogp_solve_list, mogp_kstar_dot); only
n_mean x n_mean matrices are handled here.
"""
import numpy as np
import scipy.linalg

from .formula import MeanFormula

MAX_MEAN = 32     # mogp_set_mean_vectors_list: vectors per output (csrc/api.cu MAXM)


def interpret_mean(mean):
    """-> ``None`` (zero mean) or a ``MeanFormula``; ValueError("Provided mean function is invalid") otherwise
    (GaussianProcess.py:499-512)."""
    if mean is None:
        return None
    if isinstance(mean, MeanFormula):
        spec = mean
    elif isinstance(mean, str):
        spec = MeanFormula(mean)
    else:
        raise ValueError("Provided mean function is invalid")
    if spec.n_mean == 0:
        return None
    if spec.n_mean > MAX_MEAN:
        raise ValueError("Provided mean function is invalid: at most %d design-matrix columns are supported" % MAX_MEAN)
    return spec


def design_matrix(spec, inputs):
    inputs = np.atleast_2d(np.asarray(inputs, dtype=np.float64))
    if spec is None:
        return np.zeros((inputs.shape[0], 0))
    return spec.design_matrix(inputs)


def design_matrix_inputderiv(spec, inputs):
    """dH/dx (m, D, n_mean) -- the mean function's share of the predictive derivatives."""
    inputs = np.atleast_2d(np.asarray(inputs, dtype=np.float64))
    if spec is None:
        return np.zeros(inputs.shape + (0,))
    return spec.input_deriv(inputs)


class MeanFit(object):
    """Everything the analytic mean adds to one fitted emulator."""

    __slots__ = ("beta", "alpha_mean", "U", "W", "LA", "logpost_terms")

    def __init__(self, H, y, t, W, n):
        """H (n, M) design matrix, y targets, t = K^-1 y, W = K^-1 H (n, M)."""
        M = H.shape[1]
        A = np.dot(H.T, W)                                             # calc_Ainv with B^-1 = 0
        A = 0.5 * (A + A.T)
        self.LA = scipy.linalg.cholesky(A, lower=True)
        v = np.dot(H.T, t)                                             # H^T K^-1 y
        self.beta = scipy.linalg.cho_solve((self.LA, True), v)         # calc_mean_params
        self.alpha_mean = t - np.dot(W, self.beta)                     # Kinv_t_mean = K^-1 (y - H beta)
        self.W = W
        # K^-1 H A^-1 H^T K^-1 = U U^T with U = W L_A^-T
        self.U = scipy.linalg.solve_triangular(self.LA, W.T, lower=True).T
        # data part of current_logpost: 0.5*(y^T t - v^T A^-1 v + logdet K + logdet A + (n - M) log 2 pi); the caller adds
        # y^T t and logdet K (from the device) -- here: the mean-specific corrections
        self.logpost_terms = (-np.dot(v, self.beta), 2.0 * np.sum(np.log(np.diag(self.LA))), -M)

    def data_logpost(self, quad, logdet, n):
        dq, dld, dn = self.logpost_terms
        return 0.5 * (quad + dq + logdet + dld + (n + dn) * np.log(2.0 * np.pi))

    def variance_term(self, Hs, HtKinvKs, full_cov=False):
        """R^T A^-1 R with R = H*^T - H^T K^-1 K*  (calc_R); Hs (m, M), HtKinvKs (M, m)."""
        R = Hs.T - HtKinvKs
        LAinv_R = scipy.linalg.solve_triangular(self.LA, R, lower=True)
        return np.dot(LAinv_R.T, LAinv_R) if full_cov else np.sum(LAinv_R ** 2, axis=0)
