"""Validation diagnostics over the GPU emulators -- the consumers of ``predict(full_cov=True)`` in the reference
(mogp_emulator/validation.py:8-482).

``standard_errors`` needs the predictive variances, ``pivoted_errors`` / ``mahalanobis`` the full (m, m) predictive
covariance: both come from the device (``mogp_predict`` / ``mogp_predict_cov``: V = L^-1 K* by the dataflow TRSM, K** by
the symmetric kernel-matrix kernel, one DMMA SYRK).  What is left for the host is the m x m pivoted Cholesky of that
covariance (LAPACK ``dpstrf`` through SciPy, exactly the reference's ``pivot_cholesky``, linalg/cholesky.py:284-330) and
a triangular solve with one right-hand side.

Differences from the reference that are fixes, not drift: ``compute_errors`` accepts the method names it documents
(the reference compares the *unbound* ``method.lower`` with its list of names, validation.py:215-220, so every string
raises).
"""
import numpy as np
import scipy.linalg
from scipy.linalg import lapack
from scipy.stats import f as _fisher_f

from .GaussianProcessGPU import GaussianProcessGPU
from .MultiOutputGP_GPU import MultiOutputGP_GPU


def pivot_cholesky(A):
    """Pivoted Cholesky of a symmetric, possibly singular covariance (linalg/cholesky.py:284-330): LAPACK dpstrf;
    collinear rows get a decreasing fake diagonal so that solves and log-determinants stay finite.
    Returns ``(L, P)`` with ``A[P][:, P] = L L^T`` (on the rank-revealed part)."""
    A = np.array(A, dtype=np.float64)
    assert A.ndim == 2 and A.shape[0] == A.shape[1], "A must have shape (n,n)"
    np.testing.assert_allclose(A.T, A)
    if np.any(np.diag(A) <= 0.0):
        raise scipy.linalg.LinAlgError("not pd: non-positive diagonal elements")
    L, P, rank, info = lapack.dpstrf(np.ascontiguousarray(A), lower=1)
    L = np.tril(L)
    if info < 0:
        raise scipy.linalg.LinAlgError("Illegal value in covariance matrix")
    n = A.shape[0]
    idx = np.arange(rank, n)
    divs = np.cumprod(np.arange(rank + 1, n + 1, dtype=np.float64))
    L[idx, idx] = L[rank - 1, rank - 1] / divs
    return L, P - 1


class Errors(object):
    """Base class of the error measures (validation.py:352-361): ``full_cov`` says which predictive uncertainty the
    measure needs, ``__call__(target, mean, cov)`` returns ``(errors, ordering)``."""
    full_cov = False

    def __call__(self, target, mean, cov):
        raise NotImplementedError("Errors base class does not implement a call method")


class StandardErrors(Errors):
    """(mean - target) / sqrt(var), ordered by decreasing predictive variance (validation.py:364-400)."""
    full_cov = False

    def __call__(self, target, mean, cov):
        P = np.argsort(cov)[::-1]
        return ((mean - target) / np.sqrt(cov))[P], P


class PivotErrors(Errors):
    """Errors decorrelated with the pivoted Cholesky factor of the predictive covariance: ordered by decreasing variance
    conditional on all previous errors (validation.py:403-441)."""
    full_cov = True

    def __call__(self, target, mean, cov):
        L, P = pivot_cholesky(cov)
        if L.shape == (1, 1):
            return (mean - target) / L[0, 0], P
        return scipy.linalg.solve_triangular(L, (mean - target)[P], lower=True), P


def _is_single(gp):
    return isinstance(gp, GaussianProcessGPU)


def _process_inputs(gp, valid_inputs):
    valid_inputs = np.array(valid_inputs, dtype=np.float64)
    if valid_inputs.ndim == 1:
        valid_inputs = valid_inputs.reshape(-1, 1) if gp.D == 1 else valid_inputs.reshape(1, -1)
    assert valid_inputs.ndim == 2 and valid_inputs.shape[1] == gp.D, "bad shape for validation inputs"
    return valid_inputs


def _check_valid_data(gp, valid_inputs, valid_targets):
    """validation.py:463-482."""
    assert isinstance(gp, (GaussianProcessGPU, MultiOutputGP_GPU)), "Must provide a GP to validate"
    valid_inputs = _process_inputs(gp, valid_inputs)
    valid_targets = np.array(valid_targets, dtype=np.float64)
    if _is_single(gp):
        assert valid_targets.ndim == 1, "Targets for a GP must be a 1D array"
        assert valid_targets.shape[0] == valid_inputs.shape[0], "Bad length for validation targets"
    else:
        assert valid_targets.ndim == 2, "Targets for a MultiOutputGP must be a 2D array"
        assert valid_targets.shape[1] == valid_inputs.shape[0], "Bad shape for validation targets"
        assert valid_targets.shape[0] == gp.n_emulators, "Bad shape for validation targets"
    return valid_inputs, valid_targets


def compute_errors(gp, valid_inputs, valid_targets, method):
    """General pattern (validation.py:138-237): predict at the validation inputs with the kind of uncertainty the method
    needs, then apply the method per emulator.  One ``(errors, ordering)`` tuple for a ``GaussianProcessGPU``, a list of
    them for a ``MultiOutputGP_GPU``.  An unfit emulator raises ``ValueError`` (from ``predict``)."""
    if isinstance(method, str):
        name = method.lower()
        if name in ("standard", "standarderrors"):
            methodobj = StandardErrors()
        elif name in ("pivot", "pivoterrors"):
            methodobj = PivotErrors()
        else:
            raise ValueError("Bad value for error method in compute_errors")
    else:
        methodobj = method
    assert issubclass(type(methodobj), Errors), "method must be a subclass of Errors"
    valid_inputs, valid_targets = _check_valid_data(gp, valid_inputs, valid_targets)
    res = gp.predict(valid_inputs, unc=True, deriv=False, full_cov=methodobj.full_cov)
    if _is_single(gp):
        return methodobj(valid_targets, res.mean, res.unc)
    return [methodobj(t, mu, c) for t, mu, c in zip(valid_targets, res.mean, res.unc)]


def standard_errors(gp, valid_inputs, valid_targets):
    """validation.py:240-293."""
    return compute_errors(gp, valid_inputs, valid_targets, StandardErrors())


def pivoted_errors(gp, valid_inputs, valid_targets):
    """validation.py:296-349."""
    return compute_errors(gp, valid_inputs, valid_targets, PivotErrors())


def generate_mahal_dist(gp, valid_inputs):
    """Expected distribution of the Mahalanobis distance (validation.py:98-135): a frozen ``scipy.stats.f`` with
    ``(n_valid, n - n_mean - 2)`` degrees of freedom scaled by ``n_valid``; a list of them for a multi-output emulator."""
    if not isinstance(gp, (GaussianProcessGPU, MultiOutputGP_GPU)):
        raise TypeError("Provided GP is not a GaussianProcessGPU or MultiOutputGP_GPU")
    n_valid = len(_process_inputs(gp, valid_inputs))
    n_mean = gp.n_mean if _is_single(gp) else gp._dm.shape[1]
    n_em = 1 if _is_single(gp) else gp.n_emulators
    dists = [_fisher_f(dfn=n_valid, dfd=gp.n - n_mean - 2, scale=n_valid) for _ in range(n_em)]
    return dists[0] if len(dists) == 1 else dists


def mahalanobis(gp, valid_inputs, valid_targets, scaled=False):
    """Mahalanobis distance of the validation errors, (y - mu)^T Sigma^-1 (y - mu) with Sigma the predictive covariance
    (validation.py:8-95); ``scaled=True`` subtracts the mean and divides by the standard deviation of the expected
    Fisher-Snedecor distribution.  Scalar for a single emulator, ``(n_emulators,)`` array otherwise."""
    piv = pivoted_errors(gp, valid_inputs, valid_targets)
    errors = piv[0] if _is_single(gp) else np.array([e[0] for e in piv])
    M = np.sum(errors ** 2, axis=-1)
    if scaled:
        dists = generate_mahal_dist(gp, valid_inputs)
        single = _is_single(gp)
        M_iter = [M] if single else M
        dist_iter = dists if isinstance(dists, list) else [dists]
        out = []
        for M_val, dist in zip(M_iter, dist_iter):
            mean, var = dist.stats()
            out.append((M_val - mean) / np.sqrt(var))
        M = np.array(out)
        if single:
            M = M.squeeze(axis=0)
    return M
