"""Validation diagnostics (mogp_emulator_b200/validation.py; reference mogp_emulator/validation.py): known answers of the
reference's own tests, reference-generated goldens (tests/golden/validation/*.npz), the oracle restatement, the front
end over the numpy test double on CPU, and -- marked gpu -- the same goldens through the CUDA path."""
import glob
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_equal

import gp_oracle as orc
from golden import known_answers as ka
from fake_device import FakeHandle

VALID = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "validation", "*.npz")))


def test_known_answers_of_the_reference_tests():
    """test_validation.py:58-160 of the reference, on the oracle restatement and on the product's error classes."""
    from mogp_emulator_b200.validation import StandardErrors, PivotErrors, pivot_cholesky
    idx = ka.VALID_ORDER
    A = np.linalg.cholesky(ka.VALID_COV[idx][:, idx])
    for target, mean in ((ka.VALID_TARGETS1, ka.VALID_MEAN1), (ka.VALID_TARGETS2, ka.VALID_MEAN2)):
        err = mean - target
        want_std = err[idx] / np.sqrt(np.diag(ka.VALID_COV)[idx])
        want_piv = np.linalg.solve(A, err[idx])
        for std_fn, piv_fn in ((orc.standard_errors, orc.pivoted_errors), (StandardErrors(), PivotErrors())):
            e, P = std_fn(target, mean, np.diag(ka.VALID_COV))
            assert_allclose(e, want_std)
            assert_equal(P, idx)
            e, P = piv_fn(target, mean, ka.VALID_COV)
            assert_allclose(e, want_piv)
            assert_equal(P, idx)
        assert_allclose(orc.mahalanobis(target, mean, ka.VALID_COV, 10), np.dot(err, np.linalg.solve(ka.VALID_COV, err)))
    L, P = pivot_cholesky(ka.VALID_COV)
    assert_allclose(np.dot(L, L.T), ka.VALID_COV[P][:, P])
    # a singular covariance (two identical validation points): the factor stays usable (cholesky.py:316-330)
    C = np.array([[1.0, 1.0, 0.3], [1.0, 1.0, 0.3], [0.3, 0.3, 2.0]])
    L, P = pivot_cholesky(C)
    Lo, Po = orc.pivot_cholesky(C)
    assert_allclose(L, Lo)
    assert_equal(P, Po)
    assert np.all(np.diag(L) > 0.0) and L[2, 2] == L[1, 1] / 3.0


def _golden_gp(cls_single, cls_multi, g):
    nug = float(g["nugget_in"]) if str(g["nugget_type"]) == "fixed" else str(g["nugget_type"])
    mean = str(g["mean_spec"]) if "mean_spec" in g.files else None
    n_out = int(g["n_out"])
    if n_out == 1:
        gp = cls_single(g["X"], g["y"], mean=mean, kernel=str(g["kernel"]), nugget=nug)
        gp.fit(g["theta"])
    else:
        gp = cls_multi(g["X"], g["y"], mean=mean, kernel=str(g["kernel"]), nugget=nug)
        gp.fit(np.tile(g["theta"], (n_out, 1)) + 0.1 * np.arange(n_out)[:, None])
    return gp, n_out


def _check_against_golden(val, gp, n_out, g, rtol):
    se = val.standard_errors(gp, g["Xv"], g["yv"])
    pe = val.pivoted_errors(gp, g["Xv"], g["yv"])
    if n_out == 1:
        se, pe = [se], [pe]
    want = {k: np.atleast_2d(g[k]) for k in ("std_err", "std_idx", "piv_err", "piv_idx")}
    for k in range(n_out):
        assert_equal(se[k][1], want["std_idx"][k])
        assert_allclose(se[k][0], want["std_err"][k], rtol=rtol, atol=rtol)
        assert_equal(pe[k][1], want["piv_idx"][k])
        assert_allclose(pe[k][0], want["piv_err"][k], rtol=rtol, atol=rtol * np.abs(want["piv_err"][k]).max())
    assert_allclose(val.mahalanobis(gp, g["Xv"], g["yv"]), g["mahal"], rtol=rtol)
    assert_allclose(val.mahalanobis(gp, g["Xv"], g["yv"], scaled=True), g["mahal_scaled"], rtol=rtol)
    assert np.shape(val.mahalanobis(gp, g["Xv"], g["yv"])) == (() if n_out == 1 else (n_out,))
    again = val.compute_errors(gp, g["Xv"], g["yv"], "pivot")
    assert_allclose((again if n_out == 1 else again[0])[0], pe[0][0], rtol=1e-12)
    again = val.compute_errors(gp, g["Xv"], g["yv"], "StandardErrors")
    assert_allclose((again if n_out == 1 else again[0])[0], se[0][0], rtol=1e-12)
    with pytest.raises(ValueError):
        val.compute_errors(gp, g["Xv"], g["yv"], "nonsense")
    with pytest.raises(AssertionError):
        val.standard_errors(gp, g["Xv"], g["yv"][..., :-1])
    dist = val.generate_mahal_dist(gp, g["Xv"])
    first = dist if n_out == 1 else dist[0]
    n_mean = 2 if "mean_spec" in g.files else 0
    assert first.kwds["dfn"] == len(g["Xv"]) and first.kwds["dfd"] == gp.n - n_mean - 2


@pytest.mark.parametrize("path", VALID)
def test_oracle_matches_reference_golden(path):
    g = np.load(path)
    nug = float(g["nugget_in"]) if str(g["nugget_type"]) == "fixed" else str(g["nugget_type"])
    mean = str(g["mean_spec"]) if "mean_spec" in g.files else None
    n_out = int(g["n_out"])
    Y, Yv = np.atleast_2d(g["y"]), np.atleast_2d(g["yv"])
    for k in range(n_out):
        gp = orc.OracleGP(g["X"], Y[k], kernel=str(g["kernel"]), nugget=nug, mean=mean).fit(g["theta"] + 0.1 * k)
        mu, var = gp.predict(g["Xv"])
        _, cov = gp.predict(g["Xv"], full_cov=True)
        e, P = orc.standard_errors(Yv[k], mu, var)
        assert_equal(P, np.atleast_2d(g["std_idx"])[k])
        assert_allclose(e, np.atleast_2d(g["std_err"])[k], rtol=1e-6, atol=1e-6)
        e, P = orc.pivoted_errors(Yv[k], mu, cov)
        assert_equal(P, np.atleast_2d(g["piv_idx"])[k])
        assert_allclose(e, np.atleast_2d(g["piv_err"])[k], rtol=1e-6, atol=1e-6)
        n_mean = gp.n_mean
        assert_allclose(orc.mahalanobis(Yv[k], mu, cov, gp.n, n_mean), np.atleast_1d(g["mahal"])[k], rtol=1e-6)
        assert_allclose(orc.mahalanobis(Yv[k], mu, cov, gp.n, n_mean, scaled=True), np.atleast_1d(g["mahal_scaled"])[k],
                        rtol=1e-6)


@pytest.mark.parametrize("path", VALID)
def test_front_end_over_the_test_double(path, monkeypatch):
    """Host logic of validation.py on CPU: the classes run over tests/fake_device.py (numpy stand-in for the device)."""
    from mogp_emulator_b200 import libmogp
    monkeypatch.setattr(libmogp, "Handle", FakeHandle)
    monkeypatch.setattr(libmogp, "HAVE_LIBMOGP", True)
    monkeypatch.setattr(libmogp, "gpu_usable", lambda: True)
    import mogp_emulator_b200 as mogp
    from mogp_emulator_b200 import validation
    g = np.load(path)
    gp, n_out = _golden_gp(mogp.GaussianProcessGPU, mogp.MultiOutputGP_GPU, g)
    _check_against_golden(validation, gp, n_out, g, rtol=1e-6)
    with pytest.raises(AssertionError):
        validation.standard_errors(object(), g["Xv"], g["yv"])
    with pytest.raises(TypeError):
        validation.generate_mahal_dist(object(), g["Xv"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", VALID)
def test_gpu_validation_matches_reference_golden(path):
    """The same goldens through the CUDA path: variances from mogp_predict, (m, m) covariances from mogp_predict_cov."""
    import mogp_emulator_b200 as mogp
    from mogp_emulator_b200 import validation
    g = np.load(path)
    gp, n_out = _golden_gp(mogp.GaussianProcessGPU, mogp.MultiOutputGP_GPU, g)
    _check_against_golden(validation, gp, n_out, g, rtol=1e-5)
    gp.close()


@pytest.mark.gpu
def test_gpu_unfit_emulator_raises():
    import mogp_emulator_b200 as mogp
    from mogp_emulator_b200 import validation
    g = np.load(VALID[0])
    gp = mogp.GaussianProcessGPU(g["X"], np.atleast_2d(g["y"])[0])
    with pytest.raises(ValueError):
        validation.standard_errors(gp, g["Xv"], np.atleast_2d(g["yv"])[0])
    gp.close()
