"""Pin the CPU oracle (oracle/gp_oracle.py): known answers from the reference's own tests, golden
outputs of the unmodified reference (tests/golden/*.npz), and -- when /root/reference is present
(build container only) -- a live comparison against the reference itself."""
import glob
import os

import numpy as np
import pytest
import scipy.linalg
from numpy.testing import assert_allclose

import gp_oracle as orc
from golden import known_answers as ka

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def _nugget_arg(g):
    t = str(g["nugget_type"])
    return float(g["nugget_in"]) if t == "fixed" else t


def test_r2_known_answers():
    for x1, x2, theta, want in ka.R2_CASES:
        assert_allclose(orc.calc_r2(x1, x2, theta), want)
        assert_allclose(orc.calc_r2_chunked(x1, x2, theta, rows=1), want)


def test_kernel_closed_forms():
    rng = np.random.default_rng(0)
    x1, x2 = rng.random((7, 3)), rng.random((5, 3))
    theta = np.array([0.3, -1.0, 2.0])
    r2 = orc.calc_r2(x1, x2, theta)
    assert_allclose(orc.kernel_f(x1, x2, theta, orc.SQEXP), np.exp(-0.5 * r2))
    assert_allclose(orc.kernel_f(x1, x2, theta, orc.MAT52), ka.matern52_closed_form(r2))
    assert_allclose(orc.kernel_f(x1, x1, theta, orc.MAT52).diagonal(), 1.0)


def test_r2_inf_raises():
    with pytest.raises(FloatingPointError):
        orc.calc_r2(np.array([[1.0]]), np.array([[2.0]]), np.array([800.0]))


def test_cholesky_known_answers():
    L, nug = orc.cholesky_factor(ka.CHOL_WIKI_A.copy(), 0.0, "fixed")
    assert_allclose(L, ka.CHOL_WIKI_L)
    assert nug == 0.0
    L, nug = orc.cholesky_factor(ka.CHOL_NEAR_SINGULAR_A.copy(), ka.CHOL_NEAR_SINGULAR_NUGGET, "fixed")
    assert_allclose(L, ka.CHOL_NEAR_SINGULAR_L)
    L, jitter = orc.cholesky_factor(ka.CHOL_WIKI_A.copy(), None, "adaptive")
    assert_allclose(L, ka.CHOL_WIKI_L)
    assert jitter == 0.0
    L, jitter = orc.cholesky_factor(ka.CHOL_NEAR_SINGULAR_A.copy(), None, "adaptive")
    assert_allclose(L, ka.CHOL_NEAR_SINGULAR_L)
    assert_allclose(jitter, 1.0e-6)
    with pytest.raises(scipy.linalg.LinAlgError):
        orc.jit_cholesky(ka.CHOL_NOT_PD_A.copy())


def test_variance_stability_case():
    c = ka.VAR_STABILITY
    gp = orc.OracleGP(c["x"], c["y"], nugget=c["nugget"]).fit(c["theta"])
    _, var = gp.predict(c["testing"])
    assert_allclose(np.zeros(101), var, atol=c["atol"])


SINGLE = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                if not os.path.basename(p).startswith("multi_"))
MULTI = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "multi_*.npz")))


@pytest.mark.parametrize("name", SINGLE)
def test_oracle_matches_reference_golden_single(name):
    g = _load(name)
    mean_spec = str(g["mean_spec"]) if "mean_spec" in g else None
    gp = orc.OracleGP(g["X"], g["y"], kernel=str(g["kernel"]), nugget=_nugget_arg(g), mean=mean_spec)
    gp.fit(g["theta"])
    if mean_spec is not None:       # analytic mean parameter of the constant mean function
        assert_allclose(gp.theta_mean, g["theta_mean"], rtol=1e-8)
        assert_allclose(gp.Kinv_t_mean, g["Kinv_t_mean"], rtol=1e-7, atol=1e-9 * np.abs(g["Kinv_t_mean"]).max())
    assert_allclose(gp.get_K_matrix(), g["K"], rtol=1e-13, atol=1e-300)
    assert_allclose(gp.L, g["L"], rtol=1e-9, atol=1e-12)
    assert_allclose(gp.Kinv_t, g["Kinv_t"], rtol=1e-7, atol=1e-9)
    assert_allclose(gp.current_logpost, np.squeeze(g["logpost"]), rtol=1e-10)
    if str(g["nugget_type"]) == "adaptive":
        assert gp.nugget == float(g["nugget_out"])          # exact: same jitter schedule
    else:
        assert_allclose(gp.nugget, float(g["nugget_out"]), rtol=1e-14)
    mean, var = gp.predict(g["Xs"])
    assert_allclose(mean, g["mean"], rtol=1e-6, atol=1e-9)
    nug = max(gp.nugget, 1e-12)
    assert_allclose(var, g["var"], rtol=1e-4, atol=1e-4 * nug + 1e-10)
    _, var_nn = gp.predict(g["Xs"], include_nugget=False)
    assert_allclose(var_nn, g["var_no_nugget"], rtol=1e-4, atol=1e-4 * nug + 1e-10)
    mean_fc, cov = gp.predict(g["Xs"], full_cov=True)
    assert_allclose(cov, g["cov_full"], rtol=1e-4, atol=1e-4 * nug + 1e-10)
    assert_allclose(mean_fc, g["mean"], rtol=1e-6, atol=1e-9)
    if "deriv" in g:
        assert_allclose(gp.logpost_deriv(g["theta"]), g["deriv"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", SINGLE)
def test_default_priors_match_reference(name):
    g = _load(name)
    pri = orc.default_priors(g["X"], str(g["nugget_type"]))
    for got, want in zip(pri["corr"], g["prior_corr"]):
        if np.isnan(want[0]):
            assert got is None
        else:
            assert_allclose(got, want, rtol=1e-8)
    pn = np.ravel(g["prior_nugget"])
    if np.isnan(pn[0]):
        assert pri["nugget"] is None
    else:
        assert_allclose(pri["nugget"], pn, rtol=1e-8)


@pytest.mark.parametrize("name", MULTI)
def test_oracle_matches_reference_golden_multi(name):
    g = _load(name)
    t = str(g["nugget_type"])
    nugget = float(g["nugget_in"]) if t == "fixed" else t
    mo = orc.OracleMultiOutputGP(g["X"], g["Y"], kernel=str(g["kernel"]), nugget=nugget)
    with pytest.raises(ValueError):
        mo.predict(g["Xs"])
    mo.fit(g["thetas"])
    mean, var = mo.predict(g["Xs"])
    assert mean.shape == g["mean"].shape and var.shape == g["var"].shape
    assert_allclose([e.current_logpost for e in mo.emulators], g["logposts"], rtol=1e-9)
    assert_allclose([e.nugget for e in mo.emulators], g["nuggets"], rtol=1e-14)
    assert_allclose(mean, g["mean"], rtol=1e-6, atol=1e-8)
    nug = max(float(np.max(g["nuggets"])), 1e-12)
    assert_allclose(var, g["var"], rtol=1e-4, atol=1e-4 * nug + 1e-9)


def test_multi_allow_not_fit_rows_are_nan():
    X, Y, Xs = orc.make_workload(40, 2, 3, 5, 1)
    mo = orc.OracleMultiOutputGP(X, Y, nugget=1e-6)
    mo.fit_emulator(1, np.zeros(3))
    assert mo.get_indices_fit() == [1] and mo.get_indices_not_fit() == [0, 2]
    mean, var = mo.predict(Xs, allow_not_fit=True)
    assert np.all(np.isnan(mean[[0, 2]])) and np.all(np.isfinite(mean[1]))
    assert np.all(np.isnan(var[[0, 2]])) and np.all(np.isfinite(var[1]))


def test_logpost_deriv_vs_finite_differences():
    X, Y, _ = orc.make_workload(30, 2, 1, 1, 3)
    for kernel in (orc.SQEXP, orc.MAT52):
        for nugget in (1e-4, "fit"):
            gp = orc.OracleGP(X, Y[0], kernel=kernel, nugget=nugget)
            theta = np.array([0.4, 0.9, 0.2] + ([-6.0] if nugget == "fit" else []))
            g = gp.logpost_deriv(theta)
            fd = np.zeros_like(theta)
            for i in range(len(theta)):
                e = np.zeros_like(theta)
                e[i] = 1e-6
                fd[i] = (gp.logposterior(theta + e) - gp.logposterior(theta - e)) / 2e-6
            assert_allclose(g, fd, rtol=1e-4, atol=1e-4)


def test_live_reference_if_present():
    """Build-container only: compare against the imported reference on a fresh random problem."""
    import refstub
    if not refstub.reference_available():
        pytest.skip("reference tree not present (expected on the GPU box)")
    mogp = refstub.import_reference()
    X, Y, Xs = orc.make_workload(180, 4, 1, 50, 99)
    theta = np.array([0.7, 1.1, 0.9, 1.0, 0.1])
    ref = mogp.GaussianProcess(X, Y[0], nugget=1e-6)
    ref.fit(theta)
    gp = orc.OracleGP(X, Y[0], nugget=1e-6).fit(theta)
    assert_allclose(gp.current_logpost, ref.current_logpost, rtol=1e-12)
    assert_allclose(gp.L, ref.Kinv.L, rtol=1e-10, atol=1e-13)
    rm, rv, _ = ref.predict(Xs)
    m, v = gp.predict(Xs)
    assert_allclose(m, rm, rtol=1e-9, atol=1e-11)
    assert_allclose(v, rv, rtol=1e-6, atol=1e-10)
    assert_allclose(gp.logpost_deriv(theta), ref.logpost_deriv(theta), rtol=1e-7, atol=1e-7)


def test_predict_deriv_vs_finite_differences():
    """The oracle's mean-derivative restatement (reference GPU definition) against central differences of the
    oracle's own posterior mean."""
    for kernel in (orc.SQEXP, orc.MAT52):
        X, Y, Xs = orc.make_workload(60, 3, 1, 9, seed=3)
        theta = np.array([0.7, 1.1, 0.9, 0.3])
        gp = orc.OracleGP(X, Y[0], kernel=kernel, nugget=1e-6).fit(theta)
        got = gp.predict_deriv(Xs)
        h = 1e-6
        for q in range(3):
            e = np.zeros(3)
            e[q] = h
            fd = (gp.predict(Xs + e, unc=False)[0] - gp.predict(Xs - e, unc=False)[0]) / (2 * h)
            np.testing.assert_allclose(got[:, q], fd, rtol=1e-5, atol=1e-6)


def test_constant_mean_logpost_deriv_vs_finite_differences():
    """With a mean function the reference's own logpost_deriv cannot run here (calc_A_deriv under scipy >= 1.15), so the
    oracle's collected form of GaussianProcess.py:743-778 is pinned by central differences of its log-posterior (whose
    values ARE pinned by the cmean_* goldens)."""
    for kernel, nugget, mean in ((orc.SQEXP, 1e-3, "1"), (orc.MAT52, "fit", "1"), (orc.SQEXP, "fit", "x[0]"),
                                 (orc.MAT52, 1e-3, "-1 + x[0]*x[1]")):
        X, Y, _ = orc.make_workload(70, 2, 1, 5, seed=12)
        gp = orc.OracleGP(X, Y[0] + 3.0 - X[:, 0], kernel=kernel, nugget=nugget, mean=mean)
        theta = np.array([0.6, 0.9, 0.1] + ([-6.0] if nugget == "fit" else []))
        got = gp.logpost_deriv(theta)
        for i in range(theta.size):
            e = np.zeros(theta.size)
            e[i] = 1e-4
            fd = (gp.logposterior(theta + e) - gp.logposterior(theta - e)) / 2e-4
            assert_allclose(got[i], fd, rtol=1e-6, atol=1e-8)


def test_formula_mean_predict_deriv_vs_finite_differences():
    """Predictive derivative with a formula mean: kernel share on K^-1 (y - H beta) plus d(H* beta)/dx*."""
    X, Y, Xs = orc.make_workload(60, 3, 1, 9, seed=6)
    gp = orc.OracleGP(X, Y[0] + 2.0 * X[:, 0], nugget=1e-4,
                      mean="x[0] + x[1]:x[2] + I(x[0]**2) + np.sin(x[1]) + x[2]").fit(np.array([0.7, 1.1, 0.9, 0.3]))
    got = gp.predict_deriv(Xs)
    h = 1e-6
    for q in range(3):
        e = np.zeros(3)
        e[q] = h
        fd = (gp.predict(Xs + e, unc=False)[0] - gp.predict(Xs - e, unc=False)[0]) / (2 * h)
        assert_allclose(got[:, q], fd, rtol=1e-5, atol=1e-6)
