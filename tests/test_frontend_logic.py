"""Host-side logic of the front-end classes on CPU: GaussianProcessGPU / MultiOutputGP_GPU / fit_GP_MAP run over the
numpy TEST DOUBLE of the device handle (tests/fake_device.py, installed by monkeypatching -- the product has no CPU
path) and are compared with the oracle.  What this covers is everything *above* the C ABI: argument checking and
exception types, the mean-function algebra (constant and formula means), fit-status bookkeeping, shapes, pickling, the
lock-step MAP driver.  The CUDA kernels themselves are covered by tests/test_gpu_parity.py (``-m gpu``)."""
import pickle

import numpy as np
import pytest
from numpy.testing import assert_allclose

import gp_oracle as orc
from fake_device import FakeHandle

FORMULA6 = "x[0] + x[1]:x[2] + I(x[0]**2) + np.sin(x[1]) + x[2]"


@pytest.fixture()
def fake_gpu(monkeypatch):
    from mogp_emulator_b200 import libmogp
    monkeypatch.setattr(libmogp, "Handle", FakeHandle)
    monkeypatch.setattr(libmogp, "HAVE_LIBMOGP", True)
    monkeypatch.setattr(libmogp, "gpu_usable", lambda: True)
    import mogp_emulator_b200 as mogp
    return mogp


def _workload(n=80, d=3, e=1, m=15, seed=7):
    X, Y, Xs = orc.make_workload(n, d, e, m, seed=seed)
    Y = Y + 1.5 - 2.0 * X[:, 0] + 0.8 * X[:, 1] * X[:, -1]
    return X, Y, Xs


def test_constructor_checks_without_device():
    """No library / no device -> RuntimeError (GaussianProcessGPU.py:223-229); nothing falls back to the CPU."""
    import mogp_emulator_b200 as mogp
    from mogp_emulator_b200 import libmogp
    if libmogp.gpu_usable():
        pytest.skip("a device is present")
    X, Y, _ = _workload()
    with pytest.raises(RuntimeError):
        mogp.GaussianProcessGPU(X, Y[0])
    with pytest.raises(RuntimeError):
        mogp.MultiOutputGP_GPU(X, Y)


@pytest.mark.parametrize("mean,kernel,nugget", [
    (None, "SquaredExponential", 1e-5), ("1", "Matern52", "adaptive"), ("x[0]", "SquaredExponential", "fit"),
    (FORMULA6, "Matern52", 1e-4), ("-1 + x[0]*x[1]", "SquaredExponential", 1e-4), ("y ~ x[0] + x[1]", "Matern52", "fit"),
])
def test_single_output_front_end_against_oracle(fake_gpu, mean, kernel, nugget):
    mogp = fake_gpu
    X, Y, Xs = _workload()
    y = Y[0]
    theta = np.array([0.5, 0.8, 0.3, 0.1] + ([-7.5] if nugget == "fit" else []))
    gp = mogp.GaussianProcessGPU(X, y, mean=mean, kernel=kernel, nugget=nugget)
    ref = orc.OracleGP(X, y, mean=mean, kernel=kernel, nugget=nugget)
    assert gp.n_params == ref.n_params and gp.n_mean == ref.n_mean and gp.mean == mean
    assert gp.theta.get_n_data() == ref.n_params and not gp.theta.data_has_been_set()
    assert gp.current_logpost is None and gp.L is None and gp.Kinv_t is None
    with pytest.raises(ValueError):
        gp.predict(Xs)
    with pytest.raises(RuntimeError):
        gp.fit(np.ones(ref.n_params + 1))
    assert_allclose(gp.logposterior(theta), ref.logposterior(theta), rtol=1e-10)
    assert_allclose(gp.logpost_deriv(theta), ref.logpost_deriv(theta), rtol=1e-7, atol=1e-9)
    assert_allclose(gp.theta.mean, ref.theta_mean, rtol=1e-8, atol=1e-10)
    assert_allclose(gp.Kinv_t, ref.Kinv_t, rtol=1e-9, atol=1e-9)
    assert_allclose(gp.Kinv_t_mean, ref.Kinv_t_mean, rtol=1e-7, atol=1e-8)
    assert_allclose(gp.nugget, ref.nugget, rtol=1e-14)
    res = gp.predict(Xs)
    rmean, rvar = ref.predict(Xs)
    assert_allclose(res.mean, rmean, rtol=1e-8, atol=1e-9)
    assert_allclose(res.unc, rvar, rtol=1e-6, atol=1e-10)
    assert_allclose(res.deriv, ref.predict_deriv(Xs), rtol=1e-6, atol=1e-7)
    assert_allclose(gp.predict(Xs, include_nugget=False).unc, ref.predict(Xs, include_nugget=False)[1], rtol=1e-6, atol=1e-10)
    _, rcov = ref.predict(Xs, full_cov=True)
    assert_allclose(gp.predict(Xs, deriv=False, full_cov=True).unc, rcov, rtol=1e-6, atol=1e-10)
    assert gp.predict(Xs, unc=False, deriv=False).unc is None and gp.predict(Xs, unc=False, deriv=False).deriv is None
    assert_allclose(gp(Xs), rmean, rtol=1e-8, atol=1e-9)
    mean1, unc1, deriv1 = gp.predict(Xs[0])                    # a single point given as a 1-D array
    assert mean1.shape == (1,) and unc1.shape == (1,) and deriv1.shape == (1, 3)
    # pickling drops the device object and refits on load (GaussianProcessGPU.py:656-667)
    clone = pickle.loads(pickle.dumps(gp))
    assert_allclose(clone.predict(Xs).mean, res.mean, rtol=1e-12)
    assert_allclose(clone.current_logpost, gp.current_logpost, rtol=1e-12)
    gp.theta = None
    assert gp.current_logpost is None and not gp.theta.data_has_been_set()
    gp.close()


def test_argument_errors_follow_the_reference(fake_gpu):
    mogp = fake_gpu
    X, Y, Xs = _workload()
    for bad in (1.0, "x[6]", "x[0] +", "(x[0] + x[1]):x[2]"):
        with pytest.raises(ValueError):
            mogp.GaussianProcessGPU(X, Y[0], mean=bad)
    with pytest.raises(ValueError):
        mogp.GaussianProcessGPU(X, Y[0], kernel="blah")
    with pytest.raises(ValueError):
        mogp.GaussianProcessGPU(X, Y[0], nugget="blah")
    with pytest.raises(ValueError):
        mogp.GaussianProcessGPU(X, Y[0], nugget=-1.0)
    with pytest.raises(TypeError):
        mogp.GaussianProcessGPU(X, Y[0], nugget=[1.0])
    with pytest.raises(AssertionError):
        mogp.GaussianProcessGPU(X, Y[0][:-1])
    with pytest.warns(DeprecationWarning):
        mogp.GaussianProcessGPU(X, Y[0], mean="x[0]", inputdict={"a": 0})
    gp = mogp.GaussianProcessGPU(X, Y[0], nugget=1e-6)
    with pytest.raises(mogp.GPUUnavailableError):
        gp.logpost_hessian(np.zeros(4))
    gp.fit(np.zeros(4))
    with pytest.raises(AssertionError):
        gp.predict(np.ones((4, 2)))
    # identical rows and no nugget: not positive definite -> RuntimeError, emulator left unfit
    Xd = X.copy()
    Xd[1] = Xd[0]
    gp = mogp.GaussianProcessGPU(Xd, Y[0], nugget=0.0)
    with pytest.raises(RuntimeError):
        gp.fit(np.zeros(4))
    assert not gp.theta.data_has_been_set()


def test_multi_output_front_end_with_formula_mean(fake_gpu):
    mogp = fake_gpu
    X, Y, Xs = _workload(n=70, e=4, m=12, seed=9)
    thetas = np.array([[0.5, 0.8, 0.4, 0.1], [0.9, 0.4, 0.6, 0.3], [0.2, 0.3, 0.5, -0.1], [0.1, 0.2, 0.3, 0.4]])
    mo = mogp.MultiOutputGP_GPU(X, Y, mean=FORMULA6, nugget=1e-5)
    assert mo.n_emulators == 4 and mo.get_indices_fit() == [] and mo.n_params == [4] * 4
    with pytest.raises(ValueError):
        mo.predict(Xs)
    for i in (0, 1, 3):
        mo.fit_emulator(i, thetas[i])
    assert mo.get_indices_not_fit() == [2]
    r = mo.predict(Xs, allow_not_fit=True)
    assert r.mean.shape == (4, 12) and r.unc.shape == (4, 12) and r.deriv.shape == (4, 12, 3)
    assert np.all(np.isnan(r.mean[2])) and np.all(np.isnan(r.unc[2]))
    refs = [orc.OracleGP(X, Y[i], mean=FORMULA6, nugget=1e-5).fit(thetas[i]) for i in range(4)]
    for i in (0, 1, 3):
        rm, rv = refs[i].predict(Xs)
        assert_allclose(r.mean[i], rm, rtol=1e-8, atol=1e-9)
        assert_allclose(r.unc[i], rv, rtol=1e-6, atol=1e-10)
        assert_allclose(r.deriv[i], refs[i].predict_deriv(Xs), rtol=1e-6, atol=1e-7)
        assert_allclose(mo.thetas[i].mean, refs[i].theta_mean, rtol=1e-8, atol=1e-10)
        assert_allclose(mo.logposterior(i), refs[i].current_logpost, rtol=1e-10)
    mo.fit(thetas)
    got = mo.logpost_and_deriv_batch([3, 0, 2], thetas[[3, 0, 2]])
    for i in (3, 0, 2):
        assert_allclose(got[i][0], refs[i].current_logpost, rtol=1e-10)
        assert_allclose(got[i][1], refs[i].logpost_deriv(thetas[i]), rtol=1e-7, atol=1e-9)
    _, rcov = refs[1].predict(Xs, full_cov=True)
    assert_allclose(mo.predict(Xs, deriv=False, full_cov=True).unc[1], rcov, rtol=1e-6, atol=1e-10)
    mo.reset_fit_status()
    assert mo.get_indices_fit() == []
    with pytest.raises(RuntimeError):
        mo.fit(np.zeros((3, 4)))
    with pytest.raises(RuntimeError):
        mo.fit(np.zeros((4, 5)))


@pytest.mark.parametrize("mean", [None, "x[0]"])
def test_fit_GP_MAP_over_the_front_end(fake_gpu, mean):
    """The MAP driver (scipy L-BFGS-B over the classes' logposterior / logpost_deriv; lock-step batched search for the
    multi-output class) lands where the same optimiser lands on the oracle."""
    from scipy.optimize import minimize
    mogp = fake_gpu
    X, Y, _ = _workload(n=50, d=2, e=3, m=4, seed=21)
    ref = orc.OracleGP(X, Y[0], nugget=1e-4, mean=mean)
    rr = minimize(ref.logposterior, np.zeros(3), method="L-BFGS-B", jac=ref.logpost_deriv)
    gp = mogp.fit_GP_MAP(mogp.GaussianProcessGPU(X, Y[0], nugget=1e-4, mean=mean), n_tries=1, theta0=np.zeros(3))
    assert_allclose(gp.current_logpost, rr["fun"], rtol=1e-7)
    mo = mogp.fit_GP_MAP(mogp.MultiOutputGP_GPU(X, Y, nugget=1e-4, mean=mean), n_tries=1, theta0=np.zeros(3))
    assert mo.get_indices_not_fit() == []
    assert_allclose(mo.logposterior(0), rr["fun"], rtol=1e-6)
    with pytest.raises(NotImplementedError):
        mogp.fit_GP_MAP(gp, method="Nelder-Mead")


def test_nugget_setter_priors_forms_and_strings(fake_gpu):
    """Less-travelled host paths: the nugget setter rebuilds the device object (and refits when the parameter vector keeps
    its shape), priors given as a dict / GPPriors / per-emulator list, theta given as a GPParams, __str__."""
    mogp = fake_gpu
    X, Y, Xs = _workload(n=60, d=2, e=3, m=7, seed=13)
    gp = mogp.GaussianProcessGPU(X, Y[0], nugget=1e-4)
    assert gp.nugget_type == "fixed" and gp.nugget == 1e-4 and gp.n_params == 3
    assert str(gp) == "Gaussian Process with 60 training examples and 2 input variables"
    gp.fit(np.array([0.3, 0.6, 0.1]))
    before = gp.predict(Xs, deriv=False).unc
    gp.nugget = 1e-2                                   # same theta shape: refitted with the new nugget
    assert gp.nugget == 1e-2 and gp.theta.data_has_been_set()
    after = gp.predict(Xs, deriv=False).unc
    assert np.all(after > before)
    want = orc.OracleGP(X, Y[0], nugget=1e-2, priors="weak").fit(np.array([0.3, 0.6, 0.1])).predict(Xs)[1]
    assert_allclose(after, want, rtol=1e-8, atol=1e-12)
    gp.nugget = "fit"                                  # the parameter vector grows: fit status is dropped
    assert gp.nugget_type == "fit" and gp.n_params == 4 and not gp.theta.data_has_been_set()
    gp.nugget = "adaptive"
    assert gp.nugget_type == "adaptive" and gp.n_params == 3 and gp.nugget == 0.0
    with pytest.raises(ValueError):
        gp.nugget = -1.0
    # theta as a GPParams object
    other = mogp.GaussianProcessGPU(X, Y[0], nugget=1e-4)
    other.fit(np.array([0.3, 0.6, 0.1]))
    gp2 = mogp.GaussianProcessGPU(X, Y[0], nugget=1e-4)
    gp2.fit(other.theta)
    assert_allclose(gp2.current_logpost, other.current_logpost, rtol=1e-12)
    # priors: dict and GPPriors objects give the same log-posterior as the oracle with the same prior parameters
    pri = dict(corr=[mogp.InvGammaPrior(2.0, 1.0), mogp.WeakPrior()], cov=mogp.WeakPrior(), nugget=None, n_corr=2,
               nugget_type="fixed")
    gp3 = mogp.GaussianProcessGPU(X, Y[0], nugget=1e-4, priors=pri)
    gp4 = mogp.GaussianProcessGPU(X, Y[0], nugget=1e-4, priors=mogp.GPPriors(**pri))
    th = np.array([0.2, -0.1, 0.4])
    assert_allclose(gp3.logposterior(th), gp4.logposterior(th), rtol=1e-13)
    weak = mogp.GaussianProcessGPU(X, Y[0], nugget=1e-4, priors=dict(n_corr=2, nugget_type="fixed"))
    assert gp3.logposterior(th) != weak.logposterior(th)
    assert_allclose(weak.logposterior(th), orc.OracleGP(X, Y[0], nugget=1e-4, priors="weak").logposterior(th), rtol=1e-10)
    # multi-output: one priors object for all, or a list with one entry per emulator
    mo = mogp.MultiOutputGP_GPU(X, Y, nugget=1e-4, priors=[None, pri, dict(n_corr=2, nugget_type="fixed")])
    assert "3 emulators" in str(mo) and len(mo.priors) == 3
    mo.fit(np.tile(th, (3, 1)))
    assert_allclose(mo.logposterior(2), orc.OracleGP(X, Y[2], nugget=1e-4, priors="weak").logposterior(th), rtol=1e-10)
    assert_allclose(mo.logposterior(0), orc.OracleGP(X, Y[0], nugget=1e-4).logposterior(th), rtol=1e-10)
    with pytest.raises(AssertionError):
        mogp.MultiOutputGP_GPU(X, Y, nugget=1e-4, priors=[None, None])
    assert mo.nugget == 1e-4 and mo.nugget_type == "fixed" and mo.n_corr == [2, 2, 2]
