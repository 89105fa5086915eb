"""MICEFastGP.fast_predict (mogp_emulator_b200/SequentialDesign.py; reference SequentialDesign.py:683-748): the oracle's
restatement of the reference's Woodbury route against reference-generated goldens, the identity 1 / (K^-1)_ii the device
primitive uses, the class over the numpy test double on CPU, and -- marked gpu -- the CUDA path (mogp_loo_variance)."""
import glob
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

import gp_oracle as orc
from fake_device import FakeHandle

MICE = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "history", "mice_*.npz")))


@pytest.mark.parametrize("path", MICE)
def test_oracle_and_identity_match_reference_golden(path):
    g = np.load(path)
    gp = orc.OracleGP(g["X"], g["y"], kernel=str(g["kernel"]), nugget=float(g["nugget_in"])).fit(g["theta"])
    for i in (0, 1, len(g["y"]) // 2, len(g["y"]) - 1):
        assert_allclose(orc.mice_fast_predict(gp, i), g["fast_var"][i], rtol=1e-6, atol=1e-12)
    loo = 1.0 / np.diag(orc.cho_solve(gp.L, np.eye(gp.n)))          # Schur complement == 1 / (K^-1)_ii
    assert_allclose(loo, g["fast_var"], rtol=1e-5, atol=1e-11)


def _check(mogp, g, rtol):
    gp = mogp.MICEFastGP(g["X"], g["y"], kernel=str(g["kernel"]), nugget=float(g["nugget_in"]))
    with pytest.raises(ValueError):
        gp.fast_predict(0)
    gp.fit(g["theta"])
    n = gp.n
    for i in (0, 3, n - 1):
        v = gp.fast_predict(i)
        assert v.shape == (1,)
        assert_allclose(v[0], g["fast_var"][i], rtol=rtol, atol=1e-11)
    assert_allclose(gp.loo_variances(), g["fast_var"], rtol=rtol, atol=1e-11)
    with pytest.raises(AssertionError):
        gp.fast_predict(n)
    with pytest.raises(AssertionError):
        gp.fast_predict(-1)
    gp.fit(g["theta"] + 0.2)                                           # a new fit invalidates the cached vector
    assert not np.allclose(gp.loo_variances(), g["fast_var"], rtol=1e-3)
    mean, var, _ = gp.predict(g["X"][:5])                              # still a full emulator
    assert mean.shape == (5,) and var.shape == (5,)
    gp.close()


@pytest.mark.parametrize("path", MICE)
def test_front_end_over_the_test_double(path, monkeypatch):
    from mogp_emulator_b200 import libmogp
    monkeypatch.setattr(libmogp, "Handle", FakeHandle)
    monkeypatch.setattr(libmogp, "HAVE_LIBMOGP", True)
    monkeypatch.setattr(libmogp, "gpu_usable", lambda: True)
    import mogp_emulator_b200 as mogp
    _check(mogp, np.load(path), rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("path", MICE)
def test_gpu_matches_reference_golden(path):
    import mogp_emulator_b200 as mogp
    _check(mogp, np.load(path), rtol=1e-5)


@pytest.mark.gpu
def test_gpu_loo_variances_at_size():
    """n = 1500 (12 block rows): all leave-one-out variances from one L^-1 against 1 / diag(K^-1) of the oracle."""
    import mogp_emulator_b200 as mogp
    X, Y, _ = orc.make_workload(1500, 6, 1, 4, seed=91)
    theta = np.array([0.9, 1.0, 1.1, 0.8, 1.0, 1.2, 0.1])
    gp = mogp.MICEFastGP(X, Y[0], nugget=1e-5)
    gp.fit(theta)
    ref = orc.OracleGP(X, Y[0], nugget=1e-5, priors="weak").fit(theta)
    want = 1.0 / np.diag(orc.cho_solve(ref.L, np.eye(1500)))
    assert_allclose(gp.loo_variances(), want, rtol=1e-6)
    gp.close()
