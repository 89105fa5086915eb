"""History matching (mogp_emulator_b200/HistoryMatching.py; reference mogp_emulator/HistoryMatching.py): known answers of the
reference's own tests, reference-generated goldens (tests/golden/history/*.npz), the oracle restatement, the class over the
numpy test double on CPU, and -- marked gpu -- the same goldens through the CUDA path."""
import glob
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_equal

import gp_oracle as orc
from fake_device import FakeHandle

HIST = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "history", "hist_*.npz")))


@pytest.fixture()
def fake_gpu(monkeypatch):
    from mogp_emulator_b200 import libmogp
    monkeypatch.setattr(libmogp, "Handle", FakeHandle)
    monkeypatch.setattr(libmogp, "HAVE_LIBMOGP", True)
    monkeypatch.setattr(libmogp, "gpu_usable", lambda: True)
    import mogp_emulator_b200 as mogp
    return mogp


def test_known_answers_of_the_reference_tests():
    """tests/test_HistoryMatching.py:363-458 of the reference (explicit expectations, no GP needed)."""
    from mogp_emulator_b200 import HistoryMatching, PredictResult
    exp = PredictResult(mean=np.array([2.0, 10.0]), unc=np.array([0.0, 0.0]), deriv=np.array([[1.0, 2.0]]))
    hm = HistoryMatching(obs=[1.0, 1.0], expectations=exp)
    assert_allclose(hm.get_implausibility(), [1.0, 9.0])
    assert_allclose(hm.I, [1.0, 9.0])
    assert_allclose(hm.get_implausibility(1.0), [1.0 / np.sqrt(2.0), 9.0 / np.sqrt(2.0)])
    assert_allclose(orc.implausibility(exp.mean, exp.unc, 1.0, 1.0, 1.0), hm.I)
    with pytest.raises(ValueError):
        HistoryMatching(expectations=exp).get_implausibility()                  # no observations
    with pytest.raises(AssertionError):
        HistoryMatching(obs=[1.0, 1.0], expectations=exp).get_implausibility(-1.0)
    exp2 = PredictResult(mean=np.array([[2.0, 10.0], [4.0, 6.0]]), unc=np.array([[0.5, 0.5], [0.5, 0.5]]), deriv=None)
    hm = HistoryMatching(obs=[[1.0, 5.0], 0.5], expectations=exp2)
    assert_allclose(hm.get_implausibility(), [1.0, 1.0])
    assert_allclose(hm.get_implausibility(1.0), [1.0 / np.sqrt(2.0)] * 2)
    assert_allclose(hm.get_implausibility(np.array([1.0, 1.0])), [1.0 / np.sqrt(2.0)] * 2)
    assert_allclose(orc.implausibility(exp2.mean, exp2.unc, [1.0, 5.0], [0.5, 0.5], [1.0, 1.0]), hm.I)
    # NROY / RO with the default threshold 3 (test_HistoryMatching.py:460-500)
    hm = HistoryMatching(obs=[1.0, 1.0], expectations=exp)
    assert hm.get_NROY() == [0] and hm.get_RO() == [1] and hm.threshold == 3.0
    # bookkeeping
    assert hm.ncoords == 2 and hm.ndim is None and hm.get_n_obs() == 1
    assert "History Matching tools created with" in str(hm) and "of length 1" in str(hm)
    hm.set_threshold(0.5)
    hm.I = None
    assert hm.get_NROY() == []
    wide = PredictResult(mean=np.zeros((2, 5)), unc=np.ones((2, 5)), deriv=None)
    assert HistoryMatching(obs=[[0.0, 1.0], [1.0, 1.0]], expectations=wide).ncoords == 5      # (outputs, points): points counted


def test_argument_checks():
    from mogp_emulator_b200 import HistoryMatching, PredictResult
    hm = HistoryMatching()
    assert hm.gp is None and hm.obs is None and hm.coords is None and hm.threshold == 3.0
    with pytest.raises(TypeError):
        hm.set_gp(object())
    with pytest.raises(ValueError):
        hm.set_obs([1.0, 2.0, 3.0])
    with pytest.raises(TypeError):
        hm.set_obs("a")
    with pytest.raises(AssertionError):
        hm.set_obs([1.0, -1.0])
    with pytest.raises(TypeError):
        hm.set_coords([[1.0, 2.0]])
    with pytest.raises(TypeError):
        hm.set_expectations((np.zeros(2), np.ones(2), None))                    # a plain tuple is not a PredictResult
    with pytest.raises(ValueError):
        hm.set_expectations(PredictResult(mean=np.zeros(2), unc=np.ones(3), deriv=None))
    with pytest.raises(AssertionError):
        hm.set_threshold(-1.0)
    hm.set_obs(2.0)
    assert_equal(hm.obs[0], [2.0])
    assert_equal(hm.obs[1], [0.0])
    hm.set_obs([np.array([1.0, 2.0])])
    assert_equal(hm.obs[1], [0.0])
    hm.set_coords(np.linspace(0.0, 1.0, 7))
    assert hm.coords.shape == (7, 1) and hm.ncoords == 7 and hm.ndim == 1
    with pytest.raises(ValueError):
        hm.get_implausibility()                                                  # coords but no GP, no expectations
    hm.set_coords(None)
    assert hm.coords is None


def _golden_gp(mogp, g):
    n_out = int(g["n_out"])
    if n_out == 1:
        gp = mogp.GaussianProcessGPU(g["X"], g["Y"][0], kernel=str(g["kernel"]), nugget=float(g["nugget_in"]))
        gp.fit(g["thetas"])
    else:
        gp = mogp.MultiOutputGP_GPU(g["X"], g["Y"], kernel=str(g["kernel"]), nugget=float(g["nugget_in"]))
        gp.fit(g["thetas"])
    obs = [float(g["obs_val"][0]), float(g["obs_var"][0])] if n_out == 1 else [g["obs_val"], g["obs_var"]]
    return gp, obs, n_out


def _check_against_golden(mogp, g, rtol):
    gp, obs, n_out = _golden_gp(mogp, g)
    ranks = [0] if n_out == 1 else [0, 1, n_out - 1]
    for r in ranks:
        hm = mogp.HistoryMatching(gp=gp, obs=obs, coords=g["Xq"], threshold=2.0)
        assert hm.ncoords == len(g["Xq"]) and hm.ndim == g["Xq"].shape[1]
        assert_allclose(hm.get_implausibility(rank=r), g["I_rank%d" % r], rtol=rtol)
        # index sets: identical except where a score sits within rtol of the threshold
        margin = np.abs(g["I_rank%d" % r] - 2.0) > 10 * rtol * 2.0
        nroy = np.zeros(hm.ncoords, bool)
        nroy[hm.get_NROY(rank=r)] = True
        want = np.zeros(hm.ncoords, bool)
        want[g["NROY_rank%d" % r]] = True
        assert_equal(nroy[margin], want[margin])
        assert sorted(hm.get_NROY() + hm.get_RO()) == list(range(hm.ncoords))
    hm = mogp.HistoryMatching(gp=gp, obs=obs, coords=g["Xq"])
    assert_allclose(hm.get_implausibility(discrepancy=0.3, rank=0), g["I_disc"], rtol=rtol)
    if n_out > 1:
        assert_allclose(hm.get_implausibility(discrepancy=g["disc_vec"], rank=1), g["I_disc_vec"], rtol=rtol)
    with pytest.raises(ValueError):                                               # GP + coords AND explicit expectations
        mogp.HistoryMatching(gp=gp, obs=obs, coords=g["Xq"],
                             expectations=gp.predict(g["Xq"], deriv=False)).get_implausibility()
    gp.close()


@pytest.mark.parametrize("path", HIST)
def test_oracle_matches_reference_golden(path):
    g = np.load(path)
    n_out = int(g["n_out"])
    thetas = np.atleast_2d(g["thetas"])
    preds = [orc.OracleGP(g["X"], g["Y"][k], kernel=str(g["kernel"]), nugget=float(g["nugget_in"])).fit(thetas[k]).predict(g["Xq"])
             for k in range(n_out)]
    mean, var = np.array([p[0] for p in preds]), np.array([p[1] for p in preds])
    for r in ([0] if n_out == 1 else [0, 1, n_out - 1]):
        assert_allclose(orc.implausibility(mean, var, g["obs_val"], g["obs_var"], rank=r), g["I_rank%d" % r], rtol=1e-6)
    assert_allclose(orc.implausibility(mean, var, g["obs_val"], g["obs_var"], 0.3, rank=0), g["I_disc"], rtol=1e-6)


@pytest.mark.parametrize("path", HIST)
def test_front_end_over_the_test_double(fake_gpu, path):
    _check_against_golden(fake_gpu, np.load(path), rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("path", HIST)
def test_gpu_history_matching_matches_reference_golden(path):
    import mogp_emulator_b200 as mogp
    _check_against_golden(mogp, np.load(path), rtol=1e-5)


@pytest.mark.gpu
def test_gpu_history_matching_many_query_points():
    """A wave of 20000 query points over 24 outputs (the int8 predict path): implausibility against the oracle's posterior
    for two of the outputs' worth of scores, and NROY + RO partition the query set."""
    import mogp_emulator_b200 as mogp
    X, Y, Xq = orc.make_workload(400, 4, 24, 20000, seed=77)
    thetas = np.tile(np.array([0.8, 1.0, 0.9, 1.1, 0.0]), (24, 1))
    gp = mogp.MultiOutputGP_GPU(X, Y, nugget=1e-5)
    gp.fit(thetas)
    rng = np.random.default_rng(3)
    obs = [0.4 * rng.standard_normal(24), np.full(24, 0.05)]
    hm = mogp.HistoryMatching(gp=gp, obs=obs, coords=Xq, threshold=3.0)
    gp.timings(reset=True)
    I = hm.get_implausibility(rank=0)
    assert gp.timings()["i8_block_rows"] == 4
    post = gp.predict(Xq[:300], deriv=False)
    refs = [orc.OracleGP(X, Y[k], nugget=1e-5, priors="weak").fit(thetas[k]).predict(Xq[:300]) for k in range(24)]
    want = orc.implausibility(np.array([r[0] for r in refs]), np.array([r[1] for r in refs]), obs[0], obs[1], rank=0)
    assert_allclose(I[:300], want, rtol=1e-5)
    assert_allclose(post.mean[5], refs[5][0], rtol=1e-6, atol=1e-8)
    assert sorted(hm.get_NROY(rank=0) + hm.get_RO(rank=0)) == list(range(20000))
    gp.close()
