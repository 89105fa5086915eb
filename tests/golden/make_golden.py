"""Generate golden fixtures from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference through oracle/refstub.py, runs the reference CPU GaussianProcess /
MultiOutputGP on small seeded problems and stores inputs + outputs in tests/golden/*.npz.
The fixtures travel to the GPU box; the reference does not.  Fixture sizes are kept small
(a few hundred KB in total).
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import refstub  # noqa: E402

warnings.simplefilter("ignore")
mogp = refstub.import_reference()
from mogp_emulator.Kernel import SquaredExponential, Matern52  # noqa: E402
from mogp_emulator.linalg.cholesky import jit_cholesky  # noqa: E402


def workload(n, d, n_out, m, seed):
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    Y = np.stack([np.sin(2.0 * X.sum(axis=1) + k) + 0.01 * rng.standard_normal(n) for k in range(n_out)])
    Xs = rng.random((m, d))
    return X, Y, Xs


def prior_params(gp):
    corr = []
    for p in gp.priors.corr:
        corr.append([getattr(p, "shape", np.nan), getattr(p, "scale", np.nan)])
    nug = gp.priors.nugget
    nug = [getattr(nug, "shape", np.nan), getattr(nug, "scale", np.nan)] if nug is not None else [np.nan, np.nan]
    return np.array(corr, dtype=np.float64), np.array(nug, dtype=np.float64)


ONLY = sys.argv[1:]      # optional name filters: `make_golden.py fmean` regenerates only the matching fixtures


def wanted(name):
    return not ONLY or any(k in name for k in ONLY)


def single_case(name, n, d, m, seed, kernel, nugget, theta, dup_rows=False, with_deriv=True, mean_fn=None):
    if not wanted(name):
        return
    X, Y, Xs = workload(n, d, 1, m, seed)
    y = Y[0]
    if dup_rows:
        X[1] = X[0]
        y[1] = y[0]
    kern = SquaredExponential() if kernel == "SquaredExponential" else Matern52()
    gp = mogp.GaussianProcess(X, y, mean=mean_fn, kernel=kern, nugget=nugget)
    gp.fit(theta)
    mean, var, _ = gp.predict(Xs)
    mean_nn, var_nn, _ = gp.predict(Xs, include_nugget=False)
    _, cov_full, _ = gp.predict(Xs, full_cov=True)
    out = dict(X=X, y=y, Xs=Xs, theta=np.array(theta), kernel=kernel,
               nugget_in=np.array(nugget if not isinstance(nugget, str) else np.nan),
               nugget_type=gp.nugget_type, nugget_out=np.array(gp.nugget),
               K=gp.get_K_matrix(), L=gp.Kinv.L, Kinv_t=gp.Kinv_t, logpost=np.array(gp.current_logpost),
               mean=mean, var=var, var_no_nugget=var_nn, cov_full=cov_full)
    if mean_fn is not None:
        out["mean_spec"] = mean_fn
        out["theta_mean"] = np.array(gp.theta.mean)
        out["Kinv_t_mean"] = gp.Kinv_t_mean
    corr, nug = prior_params(gp)
    out["prior_corr"] = corr
    out["prior_nugget"] = nug
    if with_deriv:
        out["deriv"] = gp.logpost_deriv(theta)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "logpost", gp.current_logpost, "nugget", gp.nugget)


def multi_case(name, n, d, n_out, m, seed, kernel, nugget, theta_scale):
    if not wanted(name):
        return
    X, Y, Xs = workload(n, d, n_out, m, seed)
    rng = np.random.default_rng(seed + 1000)
    n_params = d + 1 + (1 if nugget == "fit" else 0)
    thetas = theta_scale * rng.standard_normal((n_out, n_params))
    if nugget == "fit":
        thetas[:, -1] = -12.0 + rng.standard_normal(n_out)
    gp = mogp.MultiOutputGP(X, Y, kernel=kernel, nugget=nugget)
    gp.fit(thetas)
    mean, var, _ = gp.predict(Xs, processes=1)
    logposts = np.array([em.current_logpost for em in gp.emulators])
    nuggets = np.array([em.nugget for em in gp.emulators])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), X=X, Y=Y, Xs=Xs, thetas=thetas, kernel=kernel,
                        nugget_type=gp.emulators[0].nugget_type,
                        nugget_in=np.array(nugget if not isinstance(nugget, str) else np.nan),
                        mean=mean, var=var, logposts=logposts, nuggets=nuggets)
    print(name, "logposts", logposts)


def validation_case(name, n, d, m, seed, kernel, nugget, theta, n_out=1, mean_fn=None):
    """validation.py of the reference on a fitted GP / MultiOutputGP: standard errors, pivoted errors, Mahalanobis."""
    if not wanted(name):
        return
    from mogp_emulator.validation import standard_errors, pivoted_errors, mahalanobis
    X, Y, Xv = workload(n, d, n_out, m, seed)
    rng = np.random.default_rng(seed + 500)
    Yv = np.stack([np.sin(2.0 * Xv.sum(axis=1) + k) + 0.05 * rng.standard_normal(m) for k in range(n_out)])
    kern = SquaredExponential() if kernel == "SquaredExponential" else Matern52()
    if n_out == 1:
        gp = mogp.GaussianProcess(X, Y[0], mean=mean_fn, kernel=kern, nugget=nugget)
        gp.fit(theta)
        yv = Yv[0]
        se, sp = standard_errors(gp, Xv, yv)
        pe, pp = pivoted_errors(gp, Xv, yv)
        out = dict(std_err=se, std_idx=sp, piv_err=pe, piv_idx=pp, mahal=np.array(mahalanobis(gp, Xv, yv)),
                   mahal_scaled=np.array(mahalanobis(gp, Xv, yv, scaled=True)), y=Y[0], yv=yv)
    else:
        gp = mogp.MultiOutputGP(X, Y, mean=mean_fn, kernel=kernel, nugget=nugget)
        gp.fit(np.tile(theta, (n_out, 1)) + 0.1 * np.arange(n_out)[:, None])
        se = standard_errors(gp, Xv, Yv)
        pe = pivoted_errors(gp, Xv, Yv)
        out = dict(std_err=np.array([e[0] for e in se]), std_idx=np.array([e[1] for e in se]),
                   piv_err=np.array([e[0] for e in pe]), piv_idx=np.array([e[1] for e in pe]),
                   mahal=np.array(mahalanobis(gp, Xv, Yv)), mahal_scaled=np.array(mahalanobis(gp, Xv, Yv, scaled=True)),
                   y=Y, yv=Yv)
    out.update(X=X, Xv=Xv, theta=np.array(theta), kernel=kernel, n_out=np.array(n_out),
               nugget_in=np.array(nugget if not isinstance(nugget, str) else np.nan),
               nugget_type="fixed" if not isinstance(nugget, str) else nugget)
    if mean_fn is not None:
        out["mean_spec"] = mean_fn
    np.savez_compressed(os.path.join(HERE, "validation", name + ".npz"), **out)
    print(name, "mahalanobis", out["mahal"], out["mahal_scaled"])


def history_case(name, n, d, m, seed, kernel, nugget, theta, n_out):
    """HistoryMatching of the reference on a fitted GP / MultiOutputGP: implausibility for several ranks / discrepancies,
    NROY and RO index sets."""
    if not wanted(name):
        return
    from mogp_emulator.HistoryMatching import HistoryMatching
    X, Y, Xq = workload(n, d, n_out, m, seed)
    rng = np.random.default_rng(seed + 700)
    kern = SquaredExponential() if kernel == "SquaredExponential" else Matern52()
    if n_out == 1:
        gp = mogp.GaussianProcess(X, Y[0], kernel=kern, nugget=nugget)
        gp.fit(theta)
        obs = [0.3, 0.02]
        thetas = np.array(theta)
    else:
        gp = mogp.MultiOutputGP(X, Y, kernel=kernel, nugget=nugget)
        thetas = np.tile(theta, (n_out, 1)) + 0.1 * np.arange(n_out)[:, None]
        gp.fit(thetas)
        obs = [0.5 * rng.standard_normal(n_out), 0.01 + 0.05 * rng.random(n_out)]
    out = dict(X=X, Y=Y, Xq=Xq, thetas=thetas, kernel=kernel, n_out=np.array(n_out), nugget_in=np.array(nugget),
               obs_val=np.atleast_1d(obs[0]), obs_var=np.atleast_1d(obs[1]))
    ranks = [0] if n_out == 1 else [0, 1, n_out - 1]
    for r in ranks:
        hm = HistoryMatching(gp=gp, obs=obs, coords=Xq, threshold=2.0)
        out["I_rank%d" % r] = hm.get_implausibility(rank=r)
        out["NROY_rank%d" % r] = np.array(hm.get_NROY(rank=r), dtype=np.int64)
        out["RO_rank%d" % r] = np.array(hm.get_RO(rank=r), dtype=np.int64)
    hm = HistoryMatching(gp=gp, obs=obs, coords=Xq)
    out["I_disc"] = hm.get_implausibility(discrepancy=0.3, rank=0)
    if n_out > 1:
        disc = 0.1 * (1.0 + np.arange(n_out))
        out["disc_vec"] = disc
        out["I_disc_vec"] = HistoryMatching(gp=gp, obs=obs, coords=Xq).get_implausibility(discrepancy=disc, rank=1)
    np.savez_compressed(os.path.join(HERE, "history", name + ".npz"), **out)
    print(name, "I range", out["I_rank0"].min(), out["I_rank0"].max(), "NROY", len(out["NROY_rank0"]))


def mice_case(name, n, d, seed, kernel, nugget, theta):
    """MICEFastGP.fast_predict of the reference at every training point."""
    if not wanted(name):
        return
    from mogp_emulator.SequentialDesign import MICEFastGP
    X, Y, _ = workload(n, d, 1, 4, seed)
    kern = SquaredExponential() if kernel == "SquaredExponential" else Matern52()
    gp = MICEFastGP(X, Y[0], kernel=kern, nugget=nugget)
    gp.fit(theta)
    # the reference's fast_predict still reads the attribute ``self.L`` that GaussianProcess no longer has at v0.7.2 (the
    # factor lives in ``self.Kinv.L``): supply it, the method itself runs unmodified
    gp.L = gp.Kinv.L
    var = np.array([float(np.squeeze(gp.fast_predict(i))) for i in range(n)])
    np.savez_compressed(os.path.join(HERE, "history", name + ".npz"), X=X, y=Y[0], theta=np.array(theta), kernel=kernel,
                        nugget_in=np.array(nugget), fast_var=var)
    print(name, "loo variance range", var.min(), var.max())


if __name__ == "__main__":
    os.makedirs(os.path.join(HERE, "history"), exist_ok=True)
    mice_case("mice_sqexp_n90_d2", 90, 2, 81, "SquaredExponential", 1e-4, [0.8, 0.5, 0.1])
    mice_case("mice_mat52_n150_d3", 150, 3, 82, "Matern52", 1e-6, [0.4, 0.6, 0.2, 0.3])
    history_case("hist_sqexp_single_n80_d2", 80, 2, 60, 71, "SquaredExponential", 1e-4, [0.8, 0.5, 0.1], 1)
    history_case("hist_mat52_multi_e4_n70_d3", 70, 3, 50, 72, "Matern52", 1e-4, [0.4, 0.6, 0.2, 0.0], 4)
    os.makedirs(os.path.join(HERE, "validation"), exist_ok=True)
    validation_case("valid_sqexp_fixed_n90_d2", 90, 2, 24, 61, "SquaredExponential", 1e-4, [0.8, 0.5, 0.1])
    validation_case("valid_mat52_adaptive_n80_d3_x0", 80, 3, 20, 62, "Matern52", "adaptive", [0.4, 0.6, 0.2, 0.0], mean_fn="x[0]")
    validation_case("valid_multi_sqexp_e3_n70_d2", 70, 2, 18, 63, "SquaredExponential", 1e-4, [0.7, 0.9, 0.0], n_out=3)
    # C1-shaped (BASELINE.json configs[0]) but n=200 to keep the stored L small
    single_case("sqexp_fixed_n200_d4", 200, 4, 64, 0, "SquaredExponential", 1e-6, [1.0, 1.0, 1.0, 1.0, 0.0])
    # spans two 128-blocks with ragged edge, non-trivial theta
    single_case("sqexp_fixed_n150_d3", 150, 3, 40, 5, "SquaredExponential", 1e-4, [0.3, -0.5, 1.2, 0.7])
    single_case("mat52_adaptive_n160_d5", 160, 5, 33, 3, "Matern52", "adaptive", [1.0, 0.5, 0.0, -0.5, 1.0, 0.2])
    # duplicated rows force the jitter path (expected nugget = 1e-6 * sigma^2)
    single_case("mat52_adaptive_dup_n96_d2", 96, 2, 20, 7, "Matern52", "adaptive", [0.5, 0.5, 0.4], dup_rows=True)
    single_case("sqexp_adaptive_dup_n140_d3", 140, 3, 20, 8, "SquaredExponential", "adaptive", [1.0, 1.0, 1.0, -0.3],
                dup_rows=True)
    single_case("sqexp_fit_n100_d2", 100, 2, 25, 11, "SquaredExponential", "fit", [0.8, 1.1, 0.1, -9.0])
    single_case("mat52_fit_n130_d3", 130, 3, 25, 12, "Matern52", "fit", [0.2, 0.4, 0.6, 0.3, -7.0])
    single_case("sqexp_fixed_n1_d2", 1, 2, 5, 13, "SquaredExponential", 1e-6, [0.0, 0.0, 0.0], with_deriv=False)
    single_case("sqexp_fixed_n2_d1", 2, 1, 6, 14, "SquaredExponential", 0.0, [0.5, 0.1], with_deriv=False)
    # constant mean function with the default (weak) mean priors: analytic mean parameter.  The
    # reference's own logpost_deriv fails for n_mean > 0 under scipy >= 1.15 (calc_A_deriv), so no gradient is stored.
    single_case("cmean_sqexp_fixed_n170_d3", 170, 3, 30, 41, "SquaredExponential", 1e-5, [0.4, 0.9, 0.6, 0.2],
                with_deriv=False, mean_fn="1")
    single_case("cmean_mat52_adaptive_n140_d2", 140, 2, 25, 42, "Matern52", "adaptive", [0.7, 0.3, -0.2],
                with_deriv=False, mean_fn="1")
    # formula mean functions: the reference's GP algebra on the design matrices patsy would build
    # (oracle/refstub.py feeds them from gp_oracle.DESIGN_FUNCTIONS; patsy is not installed)
    single_case("fmean_sqexp_fixed_n160_d3", 160, 3, 30, 43, "SquaredExponential", 1e-5, [0.5, 0.8, 0.3, 0.1],
                with_deriv=False, mean_fn="x[0]")
    single_case("fmean_mat52_adaptive_n150_d3", 150, 3, 28, 44, "Matern52", "adaptive", [0.6, 0.2, 0.4, -0.1],
                with_deriv=False, mean_fn="x[0] + x[1]:x[2] + I(x[0]**2) + np.sin(x[1]) + x[2]")
    single_case("fmean_sqexp_fit_n120_d2", 120, 2, 22, 45, "SquaredExponential", "fit", [0.9, 0.7, 0.2, -8.0],
                with_deriv=False, mean_fn="y ~ x[0] + x[1]")
    multi_case("multi_sqexp_fixed_e4_n120_d3", 120, 3, 4, 30, 20, "SquaredExponential", 1e-6, 0.5)
    multi_case("multi_mat52_adaptive_e3_n90_d2", 90, 2, 3, 17, 21, "Matern52", "adaptive", 0.5)
