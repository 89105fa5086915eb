"""Exact emulation (oracle/i8_emulation.py) of the int8 predict path across the families that stress it: what the error of
the designed arithmetic is relative to the parity bar (rtol 1e-4 var + atol 1e-4 nugget), what a rigorous a-priori bound would
say, and what the a-posteriori check (32 sampled test points, 1 % of the bar) decides.  Runs on the CPU.
usage: python tools/i8_gate_study.py > profiles/r02_i8_gate_study.txt"""
import os, sys
import numpy as np
import scipy.linalg
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gp_oracle as orc
import i8_emulation as emu

rng = np.random.default_rng(7)


def inputs(kind, n, d):
    if kind == "uniform":
        return rng.random((n, d))
    if kind == "clustered":
        c = rng.random((8, d))
        return c[rng.integers(0, 8, n)] + 0.02 * rng.standard_normal((n, d))
    if kind == "near-duplicates":
        X = rng.random((n, d))
        X[n // 2:] = X[:n - n // 2] + 1e-6 * rng.standard_normal((n - n // 2, d))
        return X
    if kind == "grid-1d":
        return np.linspace(0.0, 1.0, n).reshape(-1, 1)
    raise ValueError(kind)


CASES = [("uniform", 900, 2, orc.SQEXP, 0.5), ("uniform", 900, 3, orc.SQEXP, 1.0), ("uniform", 768, 10, orc.SQEXP, 1.0),
         ("clustered", 800, 3, orc.SQEXP, 1.0), ("near-duplicates", 600, 3, orc.MAT52, 0.0), ("grid-1d", 700, 1, orc.SQEXP, 4.0),
         ("uniform", 900, 2, orc.MAT52, -1.0)]
print("# family n d kernel theta | nugget/sigma2 | cond(K) | max err / parity bar (S=7) | a-priori bound / bar | check ratio on 32 samples (accept <= 1) | decision")
for kind, n, d, kernel, th in CASES:
    X = inputs(kind, n, d)
    Xs = rng.random((256, d))
    theta = np.full(d, th)
    for nugget in (1e-6, 1e-8, 1e-10, 0.0):
        K = orc.kernel_f(X, X, theta, kernel) + nugget * np.eye(n)
        try:
            L = np.linalg.cholesky(K)
        except np.linalg.LinAlgError:
            print(kind, n, d, kernel, th, "| %g | not positive definite in FP64" % nugget)
            continue
        Ks = orc.kernel_f(X, Xs, theta, kernel)
        V = scipy.linalg.solve_triangular(L, Ks, lower=True)
        ref = 1.0 + nugget - np.sum(V * V, axis=0)
        got = emu.trsm_variance(L, Ks, 1.0, nugget, 7)
        bar = 1e-4 * np.abs(ref) + 1e-4 * nugget + 1e-300
        err = np.abs(got - ref)
        e = emu.scale_exponent(1.0, nugget)
        sv = np.linalg.svd(L, compute_uv=False)
        # rigorous worst case: every entry of T off by n 2^(2e - 7S) (dropped pairs + slicing), amplified by ||L^-1||, times 2 ||V_c||
        apriori = 2.0 * np.sqrt(n) * n * 2.0 ** (2 * e - 49) / sv[-1]
        samples = ((2 * np.arange(32) + 1) * 256) // 64
        allowed = 0.01 * (1e-4 * np.abs(ref[samples]) + 1e-4 * nugget) + 256 * np.finfo(float).eps * (1.0 + nugget)
        ratio = np.max(err[samples] / allowed)
        print(kind, n, d, kernel, th, "| %g | %.1e | %.2e | %.1e | %.2e | %s" % (nugget, (sv[0] / sv[-1]) ** 2, np.max(err / bar), apriori / np.min(bar),
                                                                                   ratio, "int8" if ratio <= 1.0 else "FP64 fall-back"))
