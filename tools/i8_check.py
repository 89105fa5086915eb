"""GPU check of the int8 / tcgen05 predict TRSM (MOGP_TRSM_I8 = 6 and 7 planes) against the FP64 DMMA path (MOGP_TRSM_I8=0) on
the same inputs, all outputs, and against the oracle on output 0; prints the error relative to the parity tolerance and the
device time of the TRSM phase.  Results of round 1: profiles/r01_i8_check.txt.
    python tools/i8_check.py [small|c3|all|c5rank|diag]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gp_oracle as orc  # noqa: E402
import mogp_emulator_b200 as mogp  # noqa: E402


def run(use_i8, X, Y, Xs, thetas, nugget, kernel, reps=1):
    os.environ["MOGP_TRSM_I8"] = str(int(use_i8))
    gp = mogp.MultiOutputGP_GPU(X, Y, kernel=kernel, nugget=nugget)
    gp.fit(thetas)
    best = None
    for _ in range(reps):
        gp.timings(reset=True)
        t0 = time.perf_counter()
        mean, var, _ = gp.predict(Xs, deriv=False)
        wall = time.perf_counter() - t0
        t = gp.timings()
        if best is None or t["trsm_ms"] < best[1]["trsm_ms"]:
            best = (wall, t)
    gp.close()
    return mean, var, best


def compare(tag, n, d, E, m, kernel, nugget, theta_corr, check_oracle=True, reps=1):
    X, Y, Xs = orc.make_workload(n, d, E, m, seed=2)
    thetas = np.tile(np.array([theta_corr] * d + [0.0]), (E, 1)) + 0.05 * np.arange(E)[:, None]
    print("%s: n=%d d=%d E=%d m=%d %s nugget=%g theta=%g" % (tag, n, d, E, m, kernel, nugget, theta_corr))
    m0, v0, b0 = run(0, X, Y, Xs, thetas, nugget, kernel, reps)
    ok = True
    ref = orc.OracleGP(X, Y[0], kernel=kernel, nugget=nugget, priors="weak").fit(thetas[0]) if check_oracle else None
    for planes in PLANES:
        ok &= compare_one(planes, m0, v0, b0, X, Y, Xs, thetas, nugget, kernel, reps, ref, m)
    return ok


PLANES = (6, 7)


def compare_one(planes, m0, v0, b0, X, Y, Xs, thetas, nugget, kernel, reps, ref, m):
    m1, v1, b1 = run(planes, X, Y, Xs, thetas, nugget, kernel, reps)
    print(" S=%d:" % planes)
    print("   DMMA : trsm %.2f ms (kstar %.2f ms)   i8: trsm %.2f ms (kstar %.2f ms)" %
          (b0[1]["trsm_ms"], b0[1]["kstar_ms"], b1[1]["trsm_ms"], b1[1]["kstar_ms"]))
    dv = np.abs(v1 - v0)
    rel = dv / np.maximum(np.abs(v0), 1e-300)
    print("   var i8 vs DMMA: max abs %.3e, max rel %.3e (var in [%.3e, %.3e]); mean identical: %s" %
          (dv.max(), rel.max(), v0.min(), v0.max(), bool(np.array_equal(m0, m1))))
    ok = np.allclose(v1, v0, rtol=1e-4, atol=1e-4 * nugget)
    if not ok and os.environ.get("I8_VERBOSE"):
        per_out = dv.max(axis=1)
        print("   per-output max abs:", " ".join("%.1e" % x for x in per_out))
        worst = int(np.argmax(per_out))
        per_panel = [dv[worst, c0:c0 + 64].max() for c0 in range(0, m, 64)]
        print("   output %d per-panel max abs:" % worst, " ".join("%.1e" % x for x in per_panel[:40]))
        per_panel0 = [dv[0, c0:c0 + 64].max() for c0 in range(0, m, 64)]
        print("   output 0 per-panel max abs:", " ".join("%.1e" % x for x in per_panel0[:40]))
        c0 = int(np.argmax(per_panel)) * 64
        print("   worst panel columns:", " ".join("%.1e" % x for x in dv[worst, c0:c0 + 64]))
    tol = 1e-4 * np.abs(v0) + 1e-4 * nugget
    print("   worst error / tolerance: %.3f" % (dv / tol).max())
    if ref is not None:
        _, rv = ref.predict(Xs)
        print("   output 0 vs oracle: i8 max abs %.3e / DMMA max abs %.3e" % (np.abs(v1[0] - rv).max(), np.abs(v0[0] - rv).max()))
        ok = ok and np.allclose(v1[0], rv, rtol=1e-4, atol=1e-4 * nugget)
    print("   parity (rtol 1e-4, atol 1e-4 nugget):", "OK" if ok else "BROKEN")
    return ok


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "small"
    ok = True
    if what == "diag":
        os.environ["I8_VERBOSE"] = "1"
        compare("T2", 256, 3, 40, 600, "SquaredExponential", 1e-6, 1.0, check_oracle=False)
        compare("T3", 300, 3, 40, 600, "SquaredExponential", 1e-6, 1.0, check_oracle=False)
        compare("T4-one-wave", 512, 3, 2, 4700, "SquaredExponential", 1e-6, 1.0, check_oracle=False)
        compare("T8", 1000, 5, 8, 5000, "SquaredExponential", 1e-6, 1.0, check_oracle=False)
    if what in ("small", "all"):
        ok &= compare("small", 300, 3, 40, 600, "SquaredExponential", 1e-6, 1.0)
        ok &= compare("medium", 1000, 5, 8, 5000, "SquaredExponential", 1e-6, 1.0)
        ok &= compare("ill-conditioned", 1153, 4, 6, 4000, "Matern52", 2e-7, -1.0)     # just above the nugget gate (1e-7 sigma2)
    if what in ("c3", "all"):
        ok &= compare("C3", 4096, 10, 32, 10000, "SquaredExponential", 1e-6, 1.0, check_oracle=False, reps=3)
    if what == "c5rank":       # one rank's share of C5: 32 outputs x n=8192 x d=15, m=10000
        PLANES = (7,)
        ok &= compare("C5/rank", 8192, 15, 32, 10000, "SquaredExponential", 1e-6, 1.0, check_oracle=False, reps=2)
    sys.exit(0 if ok else 1)
