"""The Cholesky with its history products on the int8 tensor cores (chol_i8_kernel, MOGP_CHOL_I8=1) against the FP64 DMMA
factorisation (MOGP_CHOL_I8=0) on the same inputs: L, log-determinant, alpha, posterior means / variances, timings.
usage (under gpurun): python tools/chol_i8_check.py [case ...] > gpurun_out/chol_i8_check.txt"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_thetas
from mogp_emulator_b200 import libmogp

CASES = {
    # name: (n, d, outputs, m, kernel, nugget type, nugget, theta_corr, compare L)
    "small": (384, 4, 2, 200, 0, 2, 1e-6, 1.0, True),
    "mid": (1024, 6, 3, 500, 0, 2, 1e-6, 1.0, True),
    "ragged": (1000, 5, 2, 300, 1, 2, 1e-8, 0.0, True),
    "illcond": (2048, 3, 2, 500, 0, 2, 1e-8, -1.0, True),
    "c3": (4096, 10, 32, 10000, 0, 2, 1e-6, None, False),
    "c3x4": (4096, 10, 4, 10000, 0, 2, 1e-6, None, False),
    "c2": (4096, 10, 1, 1000, 0, 2, 1e-6, None, True),
    "c4": (16384, 20, 1, 1000, 1, 0, 0.0, None, False),
    "c5rank": (8192, 15, 8, 10000, 0, 2, 1e-6, None, False),        # 8 of the 32 outputs a rank of C5 holds
}


def run(case, mode):
    n, d, E, m, kern, nt, nug, tc, cmpL = CASES[case]
    os.environ["MOGP_CHOL_I8"] = mode
    X, Y, Xs = make_workload(n, d, E, m, 2)
    thetas = make_thetas(E, d)
    if tc is not None:
        thetas[:, :d] = tc
    h = libmogp.Handle(X, Y, kern, nt, nug)
    out = {}
    for rep in range(3):
        h.timings(reset=True)
        t0 = time.perf_counter()
        quad, logdet, nugs, status = h.fit(0, thetas)
        out["fit_wall_ms"] = (time.perf_counter() - t0) * 1e3
    tm = h.timings()
    out.update(quad=quad, logdet=logdet, nug=nugs, status=status, chol_ms=tm["chol_ms"], chol_i8_outputs=tm["chol_i8_outputs"])
    if cmpL:
        out["L"] = np.tril(h.get(0, libmogp.GET_L))
        out["alpha"] = h.get(0, libmogp.GET_ALPHA)
    for rep in range(2):
        h.timings(reset=True)
        mean, var, st = h.predict(Xs)
    tm = h.timings()
    out.update(mean=mean, var=var, i8_prep_ms=tm["i8_prep_ms"], trsm_ms=tm["trsm_ms"], i8_fallbacks=tm["i8_fallbacks"])
    h.close()
    return out


for case in (sys.argv[1:] or ["small", "mid", "ragged", "illcond", "c2", "c3x4", "c3", "c4"]):
    n, d, E, m, kern, nt, nug, tc, cmpL = CASES[case]
    a = run(case, "0")
    b = run(case, "1")
    flops = E * (n ** 3) / 3.0
    line = ["%-8s n=%d E=%d" % (case, n, E),
            "chol ms fp64 %.3f (%.1f TF) | i8 %.3f (%.1f TF-equiv) outputs_on_i8=%d" % (a["chol_ms"], flops / a["chol_ms"] / 1e9, b["chol_ms"], flops / b["chol_ms"] / 1e9, b["chol_i8_outputs"]),
            "status %s/%s nug equal %s" % (a["status"].tolist()[:4], b["status"].tolist()[:4], bool(np.array_equal(a["nug"], b["nug"]))),
            "logdet max abs diff %.3e (|logdet| %.3e)" % (np.max(np.abs(a["logdet"] - b["logdet"])), np.max(np.abs(a["logdet"]))),
            "quad max rel diff %.3e" % np.max(np.abs(a["quad"] - b["quad"]) / np.abs(a["quad"])),
            "mean max rel %.3e" % (np.max(np.abs(a["mean"] - b["mean"])) / np.max(np.abs(a["mean"]))),
            "var max abs %.3e (nugget %.1e), max rel %.3e" % (np.max(np.abs(a["var"] - b["var"])), nug, np.max(np.abs(a["var"] - b["var"]) / np.maximum(np.abs(a["var"]), 1e-300))),
            "planes-of-L pass ms %.3f -> %.3f, trsm ms %.2f -> %.2f, fallbacks %d/%d" % (a["i8_prep_ms"], b["i8_prep_ms"], a["trsm_ms"], b["trsm_ms"], a["i8_fallbacks"], b["i8_fallbacks"])]
    if cmpL:
        dl = np.abs(a["L"] - b["L"])
        line.append("L max abs diff %.3e, ||dL||_F/||L||_F %.3e, alpha max rel %.3e" % (dl.max(), np.linalg.norm(dl) / np.linalg.norm(a["L"]), np.max(np.abs(a["alpha"] - b["alpha"])) / np.max(np.abs(a["alpha"]))))
    print("\n   ".join(line), flush=True)
