"""The benchmark's step (fit of all outputs, then a many-right-hand-side predict) repeated back to back at several output counts:
every fit must succeed and reproduce the first one bit for bit (log-determinants, quadratic forms), every predict the first
predict.  usage (under gpurun): python tools/fit_predict_stress.py [reps] > gpurun_out/fit_predict_stress.txt"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_thetas
from mogp_emulator_b200 import libmogp

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
for E in (16, 8, 32, 4):
    X, Y, Xs = make_workload(4096, 10, 32, 10000, 2)
    X, Y = X, Y[:E]
    thetas = make_thetas(32, 10)[:E]
    h = libmogp.Handle(X, Y, 0, 2, 1e-6)
    first = None
    bad_fit = bad_status = bad_pred = 0
    t0 = time.perf_counter()
    for r in range(reps):
        quad, logdet, nug, status = h.fit(0, thetas)
        if status.max() != 0:
            bad_status += 1
            print("  rep %d: status %s" % (r, status.tolist()), flush=True)
        mean, var, st = h.predict(Xs)
        key = (quad.tobytes(), logdet.tobytes(), mean.tobytes(), var.tobytes())
        if first is None:
            first = key
        else:
            bad_fit += key[:2] != first[:2]
            bad_pred += key[2:] != first[2:]
    tm = h.timings()
    h.close()
    print("E=%2d: %d x (fit + predict) in %.1f s; failed fits %d, fits that differ %d, predicts that differ %d; tcgen05 chol outputs %d, "
          "TRSM fall-backs %d" % (E, reps, time.perf_counter() - t0, bad_status, bad_fit, bad_pred, tm["chol_i8_outputs"], tm["i8_fallbacks"]), flush=True)
