"""Cluster width of solve_alpha_kernel against the number of outputs per launch (n = 4096): fit-solve time per launch.
usage (under gpurun): python tools/solve_sweep.py > gpurun_out/solve_sweep.txt"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_thetas
from mogp_emulator_b200 import libmogp

n, d = 4096, 10
for E in (2, 4, 8, 16, 32):
    X, Y, Xs = make_workload(n, d, E, 16, 2)
    thetas = make_thetas(E, d)
    h = libmogp.Handle(X, Y, 0, 2, 1e-6)
    row = []
    for C in ("", "1", "2", "4", "8", "16"):
        if C:
            os.environ["MOGP_SOLVE_CLUSTER"] = C
        else:
            os.environ.pop("MOGP_SOLVE_CLUSTER", None)
        h.fit(0, thetas)
        h.timings(reset=True)
        for _ in range(3):
            h.fit(0, thetas)
        row.append("%s: %.3f" % (C or "default", h.timings()["solve_ms"] / 3))
    h.close()
    print("outputs %2d  solve ms per launch by cluster width  " % E + "  ".join(row), flush=True)
