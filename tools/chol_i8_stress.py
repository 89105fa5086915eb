"""Stress of the persistent tcgen05 Cholesky (chol_i8_kernel): many back-to-back factorisations at shapes that load the
dataflow differently (many small matrices, a few, one large, ragged sizes, E > 64 groups), each repeated; every repetition
must reproduce the first one BIT FOR BIT (exact integer products, fixed FP64 order inside a tile: the result may not
depend on which CTA ran which tile when) and no launch may trap (a dependency that is never satisfied traps after 20 s).
usage (under gpurun): python tools/chol_i8_stress.py > gpurun_out/chol_i8_stress.txt"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_thetas
from mogp_emulator_b200 import libmogp

os.environ["MOGP_CHOL_I8"] = "1"
CASES = [(4096, 10, 32, 40), (4096, 10, 4, 60), (1000, 5, 70, 30), (300, 3, 130, 30), (2500, 6, 9, 30), (8192, 8, 3, 10), (16384, 12, 1, 6)]
for n, d, E, reps in CASES:
    X, Y, Xs = make_workload(n, d, E, 8, 3)
    thetas = make_thetas(E, d)
    h = libmogp.Handle(X, Y, 0, 2, 1e-6)
    t0 = time.perf_counter()
    first = None
    bad = 0
    for r in range(reps):
        quad, logdet, nug, status = h.fit(0, thetas)
        L = h.get(E - 1, libmogp.GET_L)[-257:, :]            # the last rows of the last output's factor
        key = (quad.tobytes(), logdet.tobytes(), L.tobytes())
        if first is None:
            first = key
        elif key != first:
            bad += 1
        assert status.max() == 0
    tm = h.timings()
    h.close()
    print("n=%5d E=%3d: %d fits in %.2f s, outputs on the tcgen05 path %d, repetitions that differ from the first: %d"
          % (n, E, reps, time.perf_counter() - t0, tm["chol_i8_outputs"], bad), flush=True)
    assert bad == 0
print("stress OK")
