"""Is the predict TRSM slower behind the tcgen05 Cholesky because of the power cap?  The C3 step with an idle gap between fit
and predict, for both Cholesky kernels.  usage (under gpurun): python tools/power_probe.py > gpurun_out/power_probe.txt"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_thetas
import mogp_emulator_b200 as mogp

X, Y, Xs = make_workload(4096, 10, 32, 10000, 2)
thetas = make_thetas(32, 10)
for chol in ("1", "0"):
    os.environ["MOGP_CHOL_I8"] = chol
    gp = mogp.MultiOutputGP_GPU(X, Y, nugget=1e-6)
    for gap in (0.0, 0.0, 0.05, 0.2, 0.0):
        for rep in range(3):
            gp.timings(reset=True)
            gp.fit(thetas)
            time.sleep(gap)
            gp.predict(Xs, deriv=False)
            t = gp.timings()
        print("MOGP_CHOL_I8=%s gap %.2f s: cholesky %.2f ms, K* %.2f ms, TRSM kernel %.2f ms" % (chol, gap, t["chol_ms"], t["kstar_ms"], t["i8_rows_ms"]), flush=True)
    gp.close()
