"""Where does the end-to-end (construct + fit + predict) wall time go?  Run on the GPU box."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mogp_emulator_b200 import MultiOutputGP_GPU, libmogp
E, n, d, m, kernel, nugget, seed = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
mode = sys.argv[2] if len(sys.argv) > 2 else "plain"
X, Y, Xs = bench.make_workload(n, d, E, m, seed)
thetas = bench.make_thetas(E, d)
if mode == "sampler":
    s = bench.ClockSampler(0); s.start(); time.sleep(1.0); print(s.stop())
if mode == "warm":
    gp = MultiOutputGP_GPU(X, Y, kernel=kernel, nugget=nugget)
    for _ in range(3):
        gp.fit(thetas); gp.predict(Xs, unc=True, deriv=False)
    gp.close(); del gp
for it in range(5):
    t0 = time.perf_counter(); gp = MultiOutputGP_GPU(X, Y, kernel=kernel, nugget=nugget)
    t1 = time.perf_counter(); gp.fit(thetas)
    t2 = time.perf_counter(); r = gp.predict(Xs, unc=True, deriv=False)
    t3 = time.perf_counter(); gp.close(); del gp
    t4 = time.perf_counter()
    print("iter %d: construct %.1f ms, fit %.1f ms, predict %.1f ms, destroy %.1f ms" % (it, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3))
