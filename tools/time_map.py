"""fit_GP_MAP on a multi-output emulator: batched lock-step search vs one emulator at a time (GPU box)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mogp_emulator_b200 import MultiOutputGP_GPU, GaussianProcessGPU, fit_GP_MAP
E, n, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
X, Y, Xs = bench.make_workload(n, d, E, 10, 5)
theta0 = np.zeros(d + 1)
mo = MultiOutputGP_GPU(X, Y, nugget=1e-6)
mo.priors
fit_GP_MAP(mo, n_tries=1, theta0=theta0, maxiter=20)       # warm-up
mo.reset_fit_status()
mo.timings(reset=True)
t0 = time.perf_counter(); fit_GP_MAP(mo, n_tries=1, theta0=theta0, maxiter=20); t1 = time.perf_counter()
print("batched: %d emulators n=%d d=%d: %.3f s, stats %s" % (E, n, d, t1 - t0, mo.map_fit_stats))
tm = mo.timings()
print("  device ms: fit %.1f (kmat %.1f chol %.1f solve %.1f) grad %.1f" % (tm["fit_ms"], tm["kmat_ms"], tm["chol_ms"], tm["solve_ms"], tm["grad_ms"]))
tb = mo.thetas[0].get_data().copy()
mo.close()
gp = GaussianProcessGPU(X, Y[0], nugget=1e-6); gp.priors
fit_GP_MAP(gp, n_tries=1, theta0=theta0, maxiter=20)
t0 = time.perf_counter()
for i in range(min(E, 4)):
    g = GaussianProcessGPU(X, Y[i], nugget=1e-6)
    g._priors = gp._priors
    fit_GP_MAP(g, n_tries=1, theta0=theta0, maxiter=20)
    if i == 0: ts = g.theta.get_data().copy()
    g.close()
t1 = time.perf_counter()
print("one at a time: %.3f s per emulator (x%d = %.3f s); theta[0] equal: %s" % ((t1 - t0) / min(E, 4), E, (t1 - t0) / min(E, 4) * E, np.allclose(tb, ts, rtol=1e-9)))
