"""One C3-shaped fit + predict through the int8 path (for ncu launch lists)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gp_oracle as orc
import mogp_emulator_b200 as mogp
os.environ.setdefault("MOGP_TRSM_I8", "1")
X, Y, Xs = orc.make_workload(4096, 10, 32, 10000, seed=2)
thetas = np.tile(np.array([1.0] * 10 + [0.0]), (32, 1))
gp = mogp.MultiOutputGP_GPU(X, Y, nugget=1e-6)
gp.fit(thetas)
mean, var, _ = gp.predict(Xs, deriv=False)
print(gp.timings())
