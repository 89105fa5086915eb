#!/bin/bash
# compute-sanitizer passes over a small fit + predict + gradient + full-covariance + mean-function run (GPU box)
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import numpy as np, gp_oracle as orc
from mogp_emulator_b200 import MultiOutputGP_GPU, GaussianProcessGPU
X, Y, Xs = orc.make_workload(300, 4, 3, 150, seed=0)
thetas = np.array([[1.0, 0.8, 1.2, 0.9, 0.0], [0.5, 1.0, 1.0, 1.0, 0.3], [1.0, 1.0, 1.0, 1.0, -0.2]])
mo = MultiOutputGP_GPU(X, Y, nugget=1e-6)
mo.fit(thetas)
r = mo.predict(Xs)
g = mo.logpost_and_deriv_batch([0, 2], thetas[[0, 2]])
mo.close()
gp = GaussianProcessGPU(X, Y[0] + 1.0, mean="1", kernel="Matern52", nugget="adaptive")
gp.fit(thetas[0])
c = gp.predict(Xs, full_cov=True)
gp.logpost_deriv(thetas[0])
gp.close()
# many right-hand sides: the persistent int8 tcgen05 TRSM (csrc/trsm_i8.cu) with its a-posteriori FP64 check on the side stream
X, Y, Xs = orc.make_workload(300, 3, 40, 600, seed=1)
mo = MultiOutputGP_GPU(X, Y, nugget=1e-6)
mo.fit(np.tile(np.array([1.0, 0.9, 1.1, 0.0]), (40, 1)))
r8 = mo.predict(Xs, deriv=False)
assert mo.timings()["i8_block_rows"] == 3, mo.timings()
mo.close()
# the factorisation with its history products on the int8 tensor cores (csrc/chol.cu chol_i8_kernel), forced at a small size;
# the predict then reads the planes of L the factorisation left
os.environ["MOGP_CHOL_I8"] = "1"
X, Y, Xs = orc.make_workload(500, 3, 40, 600, seed=3)
mo = MultiOutputGP_GPU(X, Y, nugget=1e-6)
mo.fit(np.tile(np.array([1.0, 0.9, 1.1, 0.0]), (40, 1)))
assert mo.timings()["chol_i8_outputs"] == 40, mo.timings()
rc8 = mo.predict(Xs, deriv=False)
assert mo.timings()["i8_block_rows"] == 4, mo.timings()
mo.close()
print("case done", float(r.mean[0, 0]), float(c.unc[0, 0]), float(r8.unc[0, 0]), float(rc8.unc[0, 0]))
PY
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 200 python /tmp/san_case.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|case done" gpurun_out/sanitizer_$tool.log | tail -3
done
