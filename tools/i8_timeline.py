"""Timeline of CTA 0 of i8_trsm_kernel (debug build: ./build.sh -DI8_TRACE): per tile, the stamps of the loader, the MMA
issuer and the consumers.  usage (under gpurun): python tools/i8_timeline.py > gpurun_out/i8_timeline.txt"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_thetas, WORKLOADS
import mogp_emulator_b200 as mogp
from mogp_emulator_b200 import libmogp

E, n, d, m, kernel, nugget, seed = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
X, Y, Xs = make_workload(n, d, E, m, seed)
gp = mogp.MultiOutputGP_GPU(X, Y, kernel=kernel, nugget=nugget)
gp.fit(make_thetas(E, d))
gp.predict(Xs, deriv=False)
gp.predict(Xs, deriv=False)
EV, TILES = 16, 2048
buf = (ctypes.c_ulonglong * (EV * TILES))()
lib = libmogp._lib
lib.mogp_debug_i8_trace.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
assert lib.mogp_debug_i8_trace(buf, EV * TILES) == 0
a = np.array(buf[:], dtype=np.int64).reshape(TILES, EV)
used = a[:, 1] > 0
a = a[used]
t0 = a[0, 1]
names = ["i", "ticket", "flag_ok", "loads_issued", "mma_start", "mma_issued", "cons_start", "acc_full", "drained", "T_ready",
         "dmma_done", "digits_done", "end"]
print("# tile-seq " + " ".join(names) + "   (us since the first ticket of CTA 0; 0 = not applicable)")
for k, r in enumerate(a):
    print(k, int(r[0]), " ".join("%9.2f" % ((v - t0) / 1e3 if v > 0 else 0.0) for v in r[1:13]))
# phase statistics over tiles with i >= 8
big = a[a[:, 0] >= 8]
def dur(x, y):
    v = (big[:, y] - big[:, x]) / 1e3
    return "%.2f" % np.median(v)
print("# medians over tiles with i >= 8 (us): mma_start->mma_issued", dur(4, 5), " acc_full->drained", dur(7, 8), " drained->T_ready", dur(8, 9),
      " T_ready->dmma_done", dur(9, 10), " dmma_done->digits_done", dur(10, 11), " digits_done->end", dur(11, 12),
      " cons_start->acc_full (waiting for the MMAs)", dur(6, 7), " mma_issued->acc_full", dur(5, 7))
nxt = (big[1:, 4] - big[:-1, 5]) / 1e3
print("# median gap mma_issued(k) -> mma_start(k+1) (us):", "%.2f" % np.median(nxt), " median tile period (us):", "%.2f" % np.median(np.diff(big[:, 12]) / 1e3))
