"""Phases of the diagonal (D) tiles of chol_dataflow_kernel (debug build: ./build.sh -DCHOL_TRACE) for one n = 4096 matrix:
load R_jj -> factor -> write L_jj -> invert 16x16 diagonal sub-blocks -> off-diagonal inverse blocks -> write inv(L_jj) + publish.
usage (under gpurun): python tools/chol_dtile_timeline.py > gpurun_out/chol_dtile.txt"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, make_thetas
import mogp_emulator_b200 as mogp
from mogp_emulator_b200 import libmogp

n, d = 4096, 10
X, Y, Xs = make_workload(n, d, 1, 16, 1)
gp = mogp.GaussianProcessGPU(X, Y[0], nugget=1e-6)
theta = make_thetas(1, d)[0]
lib = libmogp._lib
lib.mogp_debug_chol_trace.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int, ctypes.c_int]
gp.fit(theta)
gp.fit(theta)
lib.mogp_debug_chol_trace(None, 0, 1)
gp.fit(theta)
EV, N = 8, 4096
buf = (ctypes.c_ulonglong * (EV * N))()
cnt = lib.mogp_debug_chol_trace(buf, EV * N, 0)
a = np.array(buf[:], dtype=np.int64).reshape(N, EV)[:cnt]
a = a[np.argsort(a[:, 7])]
names = ["load", "factor", "write_L+logdet", "inv_diag16", "inv_offdiag", "write_Dinv+publish"]
print("# j start_us " + " ".join(names) + " total   (us)")
for r in a:
    t = r[:7]
    print(int(r[7]), "%.1f" % ((t[0] - a[0, 0]) / 1e3), " ".join("%.1f" % ((t[k + 1] - t[k]) / 1e3) for k in range(6)), "%.1f" % ((t[6] - t[0]) / 1e3))
d_ = np.diff(a[:, :7], axis=1) / 1e3
print("# medians (us):", " ".join("%s %.1f" % (nm, np.median(d_[:, k])) for k, nm in enumerate(names)), " total %.1f" % np.median((a[:, 6] - a[:, 0]) / 1e3),
      " column period %.1f" % np.median(np.diff(a[:, 0]) / 1e3))
lib.mogp_debug_chol_trace2.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
buf2 = (ctypes.c_ulonglong * (32 * cnt))()
lib.mogp_debug_chol_trace2(buf2, 32 * cnt)
b = np.array(buf2[:], dtype=np.uint64).astype(np.int64).reshape(cnt, 32)
# stamps of potf2_inv_block (shared memory, flushed after the tile): 0 start, 1 a1(0) done, then per panel s = 0..6:
# 2+3s chain warp done (a2n + b1n + a1(s+1)), 3+3s a2r + block row s of the inverse done (warp 1), 4+3s trailing update done (warp 1)
med = lambda x: float(np.median(x)) / 1e3
print("# a1(0): %.2f us" % med(b[:, 1] - b[:, 0]))
start = b[:, 1]
for s_ in range(7):
    ch, up, iv = b[:, 2 + 3 * s_], b[:, 3 + 3 * s_], b[:, 4 + 3 * s_]
    print("# panel %d (us from the barrier): chain warp %.2f | a2r + inverse row %.2f, + trailing update %.2f" % (s_, med(ch - start), med(up - start), med(iv - start)))
    start = np.maximum(ch, iv)
print(gp._handle.timings())
