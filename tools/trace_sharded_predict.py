import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from bench import make_workload, make_thetas, WORKLOADS
from mogp_emulator_b200 import MultiOutputGP_GPU
from mogp_emulator_b200.rendezvous import init_comm, env_rank_world
rank, world, local_rank = env_rank_world()
comm = init_comm(local_rank)
E, n, d, m, kernel, nugget, seed = WORKLOADS["c3x4"] if os.environ.get("SMALL") else WORKLOADS["c3"]
if os.environ.get("SMALL"): E = 4 * world
X, Y, Xs = make_workload(n, d, E, m, seed)
thetas = make_thetas(E, d)
gp = MultiOutputGP_GPU(X, Y, kernel=kernel, nugget=nugget, device=local_rank, comm=comm)
for it in range(4):
    comm.allreduce_max(0.0)
    t0 = time.perf_counter(); gp.fit(thetas); t1 = time.perf_counter()
    if it == 3 and rank == 0: os.environ["MOGP_TRACE"] = "1"
    r = gp.predict(Xs, deriv=False); t2 = time.perf_counter()
    if rank == 0: print("iter", it, "fit %.3f ms predict %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), gp.timings(reset=True), flush=True)
