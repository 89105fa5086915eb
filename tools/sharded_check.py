"""Sharded (one process per GPU) predict with an output count that does not divide by the number of ranks and one
emulator left unfit: every rank must return the same full (E, m) arrays as a single-GPU emulator.  Run under torchrun."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import gp_oracle as orc
from mogp_emulator_b200 import MultiOutputGP_GPU
from mogp_emulator_b200.rendezvous import init_comm, env_rank_world

rank, world, local_rank = env_rank_world()
comm = init_comm(local_rank)
E = 5
X, Y, Xs = orc.make_workload(300, 3, E, 77, seed=61)
thetas = np.zeros((E, 4))
thetas[:, 0] = np.linspace(0.2, 1.0, E)
mo = MultiOutputGP_GPU(X, Y, nugget=1e-6, device=local_rank, comm=comm)
mo.fit(thetas)
r = mo.predict(Xs, deriv=False)
one = MultiOutputGP_GPU(X, Y, nugget=1e-6, device=local_rank)
one.fit(thetas)
r1 = one.predict(Xs, deriv=False)
assert r.mean.shape == (E, 77) and np.array_equal(r.mean, r1.mean) and np.array_equal(r.unc, r1.unc), "sharded != single"
# leave the last output (owned by the last rank) unfit
mo.reset_fit_status()
for i in range(E - 1):
    mo.fit_emulator(i, thetas[i])
try:
    mo.predict(Xs, deriv=False)
    raise SystemExit("expected ValueError for an unfit emulator")
except ValueError:
    pass
r2 = mo.predict(Xs, deriv=False, allow_not_fit=True)
assert np.all(np.isnan(r2.mean[E - 1])) and np.array_equal(r2.mean[:E - 1], r1.mean[:E - 1])
assert mo.get_indices_not_fit() == [E - 1]
print("rank %d/%d: sharded predict OK (local outputs %s)" % (rank, world, mo.local_range))
