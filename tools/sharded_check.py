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
mo.close()

# derivatives and a mean function are gathered too (the reference returns deriv (E, m, D) by default,
# MultiOutputGP_GPU.py:185-297); again against a single-GPU emulator on the same inputs
Ym = Y + 1.5 - 0.7 * X[:, 0]
mm = MultiOutputGP_GPU(X, Ym, mean="x[0]", nugget=1e-6, device=local_rank, comm=comm)
mm.fit(thetas)
rm = mm.predict(Xs)                       # deriv=True, the default
om = MultiOutputGP_GPU(X, Ym, mean="x[0]", nugget=1e-6, device=local_rank)
om.fit(thetas)
r1m = om.predict(Xs)
assert rm.deriv.shape == (E, 77, 3)
assert np.array_equal(rm.mean, r1m.mean) and np.array_equal(rm.unc, r1m.unc) and np.array_equal(rm.deriv, r1m.deriv), "sharded (mean function, deriv) != single"
assert mm.get_indices_not_fit() == [] and mm.get_indices_fit() == list(range(E))      # status exchange at the end of fit
mm.close()
om.close()

# more ranks than outputs: the ranks without outputs still join the collective
if world > 1:
    Es = world - 1
    ms = MultiOutputGP_GPU(X, Y[:Es], nugget=1e-6, device=local_rank, comm=comm)
    ms.fit(thetas[:Es])
    rs = ms.predict(Xs, deriv=False)
    assert rs.mean.shape == (Es, 77) and np.array_equal(rs.mean, r1.mean[:Es]) and np.array_equal(rs.unc, r1.unc[:Es]), "empty rank"
    ms.close()
one.close()
print("rank %d/%d: sharded predict OK (local outputs %s)" % (rank, world, mo.local_range))
