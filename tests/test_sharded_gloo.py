"""N>1 host path on CPU: two ranks (torch.distributed, gloo) each handle their block of outputs, pack the
block exactly as libmogp_b200 does for ncclAllGather, all-gather, unpack -- every rank must end with the full
(E, m) posterior arrays.  The per-rank compute here is the CPU oracle (there is no GPU in this suite); the
partitioning / packing / ordering logic is the product's (mogp_emulator_b200.sharding)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, E, unfit, out_dir, with_deriv=False):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gp_oracle as orc
    from mogp_emulator_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, Y, Xs = orc.make_workload(60, 2, E, 11, seed=1)
    thetas = np.tile(np.array([1.0, 0.5, 0.1]), (E, 1)) + 0.05 * np.arange(E)[:, None]
    lo, hi, e_pad = sharding.shard_bounds(E, rank, world)
    gp = orc.OracleMultiOutputGP(X, Y[lo:hi], nugget=1e-6, priors="weak") if hi > lo else None
    fitted = []
    for k, i in enumerate(range(lo, hi)):
        if i not in unfit:
            gp.fit_emulator(k, thetas[i])
        fitted.append(i not in unfit)
    m, D = Xs.shape
    if hi > lo:
        mean, var = gp.predict(Xs, allow_not_fit=True)
    else:
        mean, var = np.empty((0, m)), np.empty((0, m))           # a rank without outputs still joins the collective
    deriv = None
    if with_deriv:
        deriv = np.full((hi - lo, m, D), np.nan)
        for k in range(hi - lo):
            if fitted[k]:
                deriv[k] = gp.emulators[k].predict_deriv(Xs)
    block = torch.from_numpy(sharding.pack_block(mean, var, fitted, e_pad, deriv=deriv))
    gathered = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(gathered, block)                      # the one collective of the path
    got = sharding.unpack_gathered(torch.stack(gathered).numpy(), E, world, m, d=D if with_deriv else 0)
    extra = {"deriv": got[3]} if with_deriv else {}
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), mean=got[0], var=got[1], status=got[2], **extra)
    dist.destroy_process_group()


@pytest.mark.parametrize("E,unfit,world,with_deriv", [(4, (), 2, False), (5, (3,), 2, False), (5, (1,), 2, True), (2, (), 3, True)])
def test_gather_reassembles_all_outputs(tmp_path, E, unfit, world, with_deriv):
    """Even and uneven (balanced) partitions, an unfit emulator, gathered derivatives, more ranks than outputs."""
    import gp_oracle as orc
    mp.spawn(_worker, args=(world, _free_port(), E, tuple(unfit), str(tmp_path), with_deriv), nprocs=world, join=True)
    X, Y, Xs = orc.make_workload(60, 2, E, 11, seed=1)
    thetas = np.tile(np.array([1.0, 0.5, 0.1]), (E, 1)) + 0.05 * np.arange(E)[:, None]
    full = orc.OracleMultiOutputGP(X, Y, nugget=1e-6, priors="weak")
    for i in range(E):
        if i not in unfit:
            full.fit_emulator(i, thetas[i])
    want_mean, want_var = full.predict(Xs, allow_not_fit=True)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        np.testing.assert_allclose(z["mean"], want_mean, rtol=1e-12, equal_nan=True)
        np.testing.assert_allclose(z["var"], want_var, rtol=1e-12, atol=1e-18, equal_nan=True)
        assert list(z["status"]) == [4 if i in unfit else 0 for i in range(E)]
        if with_deriv:
            for i in range(E):
                if i in unfit:
                    assert np.all(np.isnan(z["deriv"][i]))
                else:
                    np.testing.assert_allclose(z["deriv"][i], full.emulators[i].predict_deriv(Xs), rtol=1e-12, atol=1e-15)
