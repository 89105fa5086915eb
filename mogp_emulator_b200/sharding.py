"""Balanced block partition of the outputs of a multi-output emulator over ranks: with
``q, rem = divmod(E, R)`` the first ``rem`` ranks own ``q + 1`` consecutive outputs and the others ``q`` (so no rank
is empty while ``R <= E``; with more ranks than outputs the trailing ranks own none and still join the gather with an
all-padding block).  Every rank's gather block is padded to ``e_pad = ceil(E/R)`` rows so the all-gather counts are
equal.  Pure host logic, shared by the NCCL path and the gloo CPU tests."""


def shard_bounds(n_outputs, rank, world):
    """-> (lo, hi, e_pad)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    n_outputs, rank, world = int(n_outputs), int(rank), int(world)
    q, rem = divmod(n_outputs, world)
    e_pad = q + (1 if rem else 0)
    lo = rank * q + min(rank, rem)
    hi = lo + q + (1 if rank < rem else 0)
    return lo, hi, e_pad


def gathered_rows(n_outputs, world):
    """Row indices, inside the gathered ``(world*e_pad, m)`` block, of outputs 0..n_outputs-1 in order."""
    rows = []
    for r in range(world):
        lo, hi, e_pad = shard_bounds(n_outputs, r, world)
        rows.extend(r * e_pad + k for k in range(hi - lo))
    return rows


def pack_block(mean, var, fitted, e_pad, deriv=None):
    """This rank's gather block: ``e_pad`` rows of ``[mean (m) | var (m) | deriv (m*D)]`` (NaN rows for padding / unfit
    outputs) followed by ``e_pad`` status words (0 = ok, 4 = not fit, 2 = padding).  Without derivatives this is the
    ``[e_pad][2][m]`` layout libmogp_b200 packs on the device for ncclAllGather (csrc/api.cu: mogp_predict_allgather); with
    them (or with a mean function) ``MultiOutputGP_GPU.predict`` packs on the host and gathers through mogp_comm_allgather.
    A rank without outputs passes empty (0, m) arrays."""
    import numpy as np
    mean = np.asarray(mean, dtype=np.float64)
    e_loc, m = mean.shape
    width = 2 * m + (0 if deriv is None else int(np.prod(np.shape(deriv)[1:])))
    block = np.full(e_pad * width + e_pad, np.nan)
    res = block[:e_pad * width].reshape(e_pad, width)
    status = block[e_pad * width:]
    status[:] = 2.0
    for k in range(e_loc):
        if fitted[k]:
            res[k, :m] = mean[k]
            res[k, m:2 * m] = var[k]
            if deriv is not None:
                res[k, 2 * m:] = np.reshape(deriv[k], -1)
            status[k] = 0.0
        else:
            status[k] = 4.0
    return block


def unpack_gathered(gathered, n_outputs, world, m, d=0):
    """(world, e_pad*width + e_pad) gathered blocks -> mean (E, m), var (E, m), status (E,) in output order, and with
    ``d > 0`` the derivatives (E, m, d) as a fourth value."""
    import numpy as np
    e_pad = shard_bounds(n_outputs, 0, world)[2]
    width = (2 + d) * m
    gathered = np.asarray(gathered).reshape(world, e_pad * width + e_pad)
    mean = np.empty((n_outputs, m))
    var = np.empty((n_outputs, m))
    deriv = np.empty((n_outputs, m, d)) if d else None
    status = np.empty(n_outputs, dtype=np.int32)
    for r in range(world):
        lo, hi, _ = shard_bounds(n_outputs, r, world)
        res = gathered[r, :e_pad * width].reshape(e_pad, width)
        st = gathered[r, e_pad * width:]
        for k in range(hi - lo):
            mean[lo + k] = res[k, :m]
            var[lo + k] = res[k, m:2 * m]
            if d:
                deriv[lo + k] = res[k, 2 * m:].reshape(m, d)
            status[lo + k] = int(st[k])
    return (mean, var, status, deriv) if d else (mean, var, status)
