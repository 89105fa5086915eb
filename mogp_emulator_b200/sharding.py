"""Block partition of the outputs of a multi-output emulator over ranks: rank r owns
``[r*ceil(E/R), min(E, (r+1)*ceil(E/R)))``; every rank's gather block is padded to ``ceil(E/R)`` rows so the
all-gather counts are equal.  Pure host logic, shared by the NCCL path and the gloo CPU tests."""


def shard_bounds(n_outputs, rank, world):
    """-> (lo, hi, e_pad)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    e_pad = -(-int(n_outputs) // int(world))
    lo = min(n_outputs, rank * e_pad)
    hi = min(n_outputs, (rank + 1) * e_pad)
    return lo, hi, e_pad


def gathered_rows(n_outputs, world):
    """Row indices, inside the gathered ``(world*e_pad, m)`` block, of outputs 0..n_outputs-1 in order."""
    rows = []
    for r in range(world):
        lo, hi, e_pad = shard_bounds(n_outputs, r, world)
        rows.extend(r * e_pad + k for k in range(hi - lo))
    return rows
