"""Latin-hypercube condition batches and their sharding over the GPUs of one box.

The reference samples its operating conditions with scipy's LatinHypercube
(INDEPENDENT_DATASET_GENERATION/Latin_hypercube_sampling_4D.py:12-37: d=4, seed 13895, bounds
T in [870,1150] K, P in [1,3] bar, L in [0.5,1] m, u0 in [2.5,5] m/s) and then loops over the rows serially
(SURROGATE_MODEL/surrogate_model_Eoff_single_model.py:339).  Every condition is an independent initial
value problem, so the batch shards into contiguous blocks, one per rank, with no collective on the data
path; the only communication is the final gather of the [9, n] outlets.
"""
from __future__ import annotations

import numpy as np

T_BOUNDS, P_BOUNDS_BAR, L_BOUNDS, U_BOUNDS = (870.0, 1150.0), (1.0, 3.0), (0.5, 1.0), (2.5, 5.0)


def lhs_conditions(n: int, seed: int = 13895):
    """(T[K], P[Pa], L[m], u0[m/s]) float32.  scipy LatinHypercube(d=4, seed) without the O(n^2) "random-cd"
    optimisation the reference uses for its 400-point sets."""
    from scipy.stats import qmc

    u = qmc.LatinHypercube(d=4, seed=seed).random(n)
    lo = np.array([T_BOUNDS[0], P_BOUNDS_BAR[0], L_BOUNDS[0], U_BOUNDS[0]])
    hi = np.array([T_BOUNDS[1], P_BOUNDS_BAR[1], L_BOUNDS[1], U_BOUNDS[1]])
    x = qmc.scale(u, lo, hi)
    return (x[:, 0].astype(np.float32), (x[:, 1] * 1.0e5).astype(np.float32), x[:, 2].astype(np.float32),
            x[:, 3].astype(np.float32))


def shard_bounds(n: int, world_size: int, rank: int):
    """Contiguous block [lo, hi) of rank `rank`; blocks differ in size by at most one condition."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GatheredOutlets:
    """The outlets of every rank after the final gather: one [world, 9, width] buffer (width = ceil(n_total / world)) as the
    collective filled it.  Rank r's conditions are buf[r, :, :hi_r - lo_r]; `blocks()` hands out those views, `full()`
    concatenates them into [9, n_total] for callers that want one array (a copy -- the sweep itself never needs it)."""

    def __init__(self, buf, n_total: int):
        self.buf, self.n_total = buf, n_total

    def blocks(self):
        ws = self.buf.shape[0]
        return [self.buf[r, :, : hi - lo] for r, (lo, hi) in ((r, shard_bounds(self.n_total, ws, r)) for r in range(ws))]

    def full(self):
        import torch
        return torch.cat(self.blocks(), dim=1)


_gather_buffers = {}


def gather_outlets(y_local, n_total: int, group=None, as_blocks: bool = False):
    """The only collective of the sweep: all-gather the per-rank [9, n_r] outlet blocks (NCCL on GPUs, gloo on CPU).
    One `all_gather_into_tensor` into a buffer [world, 9, width] that is allocated once per (shape, dtype, device) and reused by
    every sweep -- no per-call zero fill, tensor list or concatenation.  Blocks may be ragged by one column: a rank whose block
    is narrower than `width` sends it through a persistent staging buffer (the extra column is never read back).
    Returns [9, n_total] (`as_blocks=False`; single process: `y_local` itself) or the GatheredOutlets view object."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return GatheredOutlets(y_local.unsqueeze(0), n_total) if as_blocks else y_local
    ws = dist.get_world_size(group)
    width = (n_total + ws - 1) // ws
    rows = y_local.shape[0]
    key = (rows, width, ws, y_local.dtype, y_local.device)
    bufs = _gather_buffers.get(key)
    if bufs is None:
        bufs = _gather_buffers[key] = (torch.empty((ws, rows, width), dtype=y_local.dtype, device=y_local.device),
                                       torch.zeros((rows, width), dtype=y_local.dtype, device=y_local.device))
    out, stage = bufs
    send = y_local
    if y_local.shape[1] != width or not y_local.is_contiguous():
        stage[:, : y_local.shape[1]].copy_(y_local)
        send = stage
    dist.all_gather_into_tensor(out.view(-1), send.reshape(-1), group=group)   # flat views: rank r fills out[r]
    g = GatheredOutlets(out, n_total)
    return g if as_blocks else g.full()
