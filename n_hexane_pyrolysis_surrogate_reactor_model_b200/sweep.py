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


def gather_outlets(y_local, n_total: int, group=None):
    """All-gather the per-rank [9, n_r] outlet blocks into [9, n_total] on every rank (NCCL on GPUs, gloo on CPU).
    Blocks may be ragged by one column, so they are padded to a common width for the collective."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return y_local
    ws = dist.get_world_size(group)
    width = (n_total + ws - 1) // ws
    pad = torch.zeros((y_local.shape[0], width), dtype=y_local.dtype, device=y_local.device)
    pad[:, : y_local.shape[1]] = y_local
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad, group=group)
    parts = []
    for r in range(ws):
        lo, hi = shard_bounds(n_total, ws, r)
        parts.append(out[r][:, : hi - lo])
    return torch.cat(parts, dim=1)
