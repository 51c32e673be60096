"""Multi-rank plumbing on CPU: contiguous sharding of the condition batch and the final gather over gloo
(world_size 2), which is the only collective on the sweep path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import gather_outlets, lhs_conditions, shard_bounds


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 400, 1 << 20, (1 << 20) + 3):
        for ws in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, ws, r) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_lhs_conditions_are_a_latin_hypercube():
    n = 1000
    T, P, L, U = lhs_conditions(n, seed=13895)
    assert T.dtype == np.float32 and T.shape == (n,)
    assert 870 <= T.min() and T.max() <= 1150 and 1e5 <= P.min() and P.max() <= 3e5
    assert 0.5 <= L.min() and L.max() <= 1.0 and 2.5 <= U.min() and U.max() <= 5.0
    strata = np.floor((T.astype(np.float64) - 870) / 280 * n).astype(int)
    assert len(np.unique(np.clip(strata, 0, n - 1))) >= n - 2          # one sample per stratum (float32 rounding aside)
    T2, _, _, _ = lhs_conditions(n, seed=13895)
    assert np.array_equal(T, T2)                                       # seeded, every rank draws the same batch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_total, world, rank)
    full = torch.arange(9 * n_total, dtype=torch.float64).reshape(9, n_total)
    got = gather_outlets(full[:, lo:hi].contiguous(), n_total)
    blocks = gather_outlets(full[:, lo:hi].contiguous(), n_total, as_blocks=True).blocks()    # views of the persistent buffer, no copy
    views_ok = all(torch.equal(b, full[:, slice(*shard_bounds(n_total, world, r))]) for r, b in enumerate(blocks))
    again = gather_outlets(2 * full[:, lo:hi].contiguous(), n_total)                          # the buffer is reused by the next sweep
    ret[rank] = bool(torch.equal(got, full)) and views_ok and bool(torch.equal(again, 2 * full))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [10, 11])
def test_gather_outlets_gloo_world2(n_total):
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), n_total, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def test_gather_is_identity_without_process_group():
    y = torch.ones(9, 5)
    assert gather_outlets(y, 5) is y
