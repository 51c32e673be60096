"""CPU ORACLE (test infrastructure, NOT product code): the reference's sweep drivers restated with the
reference's own structure, for timing the CPU path (bench.py cpu_baseline / --impl reference) and for
end-to-end checks.  One Python process, torch CPU ops, one condition at a time -- exactly how
SURROGATE_MODEL/surrogate_model_Eoff_single_model.py:339-369 and ...Eon_single_model.py:296-368 work.
"""
from __future__ import annotations

import numpy as np
import torch

from . import reference_path as R


def _mlp(ms_mlp):
    return R.MLPParams(ms_mlp.w, ms_mlp.b, ms_mlp.out_min, ms_mlp.out_max)


def eoff_sweep(ms, T, P, L, U, rtol=1e-6, atol=1e-6, keep_going=False):
    """...Eoff_single_model.py main(): ONE batched time-MLP call, enforce_strict per row, then the serial
    per-condition predict_n_ode loop.  Returns outlets [n, 9] float32 (state at the last knot).
    keep_going: torchdiffeq aborts the reference SCRIPT at its first assert (`underflow in dt`, non-finite state; the
    call at ...Eoff_single_model.py:185 has no handler); with keep_going that one condition gets a NaN row instead and
    the loop carries on, so that a timing sample is not lost to a single condition."""
    c0 = R.inlet_concentration(T, P)
    tg = R.time_grid(_mlp(ms.time_mlp), T, P, L, U)
    out = np.empty((len(T), 9), np.float32)
    for i in range(len(T)):
        try:
            sol = R.crnn_predict(tg[i], np.full(R.NTOTAL, T[i], np.float32), c0[i], ms.crnn.w_in, ms.crnn.w_b, ms.crnn.w_out,
                                 rtol=rtol, atol=atol)
            out[i] = sol[:, -1]
        except AssertionError:
            if not keep_going:
                raise
            out[i] = np.nan
    return out


def eon_sweep(ms, T, P, L, U, rtol=1e-6, atol=1e-6, keep_going=False):
    """...Eon_single_model.py main(): per condition a batch-1 temperature MLP, a batch-1 full-length time MLP,
    the full 801-point integration, a batch-1 short time MLP, idx_cut and the trim.  Returns [n, 9] float32.
    keep_going: as in eoff_sweep (the reference's call at ...Eon_single_model.py:154-155 would abort the script)."""
    tm, pm = _mlp(ms.time_mlp), _mlp(ms.temp_mlp)
    c0 = R.inlet_concentration(T, P)
    out = np.empty((len(T), 9), np.float32)
    for i in range(len(T)):
        Ti, Pi = T[i:i + 1], P[i:i + 1]
        Tp = R.temp_profile(pm, Ti, Pi)[0]
        t_full = R.time_grid(tm, Ti, Pi, np.full(1, R.FULL_L, np.float32), np.full(1, R.FULL_U0, np.float32))[0]
        try:
            sol = R.crnn_predict(t_full, Tp, c0[i], ms.crnn.w_in, ms.crnn.w_b, ms.crnn.w_out, rtol=rtol, atol=atol)
        except AssertionError:
            if not keep_going:
                raise
            out[i] = np.nan
            continue
        t_short = R.time_grid(tm, Ti, Pi, L[i:i + 1], U[i:i + 1])[0]
        k = R.eon_idx_cut(t_full, t_short[-1])
        out[i] = sol[:, k]
    return out


def sweep(ms, T, P, L, U, **kw):
    return (eon_sweep if ms.energy_on else eoff_sweep)(ms, T, P, L, U, **kw)


def _worker(args):
    torch.set_num_threads(1)
    ms, T, P, L, U = args
    return sweep(ms, T, P, L, U, keep_going=True)   # a failing condition costs ITS row only (NaN), not the worker's chunk


def sweep_parallel(ms, T, P, L, U, processes: int):
    """Best-case CPU scaling of the reference path: the per-condition loop spread over `processes` workers."""
    import multiprocessing as mp

    chunks = [c for c in np.array_split(np.arange(len(T)), processes) if len(c)]
    with mp.get_context("fork").Pool(len(chunks)) as pool:
        parts = pool.map(_worker, [(ms, T[c], P[c], L[c], U[c]) for c in chunks])
    return np.concatenate(parts)
