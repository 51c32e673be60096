"""Reference-format outputs of a sweep: prediction dumps and the per-case / per-species accuracy table.

  np.savetxt(pred_*.txt, fmt="%.6e") with columns [t, T, P, L, u0, H2, CH4, C2H4, C2H6, C3H6, C4H8-1, NC6H14]
      SURROGATE_MODEL/surrogate_model_Eon_single_model.py:338-368, ...Eoff_single_model.py:345-372
  RMSE / NRMSE / relative error (final and residence-time average), Frechet-type distance, max-norm per case and species
      ...Eoff_single_model.py:384-480, ...Eon_single_model.py:381-463
The file writers run on the host; the accuracy numbers of a whole batch come from one kernel (accuracy_device /
accuracy_rows_device, C entry point pfr_accuracy) that reads the [801, 9, n] dense trajectories where the integrator
left them.  accuracy_rows is the single-condition host form of the same formulas.
"""
from __future__ import annotations

import os

import numpy as np

SPECIES_OBS = ["H2", "CH4", "C2H4", "C2H6", "C3H6", "C4H8-1", "NC6H14"]
COLUMNS = ["Case_ID", "Species_ID", "T_ini [K]", "P_ini [Pa]", "L_ini [m]", "u0_ini [m/s]", "RMSE_final", "NRMSE_final",
           "RelError_final(%)", "RMSE_time_avg", "NRMSE_time_avg", "RelError_time_avg(%)", "FCD", "Max_Norm"]
EPS_REL = 1.0e-5


def prediction_table(t, Tprof, P, L, u0, species) -> np.ndarray:
    """[n_t, 12] float array of one condition: species is [>=7, n_t]; products' t = 0 value is forced to 0 like
    `pred_species[:-1, 0] = 0.0` (the n-hexane row keeps its inlet value)."""
    sp = np.array(species[:7], dtype=np.float32, copy=True)
    sp[:-1, 0] = 0.0
    t = np.asarray(t, np.float32)
    cols = [t, np.broadcast_to(np.asarray(Tprof, np.float32), t.shape), np.full_like(t, P), np.full_like(t, L), np.full_like(t, u0)]
    return np.vstack(cols + list(sp)).T


def write_prediction_files(out_dir, prefix, tgrid, Tprof, dense, T, P, L, u0, idx_cut=None, fmt="%.6e"):
    """One `prefix{idx}.txt` per condition (1-based idx).  tgrid/Tprof [801, n] (Tprof None: isothermal), dense [801, 9, n]
    host arrays; idx_cut [n] trims the Eon trajectories at `[:idx_cut + 1]`."""
    os.makedirs(out_dir, exist_ok=True)
    n = len(T)
    paths = []
    for i in range(n):
        k = NTOT if idx_cut is None else int(idx_cut[i]) + 1
        Tp = np.full(k, T[i], np.float32) if Tprof is None else Tprof[:k, i]
        tab = prediction_table(tgrid[:k, i], Tp, P[i], L[i], u0[i], dense[:k, :, i].T)
        path = os.path.join(out_dir, f"{prefix}{i + 1}.txt")
        np.savetxt(path, tab, fmt=fmt)
        paths.append(path)
    return paths


NTOT = 801


def accuracy_rows(case_id, pred, true, T, P, L, u0, absolute_denominator=False):
    """Rows of the reference's accuracy CSV for one condition.  pred/true [7, n_t] including the t = 0 column, which is
    excluded like `true = true[1:]`.  absolute_denominator: the Eon script divides by |ref| + eps, the Eoff one by ref + eps."""
    pred = np.asarray(pred, np.float32)[:, 1:]
    true = np.asarray(true, np.float32)[:, 1:]
    den = (np.abs(true) if absolute_denominator else true) + EPS_REL
    span = true.max(axis=1) - true.min(axis=1) + EPS_REL
    rmse_final = np.sqrt((pred[:, -1] - true[:, -1]) ** 2)
    rel_final = np.abs(pred[:, -1] - true[:, -1]) / den[:, -1] * 100
    rmse_time = np.sqrt(np.mean((pred - true) ** 2, axis=1))
    rel_time = np.mean(np.abs(pred - true) / den, axis=1) * 100
    fcd = np.sqrt((true.mean(axis=1) - pred.mean(axis=1)) ** 2 + (true.std(axis=1) - pred.std(axis=1)) ** 2)
    max_norm = np.max(np.abs(pred - true), axis=1) / (np.max(np.abs(true), axis=1) + EPS_REL)
    return [[case_id, SPECIES_OBS[s], T, P, L, u0, rmse_final[s], rmse_final[s] / span[s], rel_final[s], rmse_time[s],
             rmse_time[s] / span[s], rel_time[s], fcd[s], max_norm[s]] for s in range(7)]


METRICS = COLUMNS[6:]


def accuracy_device(dense, labels, idx_cut=None, absolute_denominator=False):
    """[8, 7, n] float64 CUDA tensor of the eight CSV numbers for every condition and observed species.
    dense [801, 9, n] float64/float32 CUDA (SolveResult.dense), labels [801, 7, n] float32 CUDA at the same knots,
    idx_cut [n] int32 CUDA for the Eon trim (knots 1..idx_cut) or None."""
    import torch

    from . import _lib
    from .surrogate import _ptr, _stream
    n = dense.shape[2]
    if tuple(dense.shape[:2]) != (NTOT, 9) or tuple(labels.shape) != (NTOT, 7, n) or labels.dtype != torch.float32:
        raise ValueError("dense must be [801, 9, n], labels [801, 7, n] float32")
    dense, labels = dense.contiguous(), labels.contiguous()
    idx = None if idx_cut is None else idx_cut.to(torch.int32).contiguous()
    out = torch.empty((8, 7, n), dtype=torch.float64, device=dense.device)
    _lib.check(_lib.lib().pfr_accuracy(_ptr(dense), 64 if dense.dtype == torch.float64 else 32, _ptr(labels), _ptr(idx), n,
                                       int(bool(absolute_denominator)), _ptr(out), _stream()), "pfr_accuracy")
    return out


def accuracy_rows_device(dense, labels, T, P, L, u0, idx_cut=None, absolute_denominator=False):
    """The reference CSV rows (case-major, species inside, Case_ID 1-based) for a whole batch from accuracy_device."""
    m = accuracy_device(dense, labels, idx_cut, absolute_denominator).cpu().numpy()
    rows = []
    for i in range(m.shape[2]):
        for s in range(7):
            rows.append([i + 1, SPECIES_OBS[s], T[i], P[i], L[i], u0[i]] + [m[k, s, i] for k in range(8)])
    return rows


def nearest_time_labels(t_pred, t_label, y_label):
    """Eon label matching: for every predicted time the label at the nearest label time (...Eon_single_model.py:409-417)."""
    idx = np.abs(np.asarray(t_label)[None, :] - np.asarray(t_pred)[:, None]).argmin(axis=1)
    return np.asarray(y_label)[:, idx]


def accuracy_table(rows):
    import pandas as pd
    return pd.DataFrame(rows, columns=COLUMNS)
