"""Two-mechanism validation drivers (SURROGATE_MODEL/surrogate_model_{Eoff,Eon}_validation_plot.py) without the plotting.

The reference scripts load two mechanisms' surrogates (MODEL1 = LLNL, MODEL2 = NUIG), pick three test conditions by inlet
temperature, run predict_n_ode for both and draw the curves over the Cantera labels (matplotlib, absent here).  This
module does the same selection and the same two batched solves and hands back / writes the numbers a plot would show.

  three_conditions(test_idx, T)                  ...Eoff_validation_plot.py:367-373   n//4-th, n//2-th and second-hottest
  two_model_comparison(sur1, sur2, T, P, L, u0)  ...Eoff_validation_plot.py:360-405, ...Eon_validation_plot.py (full-length
                                                 grid + temperature profile + idx_cut trim)
"""
from __future__ import annotations

import os

import numpy as np

from .report import SPECIES_OBS, prediction_table
from .surrogate import Surrogate


def three_conditions(test_idx, T_ini) -> list[int]:
    """The reference's pick: test indices sorted by inlet temperature, entries n//4, n//2 and -2."""
    order = sorted(list(test_idx), key=lambda i: float(T_ini[i]))
    n = len(order)
    if n < 2:
        raise ValueError("need at least two test conditions")
    return [int(order[n // 4]), int(order[n // 2]), int(order[-2])]


def two_model_comparison(sur1: Surrogate, sur2: Surrogate, T, P, L, u0, test_idx=None, out_dir: str | None = None,
                         names=("MODEL1", "MODEL2")) -> dict:
    """Trajectories of both surrogates at the three selected conditions.

    Returns {"conditions": [i1, i2, i3], names[0]: [table, table, table], names[1]: [...]} where every table is the
    reference's [n_t, 12] prediction layout [t, T, P, L, u0, 7 species] (Eon: trimmed at idx_cut).  With out_dir the
    tables are written as `<name>_cond<i>.txt` ('%.6e') next to a small index CSV."""
    T, P, L, u0 = (np.asarray(x, np.float32) for x in (T, P, L, u0))
    idx = three_conditions(range(len(T)) if test_idx is None else test_idx, T)
    out = {"conditions": idx}
    for name, sur in zip(names, (sur1, sur2)):
        Ts, Ps, Ls, Us = T[idx], P[idx], L[idx], u0[idx]
        if sur.energy_on:
            res = sur.predict_n_ode(Ts, Ps)
            _, tend = sur.time_grid(Ts, Ps, Ls, Us, want_grid=False, want_end=True)
            cut = sur.idx_cut(res.tgrid, tend).cpu().numpy()
        else:
            res = sur.predict_n_ode(Ts, Ps, Ls, Us)
            cut = np.full(len(idx), 800)
        res.raise_on_failure()
        tg = res.tgrid.cpu().numpy()
        Tp = None if res.Tprof is None else res.Tprof.cpu().numpy()
        dense = res.dense.cpu().numpy()
        tables = []
        for j in range(len(idx)):
            k = int(cut[j]) + 1
            Tcol = np.full(k, Ts[j], np.float32) if Tp is None else Tp[:k, j]
            tables.append(prediction_table(tg[:k, j], Tcol, Ps[j], Ls[j], Us[j], dense[:k, :, j].T))
        out[name] = tables
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "comparison_index.csv"), "w") as f:
            f.write("model,condition,T_ini [K],P_ini [Pa],L_ini [m],u0_ini [m/s],n_t," + ",".join(f"{s}_outlet" for s in SPECIES_OBS) + "\n")
            for name in names:
                for j, i in enumerate(idx):
                    tab = out[name][j]
                    np.savetxt(os.path.join(out_dir, f"{name}_cond{i + 1}.txt"), tab, fmt="%.6e")
                    f.write(f"{name},{i + 1},{T[i]:.6e},{P[i]:.6e},{L[i]:.6e},{u0[i]:.6e},{len(tab)}," +
                            ",".join(f"{v:.6e}" for v in tab[-1, 5:]) + "\n")
    return out
