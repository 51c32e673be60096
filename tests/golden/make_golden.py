"""Generates the committed fixtures under tests/golden/ from the reference checkout at /root/reference.

Run here (the reference is not present on the GPU box):   python tests/golden/make_golden.py

  containers/{LLNL,JetSurf,NUIG}.npz  the reference's trained parameter containers, repacked one file per
                                      mechanism (same float32 payload; n_hexane_..._b200.containers.pack_mechanism)
  conditions.npz                      the four header-less sampling_case_*.csv files as float64 arrays
  converter_kat.npz                   updated_p -> final_parameters known-answer pairs stored by the reference in
                                      training_history_LLNL_Eoff_wide.npz and training_history_NUIG_Eon.npz
  reference_vectors.npz               outputs of oracle/reference_path.py (torch-op restatement of the reference,
                                      torch CPU float32) on the first 16 conditions of sampling_case_4D.csv:
                                      MLP grids, temperature profiles, idx_cut, inlet concentrations, RHS values,
                                      dopri5(1e-6,1e-6) trajectories, and converged float64 outlets
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200 import containers as C  # noqa: E402
from oracle import c_oracle as CO  # noqa: E402
from oracle import reference_path as R  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
NGOLD = 16


def main():
    torch.set_num_threads(1)
    for mech in C.MECHANISMS:
        C.pack_mechanism(REF, mech, os.path.join(OUT, "containers", f"{mech}.npz"))
    cond = {
        "independent_4D": np.loadtxt(f"{REF}/INDEPENDENT_DATASET_GENERATION/sampling_case_4D.csv", delimiter=","),
        "independent_2D": np.loadtxt(f"{REF}/INDEPENDENT_DATASET_GENERATION/sampling_case_2D.csv", delimiter=","),
        "training_2D": np.loadtxt(f"{REF}/CRNN_TEMP_PRED_MODEL_TRAINING_DATASET_GENERATION/sampling_case_2D.csv", delimiter=","),
        "training_wide_2D": np.loadtxt(f"{REF}/CRNN_TEMP_PRED_MODEL_TRAINING_DATASET_GENERATION/sampling_case_wide_2D.csv", delimiter=","),
    }
    np.savez_compressed(os.path.join(OUT, "conditions.npz"), **cond)

    kat = {}
    for name in ("LLNL_Eoff_wide", "NUIG_Eon"):
        d = np.load(f"{REF}/SURROGATE_MODEL_PARAMETER_CONTAINER/training_history_{name}.npz", allow_pickle=True)
        fp = d["final_parameters"].item()
        kat[f"{name}/updated_p"] = d["updated_p"]
        for k in ("w_in", "w_b", "w_out"):
            kat[f"{name}/{k}"] = fp[k]
            assert np.array_equal(fp[k], d["parameters"][-1][k])
    np.savez_compressed(os.path.join(OUT, "converter_kat.npz"), **kat)

    a = cond["independent_4D"][:NGOLD]
    T, P = a[:, 0].astype(np.float32), (a[:, 1] * 1e5).astype(np.float32)
    L, U = a[:, 2].astype(np.float32), a[:, 3].astype(np.float32)
    c0 = R.inlet_concentration(T, P)
    vec = {"T": T, "P": P, "L": L, "U": U, "c0": c0}
    for variant in ("Eoff", "Eon"):
        ms = C.ModelSet.from_reference_dir(REF, "LLNL", variant)
        cr = ms.crnn
        tm = R.MLPParams(ms.time_mlp.w, ms.time_mlp.b, ms.time_mlp.out_min, ms.time_mlp.out_max)
        x4 = R.scale_inputs([T, P, L, U], 4)
        vec[f"{variant}/x_scaled"] = x4
        vec[f"{variant}/time_mlp_raw"] = R.mlp_forward(tm, x4)
        tshort = R.time_grid(tm, T, P, L, U)
        vec[f"{variant}/tgrid"] = tshort
        if variant == "Eoff":
            tg, Tp = tshort, np.repeat(T[:, None], 801, 1)
            rep = np.full(NGOLD, 800, np.int32)
        else:
            tg = R.time_grid(tm, T, P, np.full_like(T, R.FULL_L), np.full_like(T, R.FULL_U0))
            pm = R.MLPParams(ms.temp_mlp.w, ms.temp_mlp.b, ms.temp_mlp.out_min, ms.temp_mlp.out_max)
            Tp = R.temp_profile(pm, T, P)
            rep = np.array([R.eon_idx_cut(tg[i], tshort[i, -1]) for i in range(NGOLD)], np.int32)
            vec["Eon/tgrid_full"], vec["Eon/Tprof"], vec["Eon/idx_cut"] = tg, Tp, rep
            vec["Eon/temp_mlp_raw"] = R.mlp_forward(pm, R.scale_inputs([T, P], 2))
        vec[f"{variant}/rhs0_f64"] = R.crnn_rhs_np(Tp[:, 0].astype(np.float64), c0.astype(np.float64), cr.w_in, cr.w_b, cr.w_out)
        sols, stats = [], []
        for i in range(NGOLD):
            st = R.SolveStats()
            sols.append(R.crnn_predict(tg[i], Tp[i], c0[i], cr.w_in, cr.w_b, cr.w_out, stats=st))
            stats.append([st.nfe, st.accepted, st.rejected])
        vec[f"{variant}/dopri5_f32"] = np.stack(sols)          # [16, 9, 801]
        vec[f"{variant}/dopri5_stats"] = np.array(stats, np.int32)
        yt, yk = CO.truth_batch(tg, Tp, c0, cr.w_in, cr.w_b, cr.w_out, upto=rep, knots=True, nthreads=8)
        vec[f"{variant}/truth_outlet"] = yt
        vec[f"{variant}/truth_knots_every50"] = yk[:, ::50, :]
        # scipy DOP853 cross-check of the C truth integrator on two conditions
        for i in (0, 5):
            tr = R.converged_trajectory(tg[i], Tp[i], c0[i], cr.w_in, cr.w_b, cr.w_out, upto=int(rep[i]))
            assert np.max(np.abs(tr[-1] - yt[i]) / np.maximum(np.abs(yt[i]), 1e-3)) < 1e-9
    np.savez_compressed(os.path.join(OUT, "reference_vectors.npz"), **vec)
    for f in sorted(os.listdir(OUT)) + sorted(os.listdir(os.path.join(OUT, "containers"))):
        pth = os.path.join(OUT, f) if os.path.exists(os.path.join(OUT, f)) else os.path.join(OUT, "containers", f)
        if os.path.isfile(pth):
            print(f, os.path.getsize(pth))


if __name__ == "__main__":
    main()
