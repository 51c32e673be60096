"""Readers for the reference's on-disk parameter containers and condition files.

The surrogate path is a drop-in: it consumes the files the reference ships, unchanged.

  * CRNN training histories  SURROGATE_MODEL_PARAMETER_CONTAINER/*.npz
        keys train_loss, valid_loss, parameters[E] (object array of dict{w_in[11,9], w_b[9], w_out[9,9]} f32);
        consumers take parameters[-1]  (reference: SURROGATE_MODEL/surrogate_model_Eoff_single_model.py:223-230,
        ...Eon_single_model.py:226-230)
  * MLP weights              {TIME,TEMP}_PRED_MODEL_PARAMETER_CONTAINER/mlp_weights_*.pth
        nn.Linear state_dict fc1..fc4 (reference: ...Eoff_single_model.py:285-293, ...Eon_single_model.py:216-223)
  * MLP output scalers       .../min_max_values_mlp_*.pkl  {'min': np.float64, 'max': np.float64}
        (reference: ...Eoff_single_model.py:277-280, ...Eon_single_model.py:234-241)
  * condition files          */sampling_case_{2D,4D,wide_2D}.csv, header-less; T[K], P[bar] (, L[m], u0[m/s])
        (reference: ...Eoff_single_model.py:242,259-262); 2-column files get L=1.0 m, u0=2.5 m/s
        (CRNN_TEMP_PRED_MODEL_TRAINING_DATASET_GENERATION/simul_data_gene_fix_chain_2D.py:39-40)

A "packed" single-file form (one .npz per mechanism, same float32 payload) is also understood so that
the trained parameters can travel as test fixtures (tests/golden/make_golden.py writes them).
"""
from __future__ import annotations

import os
import pickle
from dataclasses import dataclass, field

import numpy as np

NS, NR, NTOTAL = 9, 9, 801
HIDDEN = 512
SPECIES = ["H2", "CH4", "C2H4", "C2H6", "C3H6", "C4H8-1", "NC6H14", "C4H10", "C5H10-1"]
MECHANISMS = ("LLNL", "JetSurf", "NUIG")
FULL_L, FULL_U0 = 1.0, 2.5

# mechanism registry: which CRNN history the reference's inference scripts load for each variant
CRNN_FILES = {
    ("LLNL", "Eoff"): "training_history_LLNL_Eoff_wide_v2.npz",  # ...Eoff_single_model.py:321
    ("LLNL", "Eoff_narrow"): "training_history_LLNL_Eoff.npz",
    ("LLNL", "Eoff_wide"): "training_history_LLNL_Eoff_wide.npz",
    ("LLNL", "Eon"): "training_history_LLNL_Eon.npz",  # ...Eon_single_model.py:226
    ("JetSurf", "Eoff"): "training_history_JetSurf_Eoff.npz",
    ("JetSurf", "Eon"): "training_history_JetSurf_Eon.npz",
    ("NUIG", "Eoff"): "training_history_NUIG_Eoff.npz",
    ("NUIG", "Eon"): "training_history_NUIG_Eon.npz",
}


@dataclass
class CRNNParams:
    """w_in[11,9] (rows 0-8 reaction orders, 9 Ea [kcal/mol], 10 b), w_b[9] = ln A, w_out[9,9]."""
    w_in: np.ndarray
    w_b: np.ndarray
    w_out: np.ndarray

    def __post_init__(self):
        self.w_in = np.ascontiguousarray(self.w_in, dtype=np.float32)
        self.w_b = np.ascontiguousarray(self.w_b, dtype=np.float32)
        self.w_out = np.ascontiguousarray(self.w_out, dtype=np.float32)
        if self.w_in.shape != (NS + 2, NR) or self.w_b.shape != (NR,) or self.w_out.shape != (NS, NR):
            raise ValueError(f"CRNN parameter shapes {self.w_in.shape} {self.w_b.shape} {self.w_out.shape}")


@dataclass
class MLPParams:
    """fc1..fc4 of the 512-wide predictor (nn.Linear layout [out,in], float32) + output min/max."""
    w: list
    b: list
    out_min: float
    out_max: float

    def __post_init__(self):
        self.w = [np.ascontiguousarray(a, dtype=np.float32) for a in self.w]
        self.b = [np.ascontiguousarray(a, dtype=np.float32) for a in self.b]
        dims = [a.shape for a in self.w]
        ok = (len(dims) == 4 and dims[0][0] == HIDDEN and dims[0][1] in (2, 4) and dims[1] == (HIDDEN, HIDDEN)
              and dims[2] == (HIDDEN, HIDDEN) and dims[3] == (NTOTAL - 1, HIDDEN))
        if not ok:
            raise ValueError(f"unexpected MLP layer shapes {dims}")
        self.out_min = float(self.out_min)
        self.out_max = float(self.out_max)

    @property
    def in_dim(self) -> int:
        return self.w[0].shape[1]


def load_npz_parameters(file_path: str, epoch: int = -1) -> CRNNParams:
    """parameters[epoch] of a CRNN training history (same name as the reference helper)."""
    data = np.load(file_path, allow_pickle=True)
    par = data["parameters"][epoch]
    return CRNNParams(par["w_in"], par["w_b"], par["w_out"])


def load_mlp(pth_path: str, pkl_path: str) -> MLPParams:
    import torch

    sd = torch.load(pth_path, map_location="cpu", weights_only=True)
    with open(pkl_path, "rb") as f:
        mm = pickle.load(f)
    return MLPParams([sd[f"fc{i}.weight"].numpy() for i in (1, 2, 3, 4)],
                     [sd[f"fc{i}.bias"].numpy() for i in (1, 2, 3, 4)], mm["min"], mm["max"])


def load_conditions_csv(path: str):
    """(T[K], P[Pa], L[m], u0[m/s]) float32 arrays.  bar -> Pa in float64 before the float32 cast, as
    `torch.tensor(df.iloc[:,1].values*1.0e+5, dtype=float32)` does (...Eoff_single_model.py:259-262)."""
    a = np.loadtxt(path, delimiter=",", dtype=np.float64, ndmin=2)
    if a.shape[1] not in (2, 4):
        raise ValueError(f"{path}: expected 2 or 4 columns, found {a.shape[1]}")
    T = a[:, 0].astype(np.float32)
    P = (a[:, 1] * 1.0e5).astype(np.float32)
    if a.shape[1] == 4:
        L, u0 = a[:, 2].astype(np.float32), a[:, 3].astype(np.float32)
    else:
        L = np.full_like(T, FULL_L)
        u0 = np.full_like(T, FULL_U0)
    return T, P, L, u0


@dataclass
class ModelSet:
    """Everything one surrogate variant needs: CRNN + time MLP (+ temperature MLP for Eon)."""
    mechanism: str
    variant: str  # 'Eon' | 'Eoff'
    crnn: CRNNParams
    time_mlp: MLPParams
    temp_mlp: MLPParams | None = None
    meta: dict = field(default_factory=dict)

    @property
    def energy_on(self) -> bool:
        return self.variant == "Eon"

    @classmethod
    def from_reference_dir(cls, root: str, mechanism: str = "LLNL", variant: str = "Eoff", crnn_key: str | None = None):
        """Load from a checkout of the reference repository (directory names as shipped)."""
        energy = "on" if variant == "Eon" else "off"
        crnn_file = CRNN_FILES[(mechanism, crnn_key or variant)]
        crnn = load_npz_parameters(os.path.join(root, "SURROGATE_MODEL_PARAMETER_CONTAINER", crnn_file))
        tdir = os.path.join(root, "TIME_PRED_MODEL_PARAMETER_CONTAINER")
        time_mlp = load_mlp(os.path.join(tdir, f"mlp_weights_{mechanism}_4D_time_{energy}.pth"),
                            os.path.join(tdir, f"min_max_values_mlp_{mechanism}_4D_time_{energy}.pkl"))
        temp_mlp = None
        if variant == "Eon":
            pdir = os.path.join(root, "TEMP_PRED_MODEL_PARAMETER_CONTAINER")
            temp_mlp = load_mlp(os.path.join(pdir, f"mlp_weights_{mechanism}_2D.pth"),
                                os.path.join(pdir, f"min_max_values_mlp_{mechanism}_2D.pkl"))
        return cls(mechanism, variant, crnn, time_mlp, temp_mlp, {"crnn_file": crnn_file})

    @classmethod
    def from_packed(cls, path: str, variant: str = "Eoff", crnn_key: str | None = None):
        """Load from a packed per-mechanism .npz written by `pack_mechanism`."""
        z = np.load(path, allow_pickle=False)
        mechanism = str(z["mechanism"])
        energy = "on" if variant == "Eon" else "off"
        key = crnn_key or variant
        crnn = CRNNParams(z[f"crnn/{key}/w_in"], z[f"crnn/{key}/w_b"], z[f"crnn/{key}/w_out"])

        def mlp(prefix):
            return MLPParams([z[f"{prefix}/fc{i}.weight"] for i in (1, 2, 3, 4)],
                             [z[f"{prefix}/fc{i}.bias"] for i in (1, 2, 3, 4)],
                             float(z[f"{prefix}/min"]), float(z[f"{prefix}/max"]))

        time_mlp = mlp(f"time_{energy}")
        temp_mlp = mlp("temp") if variant == "Eon" else None
        return cls(mechanism, variant, crnn, time_mlp, temp_mlp, {"packed": os.path.basename(path)})


def pack_mechanism(root: str, mechanism: str, out_path: str) -> None:
    """Repack one mechanism's containers (all CRNN variants, both time MLPs, the temperature MLP) to one .npz."""
    payload = {"mechanism": np.array(mechanism)}
    for (mech, key), fname in CRNN_FILES.items():
        if mech != mechanism:
            continue
        c = load_npz_parameters(os.path.join(root, "SURROGATE_MODEL_PARAMETER_CONTAINER", fname))
        payload[f"crnn/{key}/w_in"], payload[f"crnn/{key}/w_b"], payload[f"crnn/{key}/w_out"] = c.w_in, c.w_b, c.w_out
    tdir = os.path.join(root, "TIME_PRED_MODEL_PARAMETER_CONTAINER")
    pdir = os.path.join(root, "TEMP_PRED_MODEL_PARAMETER_CONTAINER")
    mlps = {
        "time_on": (os.path.join(tdir, f"mlp_weights_{mechanism}_4D_time_on.pth"), os.path.join(tdir, f"min_max_values_mlp_{mechanism}_4D_time_on.pkl")),
        "time_off": (os.path.join(tdir, f"mlp_weights_{mechanism}_4D_time_off.pth"), os.path.join(tdir, f"min_max_values_mlp_{mechanism}_4D_time_off.pkl")),
        "temp": (os.path.join(pdir, f"mlp_weights_{mechanism}_2D.pth"), os.path.join(pdir, f"min_max_values_mlp_{mechanism}_2D.pkl")),
    }
    for prefix, (pth, pkl) in mlps.items():
        m = load_mlp(pth, pkl)
        for i in range(4):
            payload[f"{prefix}/fc{i + 1}.weight"] = m.w[i]
            payload[f"{prefix}/fc{i + 1}.bias"] = m.b[i]
        payload[f"{prefix}/min"] = np.float64(m.out_min)
        payload[f"{prefix}/max"] = np.float64(m.out_max)
    np.savez(out_path, **payload)
