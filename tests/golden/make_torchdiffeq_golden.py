"""Emit golden vectors of the REAL torchdiffeq for row a8: run where `torchdiffeq` is installed.

    pip install torchdiffeq          # any 0.2.x; the reference does not pin a version
    python tests/golden/make_torchdiffeq_golden.py

Needs only torch, numpy, torchdiffeq and two committed fixtures of this repository (tests/golden/reference_vectors.npz:
the first 16 conditions of INDEPENDENT_DATASET_GENERATION/sampling_case_4D.csv with their MLP grids and temperature
profiles; tests/golden/containers/LLNL.npz: the trained LLNL parameter containers).  It does NOT import the oracle or the
product.  It integrates the reference's CRNN right-hand side exactly the way the reference calls the library
(SURROGATE_MODEL/surrogate_model_Eoff_single_model.py:175-186, ...Eon_single_model.py:153-156):

    sol = odeint(func, u0, t_ar, method='dopri5', atol=1e-6, rtol=1e-6);  clamp(sol.T, 1e-6, 60)

in float32 (the reference's dtype) and float64, and writes tests/golden/torchdiffeq_vectors.npz with the [16, 9, 801]
trajectories and the right-hand-side call counts.  tests/test_torchdiffeq_pin.py compares oracle/reference_path.py's
restatement of dopri5 with that file (and with torchdiffeq itself when it is importable): this is the pin that turns
"parity unpinned" for a8 into a checked statement.  This container has no torchdiffeq and no network, so the file is not
committed yet.
"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
NS, NTOTAL = 9, 801
LB, UB, ZLO, ZHI, DULO, DUHI = 1.0e-6, 6.0e1, -3.0e1, 3.0e1, -1.0e5, 1.0e5   # ...Eon_single_model.py:57-62
R_KCAL = 1.9872036e-3                                                       # ...Eon_single_model.py:63


def linear_interpolation(ts, ys):
    """...Eoff_single_model.py:106-115."""
    def interp(t):
        idx = torch.clamp(torch.searchsorted(ts, t.reshape(1), right=True), 1, ts.numel() - 1)
        t0, t1, y0, y1 = ts[idx - 1], ts[idx], ys[idx - 1], ys[idx]
        return (y0 + (y1 - y0) / (t1 - t0) * (t - t0)).squeeze()
    return interp


class CRNNFunc(torch.nn.Module):
    """...Eon_single_model.py:133-151 (the Eoff script's class at :117-155 computes the same du)."""

    def __init__(self, t_ar, T_ar, w_in, w_b, w_out):
        super().__init__()
        self.T_fn = linear_interpolation(t_ar, T_ar)
        self.w_in, self.w_b, self.w_out = w_in, w_b, w_out
        self.calls = 0

    def forward(self, t, u):
        self.calls += 1
        T = self.T_fn(t)
        Y = torch.clamp(u, LB, UB)
        w_v = torch.cat([torch.log(Y), (-1.0 / (R_KCAL * T)).reshape(1), torch.log(T).reshape(1)]).to(u.dtype)
        w_in_x = torch.clamp(self.w_in.T @ w_v + self.w_b, ZLO, ZHI)
        return torch.clamp(self.w_out @ torch.exp(w_in_x), DULO, DUHI)


def main():
    from torchdiffeq import odeint
    import torchdiffeq

    torch.set_num_threads(1)
    vec = np.load(os.path.join(HERE, "reference_vectors.npz"))
    con = np.load(os.path.join(HERE, "containers", "LLNL.npz"))
    out = {"torchdiffeq_version": np.array(getattr(torchdiffeq, "__version__", "unknown")), "torch_version": np.array(torch.__version__)}
    for variant, crnn in (("Eoff", "Eoff_wide_v2"), ("Eon", "Eon")):
        keys = [k for k in con.files if k.startswith(f"crnn/{crnn}/")]
        if not keys:
            raise SystemExit(f"containers/LLNL.npz holds no crnn/{crnn}/*: {con.files}")
        w_in, w_b, w_out = (con[f"crnn/{crnn}/{k}"] for k in ("w_in", "w_b", "w_out"))
        tgrid = vec["Eoff/tgrid"] if variant == "Eoff" else vec["Eon/tgrid_full"]
        Tprof = np.repeat(vec["T"][:, None], NTOTAL, 1) if variant == "Eoff" else vec["Eon/Tprof"]
        for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
            sols, calls = [], []
            for i in range(tgrid.shape[0]):
                t = torch.tensor(tgrid[i], dtype=dt)
                f = CRNNFunc(t, torch.tensor(Tprof[i], dtype=dt), *(torch.tensor(a, dtype=dt) for a in (w_in, w_b, w_out)))
                with torch.no_grad():
                    sol = odeint(f, torch.tensor(vec["c0"][i], dtype=dt), t, method="dopri5", atol=1e-6, rtol=1e-6)
                sols.append(torch.clamp(sol.T, LB, UB).numpy())
                calls.append(f.calls)
            out[f"{variant}/dopri5_{name}"] = np.stack(sols)
            out[f"{variant}/nfe_{name}"] = np.array(calls, np.int32)
            print(variant, name, "nfe", calls)
    path = os.path.join(HERE, "torchdiffeq_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
