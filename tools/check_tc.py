"""Tensor-core MLP path vs FP32 path vs torch CPU (development check)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions
from oracle import reference_path as R

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
ms = ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eon")
T, P, L, U = lhs_conditions(n, seed=1)
a = Surrogate(ms, mlp_mode="fp32")
b = Surrogate(ms, mlp_mode="tf32x3")
ga, _ = a.time_grid(T, P, L, U, raw=True)
print("fp32 done", flush=True)
gb, _ = b.time_grid(T, P, L, U, raw=True)
torch.cuda.synchronize()
print("tc done", flush=True)
ga, gb = ga.cpu().numpy()[1:].T, gb.cpu().numpy()[1:].T
mp = R.MLPParams(ms.time_mlp.w, ms.time_mlp.b, ms.time_mlp.out_min, ms.time_mlp.out_max)
x = R.scale_inputs([T, P, L, U], 4)
m = min(n, 2000)
r32 = R.mlp_forward(mp, x[:m]); r64 = R.mlp_forward(mp, x[:m].astype(np.float64), dtype=torch.float64)
print("max|fp32 kernel - torch32|", np.abs(ga[:m] - r32).max(), " vs f64", np.abs(ga[:m] - r64).max())
print("max|tc   kernel - torch32|", np.abs(gb[:m] - r32).max(), " vs f64", np.abs(gb[:m] - r64).max())
print("max|torch32 - f64|", np.abs(r32 - r64).max(), " max|tc - fp32 kernel|", np.abs(ga - gb).max())
pa = a.temp_profile(T, P).cpu().numpy(); pb = b.temp_profile(T, P).cpu().numpy()
print("temp profile max diff [K]", np.abs(pa - pb).max())
if n >= 65536:
    for s, nm in ((a, "fp32"), (b, "tf32x3")):
        Td, Pd, Ld, Ud = (torch.as_tensor(v).cuda() for v in (T, P, L, U))
        s.time_grid(Td, Pd, Ld, Ud, want_grid=False, want_end=True); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): s.time_grid(Td, Pd, Ld, Ud, want_grid=False, want_end=True)
        e1.record(); torch.cuda.synchronize()
        ms_ = e0.elapsed_time(e1) / 3
        print(nm, "time MLP pass ms", ms_, "TFLOP/s", n * 935936 * 2 / ms_ / 1e9)
