"""A/B of the tensor-core operand formats in ONE process, alternating: one time-MLP pass (t_end only / full grid) and the whole
Eon / Eoff sweep at 2^20 LHS conditions with mlp_mode = tf32x3 | f16x3; also the raw-output difference between the two and
against the FP32-FFMA path on 4096 conditions."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
modes = ("tf32x3", "f16x3")
T, P, L, U = [torch.as_tensor(np.asarray(v, np.float32)).cuda() for v in lhs_conditions(n, seed=1)]
ms_on = ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eon")
ms_off = ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eoff")

# numerics: raw grids of the three arithmetic modes on 4096 conditions
k = 4096
raw = {}
for mode in ("fp32",) + modes:
    s = Surrogate(ms_on, mlp_mode=mode)
    raw[mode] = (s.time_grid(T[:k], P[:k], L[:k], U[:k], want_grid=True, raw=True)[0].double().cpu(), s.temp_profile(T[:k], P[:k], raw=True).double().cpu())
    del s
for mode in modes:
    print(json.dumps({"mode": mode, "time_grid_maxdiff_vs_fp32": float((raw[mode][0] - raw["fp32"][0]).abs().max()),
                      "temp_profile_maxdiff_vs_fp32": float((raw[mode][1] - raw["fp32"][1]).abs().max())}), flush=True)
torch.cuda.empty_cache()

def timeit(fn, reps=5):
    for _ in range(3):
        keep = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        keep = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, keep

grid = torch.empty((801, n), dtype=torch.float32, device="cuda")
for rep in range(2):
    for mode in modes:
        s = Surrogate(ms_on, mlp_mode=mode)
        a, _ = timeit(lambda: s.time_grid(T, P, L, U, want_grid=False, want_end=True))
        b, _ = timeit(lambda: s.time_grid(T, P, None, None, want_grid=True, out=grid))
        c, r = timeit(lambda: s.sweep(T, P, L, U, method="bs23", rtol=3e-7, atol=1e-12))
        rec = {"mode": mode, "t_end_only_ms": a, "full_grid_ms": b, "eon_sweep_ms": c, "eon_integrator_ms": s.integrator_ms(),
               "eon_checksum": float(r.y.sum()), "failed": int((r.status != 0).sum())}
        del s, r
        torch.cuda.empty_cache()
        s = Surrogate(ms_off, mlp_mode=mode)
        d, r = timeit(lambda: s.sweep(T, P, L, U, method="dp54", rtol=1e-7, atol=1e-7))
        rec.update({"eoff_sweep_ms": d, "eoff_checksum": float(r.y.sum())})
        print(json.dumps(rec), flush=True)
        del s, r
        torch.cuda.empty_cache()
