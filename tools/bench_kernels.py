"""Kernel-level timing on one GPU (development tool, not the contract bench): MLP stages, integrator variants,
pipe peaks.  Usage: python tools/bench_kernels.py [n] ; CRNN_PFR_LIB selects a tuning build of the library."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet  # noqa: E402
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate, measure_peaks  # noqa: E402
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    what = sys.argv[2] if len(sys.argv) > 2 else "all"
    gold = os.path.join(ROOT, "tests", "golden", "containers", "LLNL.npz")
    T, P, L, U = (torch.as_tensor(a).cuda() for a in lhs_conditions(n, seed=13895))
    rep = {"n": n, "lib": os.environ.get("CRNN_PFR_LIB", "default")}
    if what in ("all", "peaks"):
        rep["peaks"] = measure_peaks()
    for variant in ("Eoff", "Eon"):
        s = Surrogate(ModelSet.from_packed(gold, variant))
        c0 = s.inlet_concentration(T, P)
        if what in ("all", "mlp"):
            ms, _ = timed(lambda: s.time_grid(T, P, L, U, want_grid=False, want_end=True))
            rep[f"{variant}/time_mlp_end_ms"] = ms
            rep[f"{variant}/time_mlp_tflops"] = n * 935936 * 2 / ms / 1e9
        _, tend = s.time_grid(T, P, L, U, want_grid=False, want_end=True)
        if variant == "Eoff":
            for srt in (False, True):
                perm = torch.argsort(T, descending=True).to(torch.int32) if srt else None
                for prec in (64, 32):
                    for meth in ("rodas4", "rodas4_tpc"):
                        ms, res = timed(lambda: s.integrate(T, c0, t_end=tend, perm=perm, precision=prec, method=meth))
                        st = res.stats.double()
                        rep[f"Eoff/{meth}_f{prec}_sort{int(srt)}_ms"] = ms
                        rep[f"Eoff/{meth}_f{prec}_steps"] = [float(st[0].mean()), float(st[1].mean()), float(st[2].mean()), float(st[0].max())]
                        rep[f"Eoff/{meth}_f{prec}_bad"] = int((res.status != 0).sum())
            perm = torch.argsort(T, descending=True).to(torch.int32)
            ms, res = timed(lambda: s.integrate(T, c0, t_end=tend, perm=perm, method="dopri5", precision=32))
            rep["Eoff/dopri5_f32_ms"] = ms
            rep["Eoff/dopri5_rhs_mean"] = float(res.stats[2].double().mean())
        else:
            tfull, _ = s.time_grid(T, P)
            Tp = s.temp_profile(T, P)
            idx = s.idx_cut(tfull, tend)
            if what in ("all", "mlp"):
                ms, _ = timed(lambda: s.temp_profile(T, P))
                rep["Eon/temp_mlp_ms"] = ms
                ms, _ = timed(lambda: s.time_grid(T, P))
                rep["Eon/time_mlp_full_ms"] = ms
            for srt in (False, True):
                perm = torch.argsort(idx, descending=True).to(torch.int32) if srt else None
                for prec in (64, 32):
                    for meth in ("rodas4", "rodas4_tpc"):
                        ms, res = timed(lambda: s.integrate(T, c0, tgrid=tfull, Tprof=Tp, idx_end=idx, perm=perm, precision=prec, method=meth), reps=2)
                        st = res.stats.double()
                        rep[f"Eon/{meth}_f{prec}_sort{int(srt)}_ms"] = ms
                        rep[f"Eon/{meth}_f{prec}_steps"] = [float(st[0].mean()), float(st[1].mean()), float(st[2].mean()), float(st[0].max())]
                        rep[f"Eon/{meth}_f{prec}_bad"] = int((res.status != 0).sum())
                        if meth == "rodas4" and prec == 64 and srt:
                            ref = s.integrate(T, c0, tgrid=tfull, Tprof=Tp, idx_end=idx, perm=perm, precision=64, method="rodas4_tpc")
                            rep["Eon/coop_vs_tpc_maxrel"] = float(((res.y - ref.y).abs() / ref.y.abs().clamp_min(1e-3)).max())
            rep["Eon/idx_cut_mean"] = float(idx.double().mean())
            ms, res = timed(lambda: s.sweep(T, P, L, U), reps=2)
            rep["Eon/sweep_ms"] = ms
        if variant == "Eoff":
            ms, res = timed(lambda: s.sweep(T, P, L, U))
            rep["Eoff/sweep_ms"] = ms
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
