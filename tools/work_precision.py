"""Work-precision table of the Rosenbrock integrators on one GPU: time and outlet error (against the same kernel family at
rtol = atol = 1e-11, whose parity with the converged oracle solution is what tests/test_gpu_parity.py checks) for the
LLNL Eon and Eoff sweeps.  Usage: python tools/work_precision.py [n] [methods comma-separated]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet  # noqa: E402
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate  # noqa: E402
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions  # noqa: E402


def timed(fn, reps=2):
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
    methods = sys.argv[2].split(",") if len(sys.argv) > 2 else ["rodas4", "ros3"]
    variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["Eon", "Eoff"]
    gold = os.path.join(ROOT, "tests", "golden", "containers", "LLNL.npz")
    T, P, L, U = (torch.as_tensor(a).cuda() for a in lhs_conditions(n, seed=13895))
    rows = []
    for variant in variants:
        s = Surrogate(ModelSet.from_packed(gold, variant))
        c0 = s.inlet_concentration(T, P)
        _, tend = s.time_grid(T, P, L, U, want_grid=False, want_end=True)
        if variant == "Eoff":
            perm = torch.argsort(T, descending=True).to(torch.int32)
            run = lambda meth, tol: s.integrate(T, c0, t_end=tend, perm=perm, method=meth, rtol=tol, atol=tol)
        else:
            tfull, _ = s.time_grid(T, P)
            Tp = s.temp_profile(T, P)
            idx = s.idx_cut(tfull, tend)
            perm = torch.argsort(idx, descending=True).to(torch.int32)
            run = lambda meth, tol: s.integrate(T, c0, tgrid=tfull, Tprof=Tp, idx_end=idx, perm=perm, method=meth, rtol=tol, atol=tol)
        ref = run("rodas4", 1e-11).y.clone()
        scale = torch.clamp(ref.abs(), min=1e-3)
        for meth in (m for m in methods if not (variant == "Eon" and m == "dp54") and not (variant == "Eoff" and m == "bs23")):
            for tol in (1e-5, 1e-6, 1e-7, 3e-8, 1e-8, 1e-9):
                ms, res = timed(lambda: run(meth, tol))
                e = ((res.y - ref).abs() / scale).amax(0)
                st = res.stats.double()
                row = dict(variant=variant, method=meth, tol=tol, ms=ms, traj_per_s=n / ms * 1e3, err_max=float(e.max()),
                           err_p99=float(torch.quantile(e[:: max(1, n // 100000)], 0.99)), err_median=float(e.median()),
                           accepted=float(st[0].mean()), rejected=float(st[1].mean()), rhs=float(st[2].mean()),
                           failed=int((res.status != 0).sum()))
                rows.append(row)
                print(json.dumps(row), flush=True)
    return rows


if __name__ == "__main__":
    main()
