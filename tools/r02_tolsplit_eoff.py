"""Which (rtol, atol) pair certifies max outlet error <= 1e-6 on the isothermal (Eoff) path at the least cost?  DP54 free stepping
to t_end, LHS conditions, error vs RODAS4 at 1e-11 on every 16th condition.  Usage: python tools/r02_tolsplit_eoff.py [n] [mech] [pairs]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet  # noqa: E402
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate  # noqa: E402
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions  # noqa: E402
from r02_explore import timed  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    mech = sys.argv[2] if len(sys.argv) > 2 else "LLNL"
    gold = os.path.join(ROOT, "tests", "golden", "containers", f"{mech}.npz")
    T, P, L, U = (torch.as_tensor(a).cuda() for a in lhs_conditions(n, seed=13895))
    s = Surrogate(ModelSet.from_packed(gold, "Eoff"))
    c0 = s.inlet_concentration(T, P)
    _, tend = s.time_grid(T, P, L, U, want_grid=False, want_end=True)
    sel = torch.arange(0, n, 16, device="cuda")
    ref = s.integrate(T[sel], c0[sel], t_end=tend[sel].contiguous(), method="rodas4", rtol=1e-11, atol=1e-11).y.clone()
    ref2 = s.integrate(T[sel], c0[sel], t_end=tend[sel].contiguous(), method="rodas4", rtol=1e-12, atol=1e-12).y
    e = ((ref2 - ref).abs() / torch.clamp(ref.abs(), min=1e-3)).amax(0)
    print(json.dumps(dict(mech=mech, exp="reference_self_consistency_1e-11_vs_1e-12", err_max=float(e.max()), err_median=float(e.median()))), flush=True)
    pairs = ((1e-7, 1e-7), (1e-7, 1e-9), (1e-7, 1e-11), (3e-8, 1e-10), (1e-8, 1e-8), (1e-8, 1e-10), (1e-8, 1e-12), (3e-9, 1e-11), (1e-9, 1e-9), (1e-9, 1e-11))
    if len(sys.argv) > 3:
        pairs = tuple(tuple(float(v) for v in pr.split(":")) for pr in sys.argv[3].split(","))
    for rtol, atol in pairs:
        ms, res = timed(lambda: s.integrate(T, c0, t_end=tend, method="dp54", rtol=rtol, atol=atol))
        y = s.integrate(T[sel], c0[sel], t_end=tend[sel].contiguous(), method="dp54", rtol=rtol, atol=atol).y
        ee = (y - ref).abs() / torch.clamp(ref.abs(), min=1e-3)
        e = ee.amax(0)
        worst = int(e.argmax())
        sp = int(ee[:, worst].argmax())
        st = res.stats.double()
        print(json.dumps(dict(mech=mech, exp="rtol_atol_eoff", method="dp54", rtol=rtol, atol=atol, ms=ms, evaluations=float(st[2].mean()),
                              attempts=float((st[0] + st[1]).mean()), failed=int((res.status != 0).sum()), err_max=float(e.max()),
                              err_p99=float(torch.quantile(e, 0.99)), err_median=float(e.median()), n_over_1e6=int((e > 1e-6).sum()),
                              worst_species=sp, worst_ref=float(ref[sp, worst]), worst_abs=float((y - ref)[sp, worst]))), flush=True)


if __name__ == "__main__":
    main()
