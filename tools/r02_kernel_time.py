"""Time the Eon integrator alone (LLNL, 2^20 LHS conditions) at given rtol:atol pairs and report the error triple against RODAS4 at
1e-11 on every 16th condition.  The library under test is chosen with CRNN_PFR_LIB.  Usage: python tools/r02_kernel_time.py tag pairs [prec]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet  # noqa: E402
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate  # noqa: E402
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions  # noqa: E402
from r02_explore import errs, timed  # noqa: E402


def main():
    tag = sys.argv[1]
    pairs = tuple(tuple(float(v) for v in pr.split(":")) for pr in sys.argv[2].split(","))
    prec = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    n = 1 << 20
    gold = os.path.join(ROOT, "tests", "golden", "containers", "LLNL.npz")
    T, P, L, U = (torch.as_tensor(a).cuda() for a in lhs_conditions(n, seed=13895))
    s = Surrogate(ModelSet.from_packed(gold, "Eon"))
    c0 = s.inlet_concentration(T, P)
    _, tend = s.time_grid(T, P, L, U, want_grid=False, want_end=True)
    tfull, _ = s.time_grid(T, P)
    Tp = s.temp_profile(T, P)
    idx = s.idx_cut(tfull, tend)
    perm = torch.argsort(idx, descending=True).to(torch.int32)
    sel = torch.arange(0, n, 16, device="cuda")
    sub = dict(tgrid=tfull[:, sel].contiguous(), Tprof=Tp[:, sel].contiguous(), idx_end=idx[sel].contiguous())
    ref = s.integrate(T[sel], c0[sel], method="rodas4", rtol=1e-11, atol=1e-11, **sub).y.clone()
    for rtol, atol in pairs:
        ms, res = timed(lambda: s.integrate(T, c0, tgrid=tfull, Tprof=Tp, idx_end=idx, perm=perm, method="bs23", precision=prec, rtol=rtol, atol=atol), reps=4)
        y = s.integrate(T[sel], c0[sel], method="bs23", precision=prec, rtol=rtol, atol=atol, **sub).y.double()
        st = res.stats.double()
        print(json.dumps(dict(tag=tag, precision=prec, rtol=rtol, atol=atol, ms=ms, attempts=float((st[0] + st[1]).mean()),
                              failed=int((res.status != 0).sum()), checksum=float(res.y.double().sum()), **errs(y, ref))), flush=True)


if __name__ == "__main__":
    main()
