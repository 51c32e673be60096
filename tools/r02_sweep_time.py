"""Whole Eon / Eoff sweep at 2^20 LHS conditions: one-call pipeline vs staged path, ms per sweep and the integrator's share."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T, P, L, U = (torch.as_tensor(a).cuda() for a in lhs_conditions(n, seed=13895))
for variant in ("Eon", "Eoff"):
    s = Surrogate(ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), variant))
    for staged in (True, False, True, False):
        keep = None
        for _ in range(3):
            keep = s.sweep(T, P, L, U, method="fast", staged=staged)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            r = s.sweep(T, P, L, U, method="fast", staged=staged)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps(dict(variant=variant, staged=staged, ms=e0.elapsed_time(e1) / 5, integrator_ms=None if staged else s.integrator_ms(),
                              failed=int((r.status != 0).sum()), stiff=r.stiff_fallbacks, checksum=float(r.y.sum()))), flush=True)
    del s
    torch.cuda.empty_cache()
