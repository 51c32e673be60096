"""Where do the attempts of the knot-limited fast path go?  Cumulative attempts / rejections up to knot k (all conditions stopped at
knot k), LLNL Eon, 65536 LHS conditions: python tools/r02_attempts_profile.py rtol atol"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions
rtol, atol = float(sys.argv[1]), float(sys.argv[2])
n = 65536
T, P, L, U = (torch.as_tensor(a).cuda() for a in lhs_conditions(n, seed=13895))
s = Surrogate(ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eon"))
c0 = s.inlet_concentration(T, P)
tf, _ = s.time_grid(T, P)
Tp = s.temp_profile(T, P)
for k in (1, 2, 3, 5, 10, 20, 50, 100, 200, 400):
    idx = torch.full((n,), k, dtype=torch.int32, device="cuda")
    r = s.integrate(T, c0, tgrid=tf, Tprof=Tp, idx_end=idx, method="bs23", rtol=rtol, atol=atol, stiff_fallback=None)
    st = r.stats.double()
    print(json.dumps(dict(rtol=rtol, atol=atol, knot=k, accepted=float(st[0].mean()), rejected=float(st[1].mean()), t_knot_median=float(tf[k].median()))), flush=True)
