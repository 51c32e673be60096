import os, sys, ctypes
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200 import _lib
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate, METHODS, _ptr, _stream
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
T, P, L, U = (torch.as_tensor(a).cuda() for a in lhs_conditions(n, seed=7))
s = Surrogate(ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eon"))
ref = s.sweep(T, P, L, U, method="bs23", rtol=3e-7, atol=1e-12, staged=True)
plan = s._sweep_plan(n)
for flags in (0, 1):
    for rep in range(6):
        y = torch.full((9, n), -7.0, dtype=torch.float64, device="cuda")
        status = torch.full((n,), 77, dtype=torch.int32, device="cuda")
        stats = torch.zeros((3, n), dtype=torch.int32, device="cuda")
        idx = torch.full((n,), -5, dtype=torch.int32, device="cuda")
        tend = torch.full((n,), -1.0, dtype=torch.float32, device="cuda")
        stiff = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(_lib.lib().pfr_sweep_run(plan, _ptr(T), _ptr(P), _ptr(L), _ptr(U), n, METHODS["bs23"], 64, 3e-7, 1e-12, 0, flags,
                                            _ptr(y), _ptr(status), _ptr(stats), _ptr(idx), _ptr(tend), _ptr(stiff), _stream()), "run")
        torch.cuda.synchronize()
        untouched = int((status == 77).sum())
        bad = (y != ref.y).any(0)
        dmax = float((y - ref.y).abs().max())
        print(f"flags {flags} rep {rep}: untouched {untouched}, status!=0 {int((status != 0).sum())}, columns differing {int(bad.sum())}, max |dy| {dmax:.3e}, "
              f"idx differs {int((idx != ref.idx_cut).sum())}, tend differs {int((tend != ref.t_end).sum())}, stats differ {int((stats != ref.stats).any(0).sum())}")
        if int(bad.sum()):
            j = int(bad.nonzero()[0])
            print("   first bad column", j, "y", y[:, j].tolist()[:3], "ref", ref.y[:, j].tolist()[:3], "idx", int(idx[j]), int(ref.idx_cut[j]), "stats", stats[:, j].tolist(), ref.stats[:, j].tolist())
