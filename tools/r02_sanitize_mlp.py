"""One small MLP pass per tensor-core mode (for compute-sanitizer)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
T, P, L, U = [torch.as_tensor(np.asarray(v, np.float32)).cuda() for v in lhs_conditions(n, seed=1)]
for mode in ("f16x3", "tf32x3"):
    s = Surrogate(ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eon"), mlp_mode=mode)
    g, e = s.time_grid(T, P, L, U, want_grid=True, want_end=True)
    p = s.temp_profile(T, P)
    torch.cuda.synchronize()
    print(mode, float(g.sum()), float(e.sum()), float(p.sum()), flush=True)
