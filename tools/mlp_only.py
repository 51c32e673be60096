"""One time-MLP pass (end-only mode) on n conditions, for profiling the MLP kernels alone: python tools/mlp_only.py [n] [reps]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
s = Surrogate(ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eoff"), chunk=chunk,
              mlp_mode=os.environ.get("PFR_AB_MODE", "f16x3"))
T, P, L, U = [torch.as_tensor(np.asarray(v, np.float32)).cuda() for v in lhs_conditions(n, seed=1)]
s.time_grid(T, P, L, U, want_grid=False, want_end=True); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): s.time_grid(T, P, L, U, want_grid=False, want_end=True)
e1.record(); torch.cuda.synchronize()
print(f"chunk {chunk}: time MLP pass, {n} conditions: {e0.elapsed_time(e1) / reps:.3f} ms")
