"""Where the host-buffer sweep spends its wall time (diagnostic): python tools/diag_e2e.py [n]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "containers")
T, P, L, U = [np.ascontiguousarray(a, np.float32) for a in lhs_conditions(n)]
dev = [torch.from_numpy(a).cuda() for a in (T, P, L, U)]
for mode in ("tf32x3", "fp32", "tf32x3"):
    sur = Surrogate(ModelSet.from_packed(os.path.join(G, "LLNL.npz"), "Eon"), mlp_mode=mode)
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sur.sweep(*dev); torch.cuda.synchronize(); t1 = time.perf_counter()
        y, st, r = sur.sweep_host(T, P, L, U); t2 = time.perf_counter()
        # phases of sweep_host by hand
        a = time.perf_counter()
        for j, x in enumerate((T, P, L, U)):
            sur._pin_in[j].copy_(torch.from_numpy(x))
        b = time.perf_counter()
        d = [sur._pin_in[j].to(sur.device, non_blocking=True) for j in range(4)]
        torch.cuda.synchronize(); c = time.perf_counter()
        res = sur.sweep(*d); torch.cuda.synchronize(); e = time.perf_counter()
        sur._pin_y.copy_(res.y, non_blocking=True); sur._pin_st.copy_(res.status, non_blocking=True)
        torch.cuda.synchronize(); f = time.perf_counter()
        print(f"{mode} it{it}: device sweep {1e3*(t1-t0):.1f} ms | sweep_host {1e3*(t2-t1):.1f} ms | stage {1e3*(b-a):.1f} h2d {1e3*(c-b):.1f} "
              f"sweep {1e3*(e-c):.1f} d2h {1e3*(f-e):.1f}", flush=True)
    del sur
