"""SM clock and power under each phase of the sweep, sampled in-process every 20 ms (NVML): is the GEMM phase power-capped?"""
import os, sys, threading, time, json
import numpy as np, torch, pynvml
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
n = 1 << 20
T, P, L, U = (torch.as_tensor(a).cuda() for a in lhs_conditions(n, seed=13895))
s = Surrogate(ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eon"))
c0 = s.inlet_concentration(T, P)
tf, _ = s.time_grid(T, P); Tp = s.temp_profile(T, P)
_, te = s.time_grid(T, P, L, U, want_grid=False, want_end=True)
idx = s.idx_cut(tf, te)
def sample(fn, secs):
    rows, stop = [], threading.Event()
    def run():
        while not stop.wait(0.02):
            rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
    th = threading.Thread(target=run); th.start()
    t0 = time.time(); it = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        fn(); it += 1
        if it % 4 == 0: torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize(); stop.set(); th.join()
    r = np.array(rows[len(rows) // 4:])
    return dict(ms_per_call=e0.elapsed_time(e1) / it, sm_mhz_median=float(np.median(r[:, 0])), sm_mhz_min=float(r[:, 0].min()), power_w_median=float(np.median(r[:, 1])), power_w_max=float(r[:, 1].max()))
print(json.dumps(dict(phase="mlp_pass_t_end_only", **sample(lambda: s.time_grid(T, P, L, U, want_grid=False, want_end=True), 4.0))))
print(json.dumps(dict(phase="bs23", **sample(lambda: s.integrate(T, c0, tgrid=tf, Tprof=Tp, idx_end=idx, method="bs23", rtol=3e-7, atol=1e-12, stiff_fallback=None), 4.0))))
print(json.dumps(dict(phase="sweep", **sample(lambda: s.sweep(T, P, L, U, method="fast"), 4.0))))
