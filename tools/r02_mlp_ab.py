"""A/B timing of one time-MLP pass (2^20 conditions, t_end-only and full-grid modes) for several builds of the library in ONE
process order that alternates them, so that clock / box differences cancel: python tools/r02_mlp_ab.py lib1.so lib2.so ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import numpy as np, torch
    sys.path.insert(0, ROOT)
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions
    n = 1 << 20
    s = Surrogate(ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eon"))
    T, P, L, U = [torch.as_tensor(np.asarray(v, np.float32)).cuda() for v in lhs_conditions(n, seed=1)]
    grid = torch.empty((801, n), dtype=torch.float32, device="cuda")
    out = []
    for name, fn in (("t_end only", lambda: s.time_grid(T, P, L, U, want_grid=False, want_end=True)),
                     ("full grid", lambda: s.time_grid(T, P, None, None, want_grid=True, out=grid)),
                     ("sweep bs23", lambda: s.sweep(T, P, L, U, method="bs23", rtol=3e-7, atol=1e-12))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            r = fn()
        e1.record(); torch.cuda.synchronize()
        out.append(f"{name} {e0.elapsed_time(e1) / 5:.2f} ms")
    print(os.path.basename(os.environ.get("CRNN_PFR_LIB", "default")), " | ".join(out), flush=True)
else:
    libs = sys.argv[1:]
    for rep in range(2):
        for lib in libs:
            env = dict(os.environ, CRNN_PFR_LIB=os.path.abspath(lib))
            subprocess.run([sys.executable, __file__, "--child"], env=env, check=False)
