"""A/B of several builds of the library (CRNN_PFR_LIB) in alternating child processes: one time-MLP pass (t_end only / full grid)
and the Eon / Eoff sweeps at 2^20 LHS conditions, mlp_mode from PFR_AB_MODE (default f16x3).
python tools/r02_lib_ab.py lib1.so lib2.so ..."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import numpy as np, torch
    sys.path.insert(0, ROOT)
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions
    n = 1 << 20
    mode = os.environ.get("PFR_AB_MODE", "f16x3")
    T, P, L, U = [torch.as_tensor(np.asarray(v, np.float32)).cuda() for v in lhs_conditions(n, seed=1)]
    grid = torch.empty((801, n), dtype=torch.float32, device="cuda")

    def timeit(fn, reps=5):
        for _ in range(3):
            keep = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            keep = fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, keep

    s = Surrogate(ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eon"), mlp_mode=mode)
    a, _ = timeit(lambda: s.time_grid(T, P, L, U, want_grid=False, want_end=True))
    b, _ = timeit(lambda: s.time_grid(T, P, None, None, want_grid=True, out=grid))
    c, r = timeit(lambda: s.sweep(T, P, L, U, method="bs23", rtol=3e-7, atol=1e-12))
    rec = {"lib": os.path.basename(os.environ.get("CRNN_PFR_LIB", "default")), "mode": mode, "t_end_only_ms": round(a, 3), "full_grid_ms": round(b, 3),
           "eon_sweep_ms": round(c, 3), "eon_integrator_ms": round(s.integrator_ms(), 3), "eon_checksum": float(r.y.sum())}
    del s, r
    torch.cuda.empty_cache()
    s = Surrogate(ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eoff"), mlp_mode=mode)
    d, r = timeit(lambda: s.sweep(T, P, L, U, method="dp54", rtol=1e-7, atol=1e-7))
    rec.update({"eoff_sweep_ms": round(d, 3), "eoff_checksum": float(r.y.sum())})
    print(json.dumps(rec), flush=True)
else:
    libs = sys.argv[1:]
    for rep in range(2):
        for lib in libs:
            env = dict(os.environ, CRNN_PFR_LIB=os.path.abspath(lib))
            subprocess.run([sys.executable, __file__, "--child"], env=env, check=False)
