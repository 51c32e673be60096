"""A/B of several builds of the library (alternating child processes): the Eon sweep at 2^20 LHS conditions with the explicit fast
path at the headline tolerance -- sweep ms, integrator ms, a checksum (the builds must agree to the bit).
python tools/r02_bs23_ab.py lib1.so lib2.so ...   (PFR_AB_METHOD=bs23|taylor4)"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import numpy as np, torch
    sys.path.insert(0, ROOT)
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
    from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions
    n = 1 << 20
    method = os.environ.get("PFR_AB_METHOD", "bs23")
    T, P, L, U = [torch.as_tensor(np.asarray(v, np.float32)).cuda() for v in lhs_conditions(n, seed=1)]
    s = Surrogate(ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eon"), mlp_mode="f16x3")
    ims = []
    for _ in range(3):
        r = s.sweep(T, P, L, U, method=method, rtol=3e-7, atol=1e-12)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        r = s.sweep(T, P, L, U, method=method, rtol=3e-7, atol=1e-12)
        ims.append(s.integrator_ms)
    e1.record(); torch.cuda.synchronize()
    st = r.stats.double()
    print(json.dumps({"lib": os.path.basename(os.environ.get("CRNN_PFR_LIB", "default")), "method": method, "sweep_ms": round(e0.elapsed_time(e1) / 5, 3),
                      "integrator_ms": round(s.integrator_ms(), 3), "checksum": float(r.y.sum()), "abs_checksum": float(r.y.abs().sum()),
                      "failed": int((r.status != 0).sum()), "evaluations": float(st[2].mean())}), flush=True)
else:
    libs = sys.argv[1:]
    for rep in range(2):
        for lib in libs:
            env = dict(os.environ, CRNN_PFR_LIB=os.path.abspath(lib))
            subprocess.run([sys.executable, __file__, "--child"], env=env, check=False)
