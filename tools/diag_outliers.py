"""Dump the LHS conditions on which the Rosenbrock kernels deviate most from their own tight-tolerance solution, with the
grids they were integrated on, so that the converged oracle solution can be computed for them on the CPU."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
T, P, L, U = (torch.as_tensor(a).cuda() for a in lhs_conditions(n, seed=13895))
s = Surrogate(ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eon"))
c0 = s.inlet_concentration(T, P)
_, tend = s.time_grid(T, P, L, U, want_grid=False, want_end=True)
tfull, _ = s.time_grid(T, P)
Tp = s.temp_profile(T, P)
idx = s.idx_cut(tfull, tend)
perm = torch.argsort(idx, descending=True).to(torch.int32)
run = lambda meth, tol: s.integrate(T, c0, tgrid=tfull, Tprof=Tp, idx_end=idx, perm=perm, method=meth, rtol=tol, atol=tol)
ref = run("rodas4", 1e-11)
out = {}
sel = []
scale = torch.clamp(ref.y.abs(), min=1e-3)
runs = {"rodas4_1e-11": ref}
for meth, tol in (("rodas4", 1e-9), ("ros3", 1e-9), ("ros3", 1e-7), ("rodas4", 1e-6), ("rodas4", 1e-12)):
    r = run(meth, tol)
    runs[f"{meth}_{tol:g}"] = r
    e = ((r.y - ref.y).abs() / scale).amax(0)
    top = torch.topk(e, 12).indices
    print(meth, tol, "worst", e[top].tolist()[:4], "idx", idx[top].tolist()[:4])
    sel.append(top)
sel = torch.unique(torch.cat(sel + [torch.arange(0, n, n // 64, device=T.device)]))
for k, r in runs.items():
    out[f"y/{k}"] = r.y[:, sel].cpu().numpy().T
    out[f"stats/{k}"] = r.stats[:, sel].cpu().numpy().T
out.update(sel=sel.cpu().numpy(), T=T[sel].cpu().numpy(), P=P[sel].cpu().numpy(), L=L[sel].cpu().numpy(), U=U[sel].cpu().numpy(),
           c0=c0[sel].cpu().numpy(), tgrid=tfull[:, sel].cpu().numpy().T, Tprof=Tp[:, sel].cpu().numpy().T, idx=idx[sel].cpu().numpy())
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "outliers.npz"), **out)
print("saved", len(sel))
