"""Small fixed workload for ncu captures: one Eon integrate, one Eoff integrate, the three MLP passes.
Usage: python tools/profile_target.py [n] [precision] [method] [tol]   (PFR_ATOL: absolute tolerance of the Eon integrate if it differs from tol)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet  # noqa: E402
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate  # noqa: E402
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    prec = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    method = sys.argv[3] if len(sys.argv) > 3 else "rodas4"
    tol = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-6
    gold = os.path.join(ROOT, "tests", "golden", "containers", "LLNL.npz")
    T, P, L, U = (torch.as_tensor(a).cuda() for a in lhs_conditions(n, seed=13895))
    on = Surrogate(ModelSet.from_packed(gold, "Eon"))
    off = Surrogate(ModelSet.from_packed(gold, "Eoff"))
    for _ in range(2):
        r1 = on.sweep(T, P, L, U, precision=prec, method=method, rtol=tol, atol=float(os.environ.get("PFR_ATOL", tol)))
        r2 = off.sweep(T, P, L, U, precision=prec, method=os.environ.get("PFR_EOFF_METHOD", "rodas4"), rtol=float(os.environ.get("PFR_EOFF_TOL", "1e-6")), atol=float(os.environ.get("PFR_EOFF_TOL", "1e-6")))
    torch.cuda.synchronize()
    print("ok", float(r1.y.sum()), float(r2.y.sum()), int(r1.status.sum()), int(r2.status.sum()))


if __name__ == "__main__":
    main()
