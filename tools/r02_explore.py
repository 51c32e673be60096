"""Round-2 exploration on one GPU (LLNL Eon, LHS conditions): what the cost sort, the gathered grid reads, the state
precision and the tolerance each cost or buy.  Prints JSON lines.  Usage: python tools/r02_explore.py [n] [tag]
(the library under test is chosen with CRNN_PFR_LIB)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet  # noqa: E402
from n_hexane_pyrolysis_surrogate_reactor_model_b200.surrogate import Surrogate  # noqa: E402
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def errs(y, ref):
    e = ((y - ref).abs() / torch.clamp(ref.abs(), min=1e-3)).amax(0)
    return dict(err_max=float(e.max()), err_p99=float(torch.quantile(e[:: max(1, e.numel() // 100000)], 0.99)), err_median=float(e.median()))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    tag = sys.argv[2] if len(sys.argv) > 2 else "default"
    quick = tag != "default"
    gold = os.path.join(ROOT, "tests", "golden", "containers", "LLNL.npz")
    T, P, L, U = (torch.as_tensor(a).cuda() for a in lhs_conditions(n, seed=13895))
    s = Surrogate(ModelSet.from_packed(gold, "Eon"))
    c0 = s.inlet_concentration(T, P)
    _, tend = s.time_grid(T, P, L, U, want_grid=False, want_end=True)
    tfull, _ = s.time_grid(T, P)
    Tp = s.temp_profile(T, P)
    idx = s.idx_cut(tfull, tend)
    perm = torch.argsort(idx, descending=True).to(torch.int32)
    sel = torch.arange(0, n, 16, device="cuda")
    sub = dict(tgrid=tfull[:, sel].contiguous(), Tprof=Tp[:, sel].contiguous(), idx_end=idx[sel].contiguous())
    ref = s.integrate(T[sel], c0[sel], method="rodas4", rtol=1e-11, atol=1e-11, **sub).y.clone()

    def emit(**row):
        print(json.dumps(dict(tag=tag, n=n, **row)), flush=True)

    def run(prec, tol, pm):
        return s.integrate(T, c0, tgrid=tfull, Tprof=Tp, idx_end=idx, perm=pm, method="bs23", precision=prec, rtol=tol, atol=tol)

    # tolerance sweep: time at full size, error on the sample
    for tol in (1e-8, 3e-9, 1e-9, 1e-10) if not quick else (1e-8, 3e-9, 1e-9, 3e-10, 1e-10):
        ms, res = timed(lambda: run(64, tol, perm))
        y = s.integrate(T[sel], c0[sel], method="bs23", rtol=tol, atol=tol, **sub).y
        st = res.stats.double()
        emit(exp="tol_sweep", precision=64, tol=tol, ms=ms, attempts=float((st[0] + st[1]).mean()), rhs=float(st[2].mean()),
             failed=int((res.status != 0).sum()), **errs(y, ref))
    if quick:
        return
    # the cost sort and the gather
    ms, _ = timed(lambda: run(64, 1e-8, None))
    emit(exp="no_sort_identity", ms=ms)
    p64 = perm.long()
    Ts, c0s, tfs, Tps, idxs = T[p64].contiguous(), c0[p64].contiguous(), tfull[:, p64].contiguous(), Tp[:, p64].contiguous(), idx[p64].contiguous()
    ms, _ = timed(lambda: s.integrate(Ts, c0s, tgrid=tfs, Tprof=Tps, idx_end=idxs, perm=None, method="bs23", rtol=1e-8, atol=1e-8))
    emit(exp="physically_sorted_identity", ms=ms)
    # proxy order: residence ratio L / u0 (what can be known before the MLPs run)
    proxy = torch.argsort(L / U, descending=True)
    Ts, c0s, tfs, Tps, idxs = T[proxy].contiguous(), c0[proxy].contiguous(), tfull[:, proxy].contiguous(), Tp[:, proxy].contiguous(), idx[proxy].contiguous()
    ms, _ = timed(lambda: s.integrate(Ts, c0s, tgrid=tfs, Tprof=Tps, idx_end=idxs, perm=None, method="bs23", rtol=1e-8, atol=1e-8))
    emit(exp="proxy_sorted_identity", ms=ms, spearman_note="order by L/u0 descending")
    del Ts, c0s, tfs, Tps, idxs
    # float32 state
    for tol in (1e-5, 1e-6, 1e-7):
        ms, res = timed(lambda: run(32, tol, perm))
        y = s.integrate(T[sel], c0[sel], method="bs23", precision=32, rtol=tol, atol=tol, **sub).y.double()
        st = res.stats.double()
        emit(exp="fp32_state", precision=32, tol=tol, ms=ms, attempts=float((st[0] + st[1]).mean()), failed=int((res.status != 0).sum()), **errs(y, ref))
    # Eoff float32
    del tfull, Tp
    s2 = Surrogate(ModelSet.from_packed(gold, "Eoff"))
    _, tend2 = s2.time_grid(T, P, L, U, want_grid=False, want_end=True)
    permT = torch.argsort(T, descending=True).to(torch.int32)
    ref2 = s2.integrate(T[sel], c0[sel], t_end=tend2[sel].contiguous(), method="rodas4", rtol=1e-11, atol=1e-11).y.clone()
    for prec, tol in ((64, 1e-7), (32, 1e-5), (32, 1e-6), (32, 1e-7)):
        ms, res = timed(lambda: s2.integrate(T, c0, t_end=tend2, perm=permT, method="dp54", precision=prec, rtol=tol, atol=tol))
        y = s2.integrate(T[sel], c0[sel], t_end=tend2[sel].contiguous(), method="dp54", precision=prec, rtol=tol, atol=tol).y.double()
        emit(exp="eoff_dp54", precision=prec, tol=tol, ms=ms, failed=int((res.status != 0).sum()), **errs(y, ref2))


if __name__ == "__main__":
    main()
