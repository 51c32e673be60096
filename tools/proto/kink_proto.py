"""CPU prototype for DESIGN 9(1): BS23 on the knot-limited Eon path with a kink-aware acceptance (a step during which a
species crosses the lower state clamp must meet the tolerance with a margin) and separate rtol / atol, on the 16 golden
conditions plus the worst LHS conditions dumped by tools/diag_outliers.py (if present).  Work = step attempts."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ros_proto import Model, ROOT, LB, UB
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet


def integrate(M, tg, Tp, kend, y0, rtol, atol, kink):
    y = y0.copy(); t = float(tg[0]); kc = 0; n_try = 0; hprop = None; k1 = None
    while kc < kend:
        tk, tk1 = float(tg[kc]), float(tg[kc + 1])
        Tk = np.float64(Tp[kc]); slope = (np.float64(Tp[kc + 1]) - Tk) / (tk1 - tk)
        Tf = lambda tt: Tk + slope * (tt - tk)
        if k1 is None:
            k1 = M.f(Tf(t), y)
            sk = atol + rtol * np.abs(y)
            d0 = np.sqrt(np.mean((y / sk) ** 2)); d1 = np.sqrt(np.mean((k1 / sk) ** 2))
            hprop = min(100 * (1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1), tg[kend] - t)
        dist = tk1 - t
        clip = hprop * 1.01 >= dist
        h = dist if clip else hprop
        k2 = M.f(Tf(t + h / 2), y + h / 2 * k1)
        k3 = M.f(Tf(t + 3 * h / 4), y + 3 * h / 4 * k2)
        yn = y + h * (2 / 9 * k1 + 1 / 3 * k2 + 4 / 9 * k3)
        k4 = M.f(Tf(t + h), yn)
        er = h * (-5 / 72 * k1 + 1 / 12 * k2 + 1 / 9 * k3 - 1 / 8 * k4)
        err = np.sqrt(np.mean((er / (atol + rtol * np.maximum(np.abs(y), np.abs(yn)))) ** 2))
        if kink and np.any((y < LB) != (yn < LB)):
            err *= kink
        n_try += 1
        fac = 0.9 * max(err, 1e-30) ** (-1 / 3)
        if np.isfinite(err) and err <= 1:
            f = min(6.0, max(0.2, fac)); hprop = max(hprop, h * f) if clip else h * f
            y = yn; k1 = k4
            if clip: t = tk1; kc += 1
            else: t += h
        else:
            hprop = h * min(max(0.2, fac), 0.9)
    return y, n_try


if __name__ == "__main__":
    from oracle import c_oracle as CO
    g = np.load(os.path.join(ROOT, "tests/golden/reference_vectors.npz"))
    ms = ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eon")
    M = Model(ms.crnn.w_in, ms.crnn.w_b, ms.crnn.w_out)
    tg, Tp, idx, c0, truth = g["Eon/tgrid_full"], g["Eon/Tprof"], g["Eon/idx_cut"], g["c0"].astype(np.float64), g["Eon/truth_outlet"]
    out = os.path.join(ROOT, "gpurun_out", "outliers.npz")
    if os.path.exists(out):
        d = np.load(out)
        c1 = np.zeros((len(d["T"]), 9)); c1[:, 6] = d["c0"]
        t1, _ = CO.truth_batch(d["tgrid"].copy(), d["Tprof"].copy(), c1.astype(np.float32), ms.crnn.w_in, ms.crnn.w_b, ms.crnn.w_out,
                               upto=d["idx"].astype(np.int32), nthreads=8)
        sel = np.arange(0, len(d["T"]), 3)
        tg, Tp, idx, c0, truth = (np.concatenate([a, b[sel]]) for a, b in ((tg, d["tgrid"]), (Tp, d["Tprof"]), (idx, d["idx"]), (c0, c1), (truth, t1)))
    truth = np.clip(truth, LB, UB)
    print(f"{len(idx)} conditions, mean outlet knot {idx.mean():.0f}")
    for rtol, atol, kink in ((1e-8, 1e-8, 0), (1e-7, 1e-7, 0), (1e-7, 1e-7, 100), (1e-7, 1e-8, 100), (1e-7, 1e-8, 10), (3e-8, 1e-8, 100), (1e-7, 1e-9, 100)):
        es, tries = [], 0
        for i in range(len(idx)):
            y, n = integrate(M, tg[i], Tp[i], int(idx[i]), c0[i], rtol, atol, kink)
            tries += n
            es.append(np.max(np.abs(np.clip(y, LB, UB) - truth[i]) / np.maximum(np.abs(truth[i]), 1e-3)))
        es = np.array(es)
        print(f"rtol {rtol:g} atol {atol:g} kink {kink:4d}: attempts/condition {tries / len(idx):7.1f}  error median {np.median(es):.1e} p90 {np.percentile(es, 90):.1e} max {es.max():.1e}")
