"""CPU prototype: explicit embedded Runge-Kutta pairs with knot-limited steps on the golden Eon conditions."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ros_proto import Model, ROOT, LB, UB
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet


def integrate(M, tg, Tp, kend, y0, tol, method="bs23", fac_max=6.0):
    y = y0.copy(); t = float(tg[0]); kc = 0
    n_acc = n_rej = n_f = 0
    hprop = None
    k1 = None
    errs = []
    p = 1 / 3 if method == "bs23" else 1 / 5
    while kc < kend:
        tk, tk1 = float(tg[kc]), float(tg[kc + 1])
        Tk = np.float64(Tp[kc]); slope = (np.float64(Tp[kc + 1]) - Tk) / (tk1 - tk)
        Tfun = lambda tt: Tk + slope * (tt - tk)
        if k1 is None:
            k1 = M.f(Tfun(t), y); n_f += 1
        if hprop is None:
            sk = tol + tol * np.abs(y)
            d0 = np.sqrt(np.mean((y / sk) ** 2)); d1 = np.sqrt(np.mean((k1 / sk) ** 2))
            h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
            hprop = min(100 * h0, tg[kend] - t)
        dist = tk1 - t
        clip = hprop * 1.01 >= dist
        h = dist if clip else hprop
        if method == "bs23":
            k2 = M.f(Tfun(t + h / 2), y + h / 2 * k1)
            k3 = M.f(Tfun(t + 3 * h / 4), y + 3 * h / 4 * k2)
            yn = y + h * (2 / 9 * k1 + 1 / 3 * k2 + 4 / 9 * k3)
            k4 = M.f(Tfun(t + h), yn); n_f += 3
            er = h * (-5 / 72 * k1 + 1 / 12 * k2 + 1 / 9 * k3 - 1 / 8 * k4)
            knew = k4
        else:  # dopri5
            a = [[1 / 5], [3 / 40, 9 / 40], [44 / 45, -56 / 15, 32 / 9], [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
                 [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656], [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]]
            c = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1, 1]
            ks = [k1]
            for i in range(6):
                yi = y + h * sum(a[i][j] * ks[j] for j in range(i + 1))
                ks.append(M.f(Tfun(t + c[i] * h), yi)); n_f += 1
            yn = yi
            e = [35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720, -2187 / 6784 + 12231 / 42400, 11 / 84 - 649 / 6300, -1 / 60]
            er = h * sum(e[j] * ks[j] for j in range(7))
            knew = ks[6]
        sk = tol + tol * np.maximum(np.abs(y), np.abs(yn))
        err = np.sqrt(np.mean((er / sk) ** 2))
        if np.isfinite(err) and err <= 1:
            f = min(fac_max, max(0.2, 0.9 * err ** (-p))) if err > 0 else fac_max
            hprop = max(hprop, h * f) if clip else h * f
            n_acc += 1; y = yn; k1 = knew
            if clip:
                errs.append(err); t = tk1; kc += 1
            else:
                t += h
        else:
            n_rej += 1
            f = max(0.2, 0.9 * err ** (-p)) if np.isfinite(err) else 0.2
            hprop = h * min(f, 0.9)
    return y, n_acc, n_rej, n_f, np.array(errs)


if __name__ == "__main__":
    g = np.load(os.path.join(ROOT, "tests/golden/reference_vectors.npz"))
    for mech in ("LLNL",):
        ms = ModelSet.from_packed(os.path.join(ROOT, f"tests/golden/containers/{mech}.npz"), "Eon")
        M = Model(ms.crnn.w_in, ms.crnn.w_b, ms.crnn.w_out)
        tg, Tp, idx, c0, truth = g["Eon/tgrid_full"], g["Eon/Tprof"], g["Eon/idx_cut"], g["c0"], g["Eon/truth_outlet"]
        for method in ("bs23", "dopri5"):
            for tol in (1e-5, 1e-6, 1e-7, 1e-8):
                acc = rej = nf = 0; es = []; ec = []
                for i in range(16):
                    y, a_, r_, f_, e_ = integrate(M, tg[i], Tp[i], int(idx[i]), c0[i].astype(np.float64), tol, method)
                    acc += a_; rej += r_; nf += f_; ec.append(e_)
                    es.append(np.max(np.abs(np.clip(y, LB, UB) - np.clip(truth[i], LB, UB)) / np.maximum(np.abs(truth[i]), 1e-3)))
                ec = np.concatenate(ec)
                print(f"{mech} {method} tol {tol:g}: acc {acc / 16:.1f} rej {rej / 16:.1f} rhs {nf / 16:.0f}  work/lane {nf / 16 * 185 / 1e3:.0f}k  outlet err max {max(es):.2e} median {np.median(es):.2e}"
                      f"  clipped-step err ratio median {np.median(ec):.1e} p99 {np.percentile(ec, 99):.1e}")
