"""CPU prototype: ROS3 on the knot-limited Eon path with the factored matrix E^-1 re-used across steps.
A re-use step keeps J (evaluated at an earlier state) and corrects for the changed step size with a truncated Neumann
series  (E_old + eps I)^-1 = E_old^-1 (I - eps E_old^-1 + eps^2 E_old^-2 ...),  eps = 1/(h gamma) - 1/(h_old gamma)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ros_proto import Model, make_tableaux, ROOT, LB, UB
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet

G = 0.43586652150845899941601945119356
C21, C31, C32 = -1.0156171083877702, 4.07599564525377, 9.207679429833079
M2, M3 = 6.169794704382824, -0.4277225654321857
E1, E2, E3 = 0.5, -2.9079558716805469, 0.2235406989781157
D1, D2, D3 = G, 0.24291996454816804, 2.1851380027664058


def integrate(M, tg, Tp, kend, y0, tol, reuse_max, theta, hband, ncorr):
    y = y0.copy(); t = float(tg[0]); kc = 0
    hprop = None
    n_acc = n_rej = n_fresh = 0
    Einv = None; h_old = None; age = 0; last_err = 1.0
    force_fresh = True
    while kc < kend:
        tk, tk1 = float(tg[kc]), float(tg[kc + 1])
        Tk = np.float64(Tp[kc]); slope = (np.float64(Tp[kc + 1]) - Tk) / (tk1 - tk)
        Tfun = lambda tt: Tk + slope * (tt - tk)
        if hprop is None:
            f0 = M.f(Tfun(t), y); sk = tol + tol * np.abs(y)
            d0 = np.sqrt(np.mean((y / sk) ** 2)); d1 = np.sqrt(np.mean((f0 / sk) ** 2))
            h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
            hprop = min(100 * h0, tg[kend] - t)
        dist = tk1 - t
        clip = hprop * 1.01 >= dist
        h = dist if clip else hprop
        fresh = force_fresh or Einv is None or age >= reuse_max or last_err > theta or abs(h / h_old - 1) > hband
        f0, J, g = M.jac(Tfun(t), y)     # (the prototype evaluates J always; a re-use step only needs f0 and g)
        ft = M.wout @ (g * M.dkT(Tfun(t))) * slope
        if fresh:
            Einv = np.linalg.inv(np.eye(9) / (h * G) - J); h_old = h; age = 0; n_fresh += 1
            solve = lambda b: Einv @ b
        else:
            eps = 1 / (h * G) - 1 / (h_old * G)
            def solve(b):
                x = Einv @ b
                c = x
                for _ in range(ncorr):
                    c = -eps * (Einv @ c)
                    x = x + c
                return x
        k1 = solve(f0 + h * D1 * ft)
        f2 = M.f(Tfun(t + G * h), y + k1)
        k2 = solve(f2 + C21 / h * k1 + h * D2 * ft)
        k3 = solve(f2 + (C31 * k1 + C32 * k2) / h + h * D3 * ft)
        yn = y + k1 + M2 * k2 + M3 * k3
        er = E1 * k1 + E2 * k2 + E3 * k3
        sk = tol + tol * np.maximum(np.abs(y), np.abs(yn))
        err = np.sqrt(np.mean((er / sk) ** 2))
        if np.isfinite(err) and err <= 1:
            f = min(6.0, max(0.2, 0.9 * err ** (-1 / 3))) if err > 0 else 6.0
            hprop = max(hprop, h * f) if clip else h * f
            n_acc += 1; y = yn; age += 1; last_err = err; force_fresh = False
            if clip: t = tk1; kc += 1
            else: t += h
        else:
            n_rej += 1
            f = max(0.2, 0.9 * err ** (-1 / 3)) if np.isfinite(err) else 0.2
            hprop = h * min(f, 0.9)
            force_fresh = True
    return y, n_acc, n_rej, n_fresh


if __name__ == "__main__":
    g = np.load(os.path.join(ROOT, "tests/golden/reference_vectors.npz"))
    ms = ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eon")
    M = Model(ms.crnn.w_in, ms.crnn.w_b, ms.crnn.w_out)
    tg, Tp, idx, c0, truth = g["Eon/tgrid_full"], g["Eon/Tprof"], g["Eon/idx_cut"], g["c0"], g["Eon/truth_outlet"]
    print("knot spacing ratio h_{k+1}/h_k percentiles:", np.percentile(np.abs(np.diff(tg, axis=1)[:, 1:] / np.diff(tg, axis=1)[:, :-1] - 1), [50, 90, 99]))
    FRESH, REUSE = 1280.0, 772.0
    for tol in (1e-7,):
        for (rm, th, hb, nc) in ((0, 0, 0, 0), (2, 1e-2, 0.1, 2), (4, 1e-2, 0.1, 2), (8, 1e-2, 0.1, 2), (4, 1e-1, 0.2, 2), (4, 1e-2, 0.2, 3), (8, 1e-1, 0.3, 3), (16, 1e-1, 0.3, 3)):
            acc = rej = fresh = 0; worst = 0; es = []
            for i in range(16):
                y, a_, r_, f_ = integrate(M, tg[i], Tp[i], int(idx[i]), c0[i].astype(np.float64), tol, rm, th, hb, nc)
                acc += a_; rej += r_; fresh += f_
                e = np.max(np.abs(np.clip(y, LB, UB) - np.clip(truth[i], LB, UB)) / np.maximum(np.abs(truth[i]), 1e-3))
                es.append(e)
            steps = acc + rej
            work = fresh * FRESH + (steps - fresh) * (REUSE + (nc - 2) * 81)
            print(f"tol {tol:g} reuse_max {rm} theta {th:g} hband {hb} ncorr {nc}: steps {steps / 16:.1f} fresh {fresh / 16:.1f} rej {rej / 16:.1f} "
                  f"work/traj {work / 16 / 1e3:.0f}k (vs all-fresh {steps * FRESH / 16 / 1e3:.0f}k)  outlet err max {max(es):.2e} median {np.median(es):.2e}")
