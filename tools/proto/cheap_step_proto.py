"""CPU prototype (tools only): knot-limited BS23 where intervals that are SHORT against the controller's step proposal are
taken with a one-evaluation second-order step (Heun with the previous end slope reused as the start slope) instead of the
three-evaluation BS23 step.  Measures right-hand sides per trajectory and the outlet error against the converged solution on
LHS conditions (grids from the torch-CPU MLPs of the oracle), rtol / atol split as in the bench headline.

usage: python tools/proto/cheap_step_proto.py [n_conditions]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ros_proto import Model, ROOT, LB, UB
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions
from oracle import reference_path as R
from oracle import c_oracle as CO


def integrate(M, tg, Tp, kend, y0, rtol, atol, rho=0.0, refresh=0, variant="heun", rho3=0.0):
    """rho: an interval of length h <= rho * hprop is taken with the cheap step; refresh: force a BS23 step after this many cheap
    steps in a row (0 = never)."""
    y = y0.copy(); t = float(tg[0]); kc = 0
    n_f = n_cheap = n_bs = n_rej = 0
    hprop = None
    k1 = None
    run = 0
    e3 = None      # normalised error of the last BS23 step per h^3
    while kc < kend:
        tk, tk1 = float(tg[kc]), float(tg[kc + 1])
        Tk = np.float64(Tp[kc]); slope = (np.float64(Tp[kc + 1]) - Tk) / (tk1 - tk)
        Tfun = lambda tt: Tk + slope * (tt - tk)
        if k1 is None:
            k1 = M.f(Tfun(t), y); n_f += 1
        if hprop is None:
            sk = atol + rtol * np.abs(y)
            d0 = np.sqrt(np.mean((y / sk) ** 2)); d1 = np.sqrt(np.mean((k1 / sk) ** 2))
            h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
            hprop = min(100 * h0, tg[kend] - t)
        dist = tk1 - t
        clip = hprop * 1.01 >= dist
        h = dist if clip else hprop
        if variant == "heun_e" and clip and e3 is not None and e3 * h ** 3 <= rho and run < refresh:
            k2 = M.f(Tp[kc + 1], y + h * k1); n_f += 1
            y = y + 0.5 * h * (k1 + k2)
            k1 = k2
            n_cheap += 1; run += 1
            t = tk1; kc += 1
            continue
        if variant != "heun_e" and clip and rho3 > 0 and h <= rho3 * hprop and not (rho > 0 and h <= rho * hprop) and (refresh == 0 or run < refresh):
            # Kutta's third-order method with the previous end slope reused; the last stage sits at t + h on a second-order
            # approximation of y1 and is reused as the next start slope: two evaluations per step
            k2 = M.f(Tfun(t + h / 2), y + h / 2 * k1)
            k3 = M.f(Tp[kc + 1], y + h * (2 * k2 - k1)); n_f += 2
            y = y + h * (k1 + 4 * k2 + k3) / 6
            k1 = k3
            n_cheap += 1; run += 1
            t = tk1; kc += 1
            continue
        if variant != "heun_e" and clip and rho > 0 and h <= rho * hprop and (refresh == 0 or run < refresh):
            if variant == "heun":
                k2 = M.f(Tp[kc + 1], y + h * k1); n_f += 1
                y = y + 0.5 * h * (k1 + k2)
                k1 = k2
            elif variant == "heun2":      # two evaluations: exact end slope
                k2 = M.f(Tp[kc + 1], y + h * k1)
                y = y + 0.5 * h * (k1 + k2)
                k1 = M.f(Tp[kc + 1], y); n_f += 2
            n_cheap += 1; run += 1
            t = tk1; kc += 1
            continue
        run = 0
        k2 = M.f(Tfun(t + h / 2), y + h / 2 * k1)
        k3 = M.f(Tfun(t + 3 * h / 4), y + 3 * h / 4 * k2)
        yn = y + h * (2 / 9 * k1 + 1 / 3 * k2 + 4 / 9 * k3)
        k4 = M.f(Tfun(t + h) if not clip else Tp[kc + 1], yn); n_f += 3
        er = h * (-5 / 72 * k1 + 1 / 12 * k2 + 1 / 9 * k3 - 1 / 8 * k4)
        sk = atol + rtol * np.maximum(np.abs(y), np.abs(yn))
        err = np.sqrt(np.mean((er / sk) ** 2))
        if np.isfinite(err) and h > 0:
            e3 = err / h ** 3
        if np.isfinite(err) and err <= 1:
            f = min(6.0, max(0.2, 0.9 * err ** (-1 / 3))) if err > 0 else 6.0
            hprop = max(hprop, h * f) if clip else h * f
            n_bs += 1; y = yn; k1 = k4
            if clip:
                t = tk1; kc += 1
            else:
                t += h
        else:
            n_rej += 1
            f = max(0.2, 0.9 * err ** (-1 / 3)) if np.isfinite(err) else 0.2
            hprop = h * min(f, 0.9)
    return y, n_f, n_bs, n_cheap, n_rej


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    ms = ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eon")
    M = Model(ms.crnn.w_in, ms.crnn.w_b, ms.crnn.w_out)
    T, P, L, U = lhs_conditions(n, seed=7)
    tm = R.MLPParams(ms.time_mlp.w, ms.time_mlp.b, ms.time_mlp.out_min, ms.time_mlp.out_max)
    pm = R.MLPParams(ms.temp_mlp.w, ms.temp_mlp.b, ms.temp_mlp.out_min, ms.temp_mlp.out_max)
    ones = np.ones_like(T)
    tfull = R.time_grid(tm, T, P, ones * 1.0, ones * 2.5)
    tshort = R.time_grid(tm, T, P, L, U)
    Tp = R.temp_profile(pm, T, P)
    idx = np.array([R.eon_idx_cut(tfull[i], tshort[i, -1]) for i in range(n)], np.int32)
    c0 = R.inlet_concentration(T, P)
    u0 = np.asarray(c0, np.float32)
    truth, _ = CO.truth_batch(tfull, Tp, u0, ms.crnn.w_in, ms.crnn.w_b, ms.crnn.w_out, upto=idx)
    print(f"{n} LHS conditions, mean outlet knot {idx.mean():.0f}")
    rtol, atol = 3e-7, 1e-12
    for variant, rho, refresh, rho3 in (("heun", 0.0, 0, 0.0), ("heun_e", 1e-4, 8, 0.0), ("heun_e", 1e-3, 8, 0.0), ("heun_e", 1e-3, 3, 0.0), ("heun_e", 3e-3, 8, 0.0),
                                        ("heun_e", 1e-2, 8, 0.0), ("heun_e", 1e-2, 3, 0.0), ("heun_e", 3e-2, 8, 0.0)):
        nf = nb = nc = nr = 0; es = []
        for i in range(n):
            y, f_, b_, c_, r_ = integrate(M, tfull[i].astype(np.float64), Tp[i], int(idx[i]), u0[i].astype(np.float64), rtol, atol, rho, refresh, variant, rho3)
            nf += f_; nb += b_; nc += c_; nr += r_
            es.append(np.max(np.abs(np.clip(y, LB, UB) - np.clip(truth[i], LB, UB)) / np.maximum(np.abs(truth[i]), 1e-3)))
        es = np.array(es)
        print(f"{variant} rho {rho:4.2f} rho3 {rho3:4.2f} refresh {refresh:2d}: rhs {nf / n:7.1f}  bs23 steps {nb / n:6.1f} cheap {nc / n:6.1f} rej {nr / n:4.1f}   "
              f"outlet err median {np.median(es):.2e} p90 {np.percentile(es, 90):.2e} max {es.max():.2e}", flush=True)


if __name__ == "__main__":
    main()
