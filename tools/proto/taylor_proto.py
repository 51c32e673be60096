"""CPU prototype (tools only): explicit TAYLOR-SERIES integrator specialised to the CRNN right-hand side
    f(T, y) = W exp(kT(T) + nu^T ln clip(y)),
whose time derivatives along the solution follow from the standard series recurrences for log and exp:
    y_[i+1] = f_[i] / (i + 1),   L = ln Y: L_[i] = (y_[i] - (1/i) sum_{m=1}^{i-1} m L_[m] y_[i-m]) / y_[0],
    z_[i] = kT_[i] + nu^T L_[i],  r = exp z: r_[i] = (1/i) sum_{m=1}^{i} m z_[m] r_[i-m],   f_[i] = W r_[i]
so a step of order p costs ONE log / exp evaluation and 2 p mat-vecs instead of BS23's three full right-hand sides.
Knot-limited like bs23_kernel (T linear inside a knot interval: kT_[i] is analytic).  Compares right-hand-side work and
outlet error against BS23 on LHS conditions.  usage: python tools/proto/taylor_proto.py [n]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ros_proto import Model, ROOT, LB, UB, ZLO, ZHI, R_KCAL
from cheap_step_proto import integrate as bs23_integrate
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet
from n_hexane_pyrolysis_surrogate_reactor_model_b200.sweep import lhs_conditions
from oracle import reference_path as R
from oracle import c_oracle as CO


def taylor_coeffs(M, T0, s, y, p):
    """y_[0..p] at (T0, y) with dT/dt = s."""
    Y = np.clip(y, LB, UB)
    free = (y >= LB) & (y <= UB)
    q = np.where(free, 1.0 / Y, 0.0)
    rho = -s / T0
    yc = [y.copy()]
    L = [np.log(Y)]
    z0 = M.kT(T0) + L[0] @ M.nu
    zfree = (z0 >= ZLO) & (z0 <= ZHI)
    z = [z0]
    r = [np.exp(np.clip(z0, ZLO, ZHI))]
    yc.append(M.wout @ r[0])
    for i in range(1, p):
        # L_[i]
        acc = yc[i].copy()
        for m in range(1, i):
            acc -= (m / i) * L[m] * yc[i - m]
        L.append(acc * q)           # clamped species: ln Y constant
        kTi = -(M.Ea / R_KCAL) * rho ** i / T0 - M.b * rho ** i / i
        zi = np.where(zfree, kTi + L[i] @ M.nu, 0.0)
        z.append(zi)
        ri = np.zeros(9)
        for m in range(1, i + 1):
            ri += (m / i) * z[m] * r[i - m]
        r.append(ri)
        yc.append((M.wout @ ri) / (i + 1))
    return yc


def poly(yc, h, p):
    yn = yc[p].copy()
    for i in range(p - 1, -1, -1):
        yn = yn * h + yc[i]
    return yn


def crossing_time(yc, k, h, p):
    """smallest tau in (0, h] at which species k reaches LB (bisection on the step polynomial; the kernel would use Newton)"""
    lo, hi = 0.0, h
    s0 = yc[0][k] < LB
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        v = sum(yc[i][k] * mid ** i for i in range(p + 1))
        if (v < LB) == s0:
            lo = mid
        else:
            hi = mid
    return hi


def integrate(M, tg, Tp, kend, y0, rtol, atol, p=4, safety=0.9, events=False):
    """Order-p Taylor step, error estimate = the last term (error of the order p - 1 solution), exponent 1 / p."""
    y = y0.copy(); t = float(tg[0]); kc = 0
    n_steps = n_rej = 0
    hprop = None
    while kc < kend:
        tk, tk1 = float(tg[kc]), float(tg[kc + 1])
        Tk = np.float64(Tp[kc]); slope = (np.float64(Tp[kc + 1]) - Tk) / (tk1 - tk)
        Tt = Tk + slope * (t - tk)
        yc = taylor_coeffs(M, Tt, slope, y, p)
        if hprop is None:
            sk = atol + rtol * np.abs(y)
            d0 = np.sqrt(np.mean((y / sk) ** 2)); d1 = np.sqrt(np.mean((yc[1] / sk) ** 2))
            h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
            hprop = min(100 * h0, tg[kend] - t)
        while True:
            dist = tk1 - t
            clip = hprop * 1.01 >= dist
            h = dist if clip else hprop
            event = False
            yn = poly(yc, h, p)
            if events:
                crossed = (y < LB) != (yn < LB)
                if crossed.any():
                    h = min(crossing_time(yc, k, h, p) for k in np.nonzero(crossed)[0])
                    clip = False
                    event = True
                    yn = poly(yc, h, p)
            er = yc[p] * h ** p
            sk = atol + rtol * np.maximum(np.abs(y), np.abs(yn))
            err = np.sqrt(np.mean((er / sk) ** 2))
            n_steps += 1
            if np.isfinite(err) and err <= 1:
                f = min(6.0, max(0.2, safety * err ** (-1 / p))) if err > 0 else 6.0
                hprop = max(hprop, h * f) if (clip or event) else h * f    # a step cut short by a knot or an event does not shrink the proposal
                y = yn
                if clip:
                    t = tk1; kc += 1
                else:
                    t += h
                break
            n_rej += 1
            f = max(0.2, safety * err ** (-1 / p)) if np.isfinite(err) else 0.2
            hprop = h * min(f, 0.9)      # a rejected step re-uses the coefficients: only the polynomial is re-evaluated
    return y, n_steps, n_rej


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
    u0 = np.asarray(R.inlet_concentration(T, P), np.float32)
    truth, _ = CO.truth_batch(tfull, Tp, u0, ms.crnn.w_in, ms.crnn.w_b, ms.crnn.w_out, upto=idx)
    print(f"{n} LHS conditions, mean outlet knot {idx.mean():.0f}")

    def report(name, ys, extra):
        es = np.array([np.max(np.abs(np.clip(ys[i], LB, UB) - np.clip(truth[i], LB, UB)) / np.maximum(np.abs(truth[i]), 1e-3)) for i in range(n)])
        print(f"{name}: {extra}   outlet err median {np.median(es):.2e} p90 {np.percentile(es, 90):.2e} max {es.max():.2e}", flush=True)

    rtol, atol = 3e-7, 1e-12
    out = [bs23_integrate(M, tfull[i].astype(np.float64), Tp[i], int(idx[i]), u0[i].astype(np.float64), rtol, atol) for i in range(n)]
    report("bs23 3e-7/1e-12", [o[0] for o in out], f"rhs {np.mean([o[1] for o in out]):7.1f} steps {np.mean([o[2] + o[4] for o in out]):6.1f}")
    for p, rt, at, ev in ((4, 3e-7, 1e-12, False), (4, 3e-7, 1e-12, True), (4, 1e-6, 1e-12, True), (4, 3e-6, 1e-12, True), (4, 1e-5, 1e-12, True), (3, 3e-7, 1e-12, True),
                          (5, 1e-5, 1e-12, True), (4, 1e-6, 1e-10, True)):
        out = [integrate(M, tfull[i].astype(np.float64), Tp[i], int(idx[i]), u0[i].astype(np.float64), rt, at, p, events=ev) for i in range(n)]
        report(f"taylor-{p} {rt:g}/{at:g} events {ev}", [o[0] for o in out], f"evaluations {np.mean([o[1] - o[2] for o in out]):6.1f} (+{np.mean([o[2] for o in out]):5.1f} re-evaluated polynomials)")


if __name__ == "__main__":
    main()
