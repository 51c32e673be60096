"""CPU prototype: knot-limited Rosenbrock stepping on the golden Eon conditions with different tableaux.
Measures accepted / rejected steps, a work estimate and the outlet error against the converged oracle solution."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from n_hexane_pyrolysis_surrogate_reactor_model_b200.containers import ModelSet

R_KCAL = np.float64(np.float32(1.9872036e-3))
LB, UB, ZLO, ZHI, DULO, DUHI = 1e-6, 60.0, -30.0, 30.0, -1e10, 1e10

def make_tableaux():
    T = {}
    T["rodas4"] = dict(gamma=0.25,
        a=[[], [1.544], [0.9466785280815826, 0.2557011698983284], [3.314825187068521, 2.896124015972201, 0.9986419139977817],
           [1.221224509226641, 6.019134481288629, 12.53708332932087, -0.687886036105895], None],
        C=[[], [-5.6688], [-2.430093356833875, -0.2063599157091915], [-0.1073529058151375, -9.594562251023355, -20.47028614809616],
           [7.496443313967647, -10.24680431464352, -33.99990352819905, 11.7089089320616],
           [8.083246795921522, -7.981132988064893, -31.52159432874371, 16.31930543123136, -6.058818238834054]],
        c=[0, 0.386, 0.21, 0.63, 1, 1], d=[0.25, -0.1043, 0.1035, -0.03620000000000023, 0, 0], special="rodas4")
    ig = 1 / (0.5 + np.sqrt(3) / 6)
    T["ros3p"] = dict(gamma=1 / ig, a=[[], [ig], [ig, 0.0]], C=[[], [-ig * ig], [-3.464101615137755, -1.732050807568877]],
        c=[0, 1, 1], d=[0.7886751345948129, -0.2113248654051871, -1.077350269189626],
        m=[2.0, 0.5773502691896258, 0.4226497308103742], mh=[2.113248654051871, 1.0, 0.4226497308103742], newf=[1, 1, 0])
    T["rodas3"] = dict(gamma=0.5, a=[[], [0.0], [2.0, 0.0], [2.0, 0.0, 1.0]], C=[[], [4.0], [1.0, -1.0], [1.0, -1.0, -8.0 / 3.0]],
        c=[0, 0, 1, 1], d=[0.5, 1.5, 0, 0], m=[2.0, 0, 1.0, 1.0], e=[0, 0, 0, 1.0], newf=[1, 0, 1, 1])
    g = 0.43586652150845899941601945119356
    T["ros3"] = dict(gamma=g, a=[[], [1.0], [1.0, 0.0]], C=[[], [-1.0156171083877702091975600115545], [4.0759956452537699824805835358067, 9.2076794298330791242156818474003]],
        c=[0, g, g], d=[g, 0.24291996454816804366592249683314, 2.1851380027664058511513169485832],
        m=[1.0, 6.1697947043828245592553615689730, -0.4277225654321857332623837380651],
        e=[0.5, -2.9079558716805469821718236208017, 0.2235406989781156962736090927619], newf=[1, 1, 0])
    g2 = 1 + 1 / np.sqrt(2)
    T["ros2"] = dict(gamma=g2, a=[[], [1 / g2]], C=[[], [-2 / g2]], c=[0, 1], d=[g2, -g2], m=[1.5 / g2, 0.5 / g2], e=[0.5 / g2, 0.5 / g2], newf=[1, 1])
    T["ros4"] = dict(gamma=0.57282, a=[[], [2.0], [1.867943637803922, 0.2344449711399156], [1.867943637803922, 0.2344449711399156, 0.0]],
        C=[[], [-7.137615036412310], [2.580708087951457, 0.6515950076447975], [-2.137148994382534, -0.3214669691237626, -0.6949742501781779]],
        c=[0, 1.14564, 0.65521686381559, 0.65521686381559], d=[0.57282, -1.769193891319233, 0.7592633437920482, -0.1049021087100450],
        m=[2.255570073418735, 0.2870493262186792, 0.4353179431840180, 1.093502252409163],
        e=[-0.2815431932141155, -0.07276199124938920, -0.1082196201495311, -1.093502252409163], newf=[1, 1, 1, 0])
    return T

class Model:
    def __init__(self, w_in, w_b, w_out):
        self.nu = np.asarray(w_in[:9], np.float64)       # [k][j]
        self.Ea = np.asarray(w_in[9], np.float64); self.b = np.asarray(w_in[10], np.float64)
        self.lnA = np.asarray(w_b, np.float64); self.wout = np.asarray(w_out, np.float64)
    def kT(self, T):
        return self.lnA + self.Ea * (-1 / (R_KCAL * T)) + self.b * np.log(T)
    def dkT(self, T):
        return (self.b + self.Ea / (R_KCAL * T)) / T
    def f(self, T, y):
        Y = np.clip(y, LB, UB)
        z = self.kT(T) + np.log(Y) @ self.nu
        r = np.exp(np.clip(z, ZLO, ZHI))
        return self.wout @ r
    def jac(self, T, y):
        Y = np.clip(y, LB, UB)
        z = self.kT(T) + np.log(Y) @ self.nu
        r = np.exp(np.clip(z, ZLO, ZHI))
        g = np.where((z >= ZLO) & (z <= ZHI), r, 0.0)
        q = np.where((y >= LB) & (y <= UB), 1 / Y, 0.0)
        J = (self.wout * g) @ self.nu.T * q
        return self.wout @ r, J, g

def ros_step(M, tab, t, y, h, Tfun, slope):
    """One step in Hairer's transformed form. Returns ynew, err vector, number of f evals."""
    s = len(tab["c"])
    f0, J, g = M.jac(Tfun(t), y)
    ft = M.wout @ (g * M.dkT(Tfun(t))) * slope
    E = np.eye(9) / (h * tab["gamma"]) - J
    Einv = np.linalg.inv(E)
    k = []
    nf = 1
    if tab.get("special") == "rodas4":
        a, C, c, d = tab["a"], tab["C"], tab["c"], tab["d"]
        k.append(Einv @ (f0 + h * d[0] * ft))
        yn = None
        for i in range(1, 4):
            yi = y + sum(a[i][j] * k[j] for j in range(i))
            fi = M.f(Tfun(t + c[i] * h), yi); nf += 1
            k.append(Einv @ (fi + sum(C[i][j] / h * k[j] for j in range(i)) + h * d[i] * ft))
        yi = y + sum(a[4][j] * k[j] for j in range(4))
        fi = M.f(Tfun(t + h), yi); nf += 1
        k.append(Einv @ (fi + sum(C[4][j] / h * k[j] for j in range(4))))
        yi = yi + k[4]
        fi = M.f(Tfun(t + h), yi); nf += 1
        k.append(Einv @ (fi + sum(C[5][j] / h * k[j] for j in range(5))))
        return yi + k[5], k[5], nf
    a, C, c, d, m = tab["a"], tab["C"], tab["c"], tab["d"], tab["m"]
    fi = f0
    for i in range(s):
        if i > 0 and tab["newf"][i]:
            yi = y + sum(a[i][j] * k[j] for j in range(i))
            fi = M.f(Tfun(t + c[i] * h), yi); nf += 1
        k.append(Einv @ (fi + sum(C[i][j] / h * k[j] for j in range(i)) + h * d[i] * ft))
    ynew = y + sum(m[i] * k[i] for i in range(s))
    if "e" in tab:
        err = sum(tab["e"][i] * k[i] for i in range(s))
    else:
        err = sum((m[i] - tab["mh"][i]) * k[i] for i in range(s))
    return ynew, err, nf

def order_test(name, tab):
    """nonautonomous stiff-ish scalar-coupled test on the CRNN itself with a linear T ramp"""
    ms = ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eon")
    M = Model(ms.crnn.w_in, ms.crnn.w_b, ms.crnn.w_out)
    y0 = np.load(os.path.join(ROOT, "tests/golden/reference_vectors.npz"))["Eon/truth_knots_every50"][0][3].copy()
    Tfun = lambda t: 1000.0 + 400.0 * t
    tend = 0.05
    def run(nsteps, est=False):
        y = y0.copy(); h = tend / nsteps; t = 0.0
        e_acc = 0
        for _ in range(nsteps):
            yn, er, _ = ros_step(M, tab, t, y, h, Tfun, 400.0)
            if est: y = yn - er   # embedded solution
            else: y = yn
            t += h
        return y
    ref = run(4096) if name != "ros2" else None
    tabs = make_tableaux()
    yref = None
    # reference with rodas4 fine
    y = y0.copy(); t = 0.0; N = 8192; h = tend / N
    for _ in range(N):
        y, _, _ = ros_step(M, tabs["rodas4"], t, y, h, Tfun, 400.0); t += h
    yref = y
    out = []
    for N in (16, 32, 64, 128):
        e1 = np.max(np.abs(run(N) - yref) / np.maximum(np.abs(yref), 1e-3))
        e2 = np.max(np.abs(run(N, True) - yref) / np.maximum(np.abs(yref), 1e-3))
        out.append((N, e1, e2))
    for (N, e1, e2), (N2, f1, f2) in zip(out[:-1], out[1:]):
        print(f"  {name}: N {N}->{N2}: order {np.log2(e1 / f1):.2f} (err {f1:.2e}), embedded {np.log2(e2 / f2):.2f} (err {f2:.2e})")

def integrate(M, tab, order_emb, tg, Tp, kend, y0, rtol, atol, fac_max=6.0):
    n_acc = n_rej = n_f = 0
    y = y0.copy(); t = float(tg[0]); kc = 0
    hprop = None
    p = 1.0 / (order_emb + 1)
    errs_clip = []
    while kc < kend:
        tk, tk1 = float(tg[kc]), float(tg[kc + 1])
        Tk = np.float64(Tp[kc]); slope = (np.float64(Tp[kc + 1]) - Tk) / (tk1 - tk)
        Tfun = lambda tt: Tk + slope * (tt - tk)
        if hprop is None:
            f0 = M.f(Tfun(t), y); sk = atol + rtol * np.abs(y)
            d0 = np.sqrt(np.mean((y / sk) ** 2)); d1 = np.sqrt(np.mean((f0 / sk) ** 2))
            h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
            hprop = min(100 * h0, tg[kend] - t)
        dist = tk1 - t
        clip = hprop * 1.01 >= dist
        h = dist if clip else hprop
        yn, er, nf = ros_step(M, tab, t, y, h, Tfun, slope); n_f += nf
        sk = atol + rtol * np.maximum(np.abs(y), np.abs(yn))
        err = np.sqrt(np.mean((er / sk) ** 2))
        if np.isfinite(err) and err <= 1:
            f = min(fac_max, max(0.2, 0.9 * err ** (-p))) if err > 0 else fac_max
            hprop = max(hprop, h * f) if clip else h * f
            n_acc += 1; y = yn
            if clip:
                errs_clip.append(err)
                t = tk1; kc += 1
            else:
                t += h
        else:
            n_rej += 1
            f = max(0.2, 0.9 * err ** (-p)) if np.isfinite(err) else 0.2
            hprop = h * min(f, 0.9)
    return y, n_acc, n_rej, n_f, np.array(errs_clip)

if __name__ == "__main__":
    tabs = make_tableaux()
    if "order" in sys.argv:
        for nm in ("ros4",):
            order_test(nm, tabs[nm])
    g = np.load(os.path.join(ROOT, "tests/golden/reference_vectors.npz"))
    ms = ModelSet.from_packed(os.path.join(ROOT, "tests/golden/containers/LLNL.npz"), "Eon")
    M = Model(ms.crnn.w_in, ms.crnn.w_b, ms.crnn.w_out)
    tg, Tp, idx, c0, truth = g["Eon/tgrid_full"], g["Eon/Tprof"], g["Eon/idx_cut"], g["c0"], g["Eon/truth_outlet"]
    emb = dict(ros4=3, rodas4=3, ros3p=2, rodas3=2, ros3=2, ros2=1)
    # work model (FP64 instr): jac 891 + factor 285 + per solve 81 + per f 421 + sums
    for tol in (1e-5, 1e-6, 1e-7, 1e-8):
        for nm in ("rodas4", "ros4", "ros3p", "ros3"):
            tab = tabs[nm]; s = len(tab["c"])
            acc = rej = nf = 0; worst = 0; errs = []
            ncond = 16
            for i in range(ncond):
                y, a_, r_, f_, ec = integrate(M, tab, emb[nm], tg[i], Tp[i], int(idx[i]), c0[i].astype(np.float64), tol, tol)
                acc += a_; rej += r_; nf += f_; errs.append(ec)
                e = np.max(np.abs(np.clip(y, LB, UB) - np.clip(truth[i], LB, UB)) / np.maximum(np.abs(truth[i]), 1e-3))
                worst = max(worst, e)
            steps = acc + rej
            work = steps * (891 + 285 + 43 + 81 * s + 40 * s) + nf * 421
            ec = np.concatenate(errs)
            print(f"tol {tol:g} {nm:7s}: acc {acc / ncond:7.1f} rej {rej / ncond:5.1f} f-evals {nf / ncond:7.1f} work/traj {work / ncond / 1e6:.3f} M  "
                  f"worst outlet err {worst:.2e}  clipped-step err ratio median {np.median(ec):.1e} p99 {np.percentile(ec, 99):.1e} knots {np.mean(idx[:ncond]):.0f}")
