// taylor4_kernel: explicit TAYLOR-SERIES integrator of order 4 for the coupled (Eon) sweep, one PFR condition per thread,
// knot-limited like bs23_kernel (integrate_explicit.cuh), whose work queue, coefficient broadcasts and controller it shares.
//
// Why.  The CRNN right-hand side is   f(T, y) = W exp( kT(T) + nu^T ln clip(y) ),   and along a solution every time
// derivative of it follows from the power-series recurrences of log and exp (automatic differentiation of the ODE; y_[i] is the
// i-th Taylor coefficient of y(t + tau) in tau):
//     L = ln Y :  L_[i] = ( y_[i] - (1/i) sum_{m=1}^{i-1} m L_[m] y_[i-m] ) / y_[0]          (0 for a species held by a clamp)
//     z_[i] = kT_[i] + nu^T L_[i]                                                              (0 for a clamped exponent)
//     r = e^z  :  r_[i] = (1/i) sum_{m=1}^{i} m z_[m] r_[i-m]
//     y_[i+1] = W r_[i] / (i + 1)
// Inside a knot interval T is linear in t, so kT_[i] is analytic:  with rho = -(dT/dt) / T,
//     kT_j = lnA_j - Ea_j / (R T) + b_j ln T   =>   kT_[i],j = rho^i ( -Ea_j / (R T) - b_j / i ).
// One step of order 4 therefore costs ONE set of nine logarithms and nine exponentials and eight 9x9 mat-vecs, where the
// Bogacki-Shampine step costs three full right-hand sides (27 + 27 transcendental evaluations, six mat-vecs) -- ~2 200 instead of
// ~2 950 instructions -- and the step is of order 4 instead of 3: on the LHS sweep it needs 7 % fewer steps at the same
// tolerances with smaller outlet errors (CPU prototype tools/proto/taylor_proto.py).  The last term y_[4] h^4 is
// the error estimate (it is the local error of the order-3 polynomial; the order-4 polynomial is what is propagated -- the same
// relation as between the two members of the Bogacki-Shampine pair), exponent 1/4 in the controller.
//
// Kinks.  ln clip(y, lb, ub) has a kink where a species crosses the lower clamp (every product species does, once, right after
// the inlet: they start at 0 < lb = 1e-6).  A Runge-Kutta method samples f on both sides of such a kink; a Taylor polynomial
// built at the start of the step knows only one side.  The step polynomial is therefore checked for species that change sides,
// and such a step is cut at the crossing (Newton iteration on the quartic of the earliest crossing species): the expansion is
// exact on each side.  At most TAYLOR_MAX_EVENTS cuts per trajectory (a species that chatters around the clamp must not stall
// the integration); after that the kink is crossed like any explicit method crosses it.  An exponent at its clamp [-30, 30]
// has a constant rate (z_[i>0] = 0); a right-hand side at its output clamp (+-1e5: never seen with a trained model) is handed to
// the Rosenbrock kernel (PFR_ST_STIFF).
//
// Data.  y_[0], the mat-vec accumulators and one coefficient row live in registers; y_[1..3], 1/Y, r_[0], z_[1], z_[2]
// (63 values per thread) in shared memory, [vector][entry][thread] (conflict-free); r_[1], r_[2], L_[1], L_[2] are recomputed from those
// where they are needed (one or two multiplications) instead of being stored: 64.5 KB per CTA of 128 threads, three CTAs per SM.
#pragma once
#include "integrate_explicit.cuh"

namespace pfr {

constexpr int TAYLOR_BLOCK = 128;
#ifndef PFR_TAYLOR_MINB
#define PFR_TAYLOR_MINB 3
#endif
constexpr int TAYLOR_CTAS_PER_SM = PFR_TAYLOR_MINB;
constexpr int TAYLOR_VECS = 7;          // y1 y2 y3 1/Y r0 z1 z2
constexpr int TAYLOR_MAX_EVENTS = 48;
template <typename real> constexpr size_t taylor_smem_bytes() { return (size_t)TAYLOR_VECS * NS * TAYLOR_BLOCK * sizeof(real); }

template <typename real> __device__ __forceinline__ real below(real x);
template <> __device__ __forceinline__ double below<double>(double x) { return __longlong_as_double(__double_as_longlong(x) - 1); }   // x > 0
template <> __device__ __forceinline__ float below<float>(float x) { return __int_as_float(__float_as_int(x) - 1); }

template <typename real, bool kRamp>
__global__ void __launch_bounds__(TAYLOR_BLOCK, PFR_TAYLOR_MINB)
taylor4_kernel(const __grid_constant__ CrnnParams<real> p, const RodasArgs a) {
    __shared__ __align__(16) TpcCoef<real> sc;
    load_tpc_coef<real, TAYLOR_BLOCK>(sc, p, a.tables);
    const int zthr = bound_key(m_min(m_abs(p.zlo), m_abs(p.zhi))), dthr = bound_key(m_min(m_abs(p.dulo), m_abs(p.duhi)));
    const int lane = threadIdx.x & 31;
    const size_t n = (size_t)a.n;
    real* __restrict__ y_out = static_cast<real*>(a.y_out);
    real* __restrict__ y_dense = static_cast<real*>(a.y_dense);
    const bool dense = y_dense != nullptr;
    const bool raw = (a.flags & 1) != 0;
    const real rtol = real(a.rtol), atol = real(a.atol);

    extern __shared__ __align__(16) unsigned char ty_dyn[];
    real* const stv = reinterpret_cast<real*>(ty_dyn) + threadIdx.x;
#define TY_Y1(k) stv[(0 * NS + (k)) * TAYLOR_BLOCK]
#define TY_Y2(k) stv[(1 * NS + (k)) * TAYLOR_BLOCK]
#define TY_Y3(k) stv[(2 * NS + (k)) * TAYLOR_BLOCK]
#define TY_Q(k) stv[(3 * NS + (k)) * TAYLOR_BLOCK]
#define TY_R0(j) stv[(4 * NS + (j)) * TAYLOR_BLOCK]
#define TY_Z1(j) stv[(5 * NS + (j)) * TAYLOR_BLOCK]
#define TY_Z2(j) stv[(6 * NS + (j)) * TAYLOR_BLOCK]

    // lane-level work queue and knot bookkeeping: as in bs23_kernel
    bool have = false, fresh = false, exhausted = false;
    int i = 0, kend = 0, kc = 0, nacc = 0, nrej = 0, nrhs = 0, status = 0, stiff_cap = 0, nevents = 0;
    double t = 0.0, t_final = 0.0, tk = 0.0, tk1 = 0.0, hprop = 0.0;
    real Tk = real(0), Tk1 = real(0), slope = real(0);
    float t_ahead = 0.f, T_ahead = 0.f;
    real y[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) y[k] = real(0);

    while (true) {
        const unsigned want = __ballot_sync(0xffffffffu, !have && !exhausted);
        if (want) {
            int base = 0;
            const int leader = __ffs(want) - 1;
            if (lane == leader) base = atomicAdd(a.work_counter, __popc(want));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!have && !exhausted) {
                const int slot = base + __popc(want & ((1u << lane) - 1u));
                if (slot >= a.n) {
                    exhausted = true;
                } else {
                    i = a.perm ? a.perm[slot] : slot;
                    kend = a.idx_end ? a.idx_end[i] : NTOT - 1;
#pragma unroll
                    for (int k = 0; k < NS; k++) y[k] = real(0);
                    y[NS - 3] = real(a.c0[i]);
                    t = (double)a.tgrid[i];
                    t_final = (double)a.tgrid[(size_t)kend * n + i];
                    kc = 0;
                    tk = t;
                    tk1 = (double)a.tgrid[n + i];
                    Tk = Tk1 = real(a.T0[i]);
                    slope = real(0);
                    if (kRamp) {
                        Tk = real(a.Tprof[i]);
                        Tk1 = real(a.Tprof[n + i]);
                        slope = (Tk1 - Tk) / real(tk1 - tk);
                    }
                    const size_t k2i = NTOT > 2 ? 2 : NTOT - 1;
                    t_ahead = a.tgrid[k2i * n + i];
                    T_ahead = kRamp ? a.Tprof[k2i * n + i] : 0.f;
                    nacc = nrej = nrhs = status = nevents = 0;
                    stiff_cap = 8 * kend + 512;
                    hprop = 0.0;
                    if (dense) {
#pragma unroll
                        for (int k = 0; k < NS; k++) y_dense[(size_t)k * n + i] = raw ? y[k] : m_min(m_max(y[k], p.lb), p.ub);
                    }
                    have = true;
                    fresh = (kend != 0) && (t_final > t);
                    if (!fresh) kc = -1;   // nothing to integrate: falls through to the output code below
                }
            }
        }
        if (__all_sync(0xffffffffu, !have)) break;
        bool done = have && !fresh && kc < 0;
        if (have && !done) {
            // ---- Taylor coefficients y_[1..4] at (t, y) -------------------------------------------------------------------
            const real tau = real(t - tk);
            const real Tt = kRamp ? fma(slope, tau, Tk) : Tk;
            const real invT = rcp_full(Tt);
            const real mE = -p.inv_R * invT;
            const real lnT = t_log<real>(Tt, sc.ft);
            const real rho = kRamp ? -slope * invT : real(0);
            real z[NR], acc[NS];
            unsigned zfree = 0x1ffu;   // bit j: exponent j is strictly inside its clamp
#pragma unroll
            for (int j = 0; j < NR; j++) {
                real Ea, b, lnA, pad;
                lds2(&sc.arr[j][0], Ea, b, lnT);
                lds2(&sc.arr[j][2], lnA, pad, lnT);
                z[j] = fma(Ea, mE, fma(b, lnT, lnA));
            }
            // order 0: L_[0] = ln Y, z_[0], r_[0] = e^z, y_[1] = W r_[0]
#pragma unroll
            for (int k = 0; k < NS; k++) {
                const real Y = m_min(m_max(y[k], p.lb), p.ub);
                const real l = t_log<real>(Y, sc.ft);
                TY_Q(k) = (y[k] >= p.lb && y[k] <= p.ub) ? rcp_norm(Y) : real(0);
                real c[10];
                lds9(sc.nu[k], c, l);
#pragma unroll
                for (int j = 0; j < NR; j++) z[j] = fma(c[j], l, z[j]);
            }
            bool near = false;
#pragma unroll
            for (int j = 0; j < NR; j++) near = near || maybe_outside(z[j], zthr);
            if (near) {
#pragma unroll
                for (int j = 0; j < NR; j++) {
                    if (!(z[j] > p.zlo && z[j] < p.zhi)) zfree &= ~(1u << j);
                    z[j] = m_min(m_max(z[j], p.zlo), p.zhi);
                }
            }
#pragma unroll
            for (int k = 0; k < NS; k++) acc[k] = real(0);
#pragma unroll
            for (int j = 0; j < NR; j++) {
                const real r = t_exp<real>(z[j], sc.ft);
                TY_R0(j) = r;
                real c[10];
                lds9(sc.woutT[j], c, r);
#pragma unroll
                for (int k = 0; k < NS; k++) acc[k] = fma(c[k], r, acc[k]);
            }
            bool clamped_out = false;
#pragma unroll
            for (int k = 0; k < NS; k++) clamped_out = clamped_out || maybe_outside(acc[k], dthr);
            nrhs++;
            // order 1: L_[1] = y_[1] / Y, z_[1], r_[1] = z_[1] r_[0], y_[2] = W r_[1] / 2
#pragma unroll
            for (int j = 0; j < NR; j++) {
                real Ea, b, after = acc[0];
                lds2(&sc.arr[j][0], Ea, b, after);
                z[j] = rho * fma(Ea, mE, -b);
            }
#pragma unroll
            for (int k = 0; k < NS; k++) {
                TY_Y1(k) = acc[k];
                const real l1 = TY_Q(k) * acc[k];
                real c[10];
                lds9(sc.nu[k], c, l1);
#pragma unroll
                for (int j = 0; j < NR; j++) z[j] = fma(c[j], l1, z[j]);
            }
            if (zfree != 0x1ffu) {
#pragma unroll
                for (int j = 0; j < NR; j++) z[j] = ((zfree >> j) & 1u) ? z[j] : real(0);
            }
#pragma unroll
            for (int k = 0; k < NS; k++) acc[k] = real(0);
#pragma unroll
            for (int j = 0; j < NR; j++) {
                const real r1h = real(0.5) * z[j] * TY_R0(j);      // r_[1] / 2
                TY_Z1(j) = z[j];
                real c[10];
                lds9(sc.woutT[j], c, r1h);
#pragma unroll
                for (int k = 0; k < NS; k++) acc[k] = fma(c[k], r1h, acc[k]);
            }
            // order 2: L_[2] = (y_[2] - L_[1] y_[1] / 2) / Y, z_[2], r_[2] = z_[1] r_[1] / 2 + z_[2] r_[0], y_[3] = W r_[2] / 3
            const real rho2 = rho * rho;
#pragma unroll
            for (int j = 0; j < NR; j++) {
                real Ea, b, after = acc[0];
                lds2(&sc.arr[j][0], Ea, b, after);
                z[j] = rho2 * fma(Ea, mE, real(-0.5) * b);
            }
#pragma unroll
            for (int k = 0; k < NS; k++) {
                const real y1k = TY_Y1(k), qk = TY_Q(k);
                const real l1 = qk * y1k;
                const real l2 = qk * fma(real(-0.5) * l1, y1k, acc[k]);
                TY_Y2(k) = acc[k];
                real c[10];
                lds9(sc.nu[k], c, l2);
#pragma unroll
                for (int j = 0; j < NR; j++) z[j] = fma(c[j], l2, z[j]);
            }
            if (zfree != 0x1ffu) {
#pragma unroll
                for (int j = 0; j < NR; j++) z[j] = ((zfree >> j) & 1u) ? z[j] : real(0);
            }
#pragma unroll
            for (int k = 0; k < NS; k++) acc[k] = real(0);
#pragma unroll
            for (int j = 0; j < NR; j++) {
                const real r0 = TY_R0(j), z1 = TY_Z1(j);
                const real r1 = z1 * r0;
                const real r2t = real(1.0 / 3.0) * fma(real(0.5) * z1, r1, z[j] * r0);   // r_[2] / 3
                TY_Z2(j) = z[j];
                real c[10];
                lds9(sc.woutT[j], c, r2t);
#pragma unroll
                for (int k = 0; k < NS; k++) acc[k] = fma(c[k], r2t, acc[k]);
            }
            // order 3: L_[3] = (y_[3] - L_[1] y_[2] / 3 - 2 L_[2] y_[1] / 3) / Y, z_[3],
            //          r_[3] = z_[1] r_[2] / 3 + 2 z_[2] r_[1] / 3 + z_[3] r_[0], y_[4] = W r_[3] / 4
            const real rho3 = rho2 * rho;
#pragma unroll
            for (int j = 0; j < NR; j++) {
                real Ea, b, after = acc[0];
                lds2(&sc.arr[j][0], Ea, b, after);
                z[j] = rho3 * fma(Ea, mE, real(-1.0 / 3.0) * b);
            }
#pragma unroll
            for (int k = 0; k < NS; k++) {
                const real y1k = TY_Y1(k), y2k = TY_Y2(k), qk = TY_Q(k);
                const real l1 = qk * y1k;
                const real l2k = qk * fma(real(-0.5) * l1, y1k, y2k);
                const real l3 = qk * fma(real(-2.0 / 3.0) * l2k, y1k, fma(real(-1.0 / 3.0) * l1, y2k, acc[k]));
                TY_Y3(k) = acc[k];
                real c[10];
                lds9(sc.nu[k], c, l3);
#pragma unroll
                for (int j = 0; j < NR; j++) z[j] = fma(c[j], l3, z[j]);
            }
            if (zfree != 0x1ffu) {
#pragma unroll
                for (int j = 0; j < NR; j++) z[j] = ((zfree >> j) & 1u) ? z[j] : real(0);
            }
#pragma unroll
            for (int k = 0; k < NS; k++) acc[k] = real(0);
#pragma unroll
            for (int j = 0; j < NR; j++) {
                const real r0 = TY_R0(j), z1 = TY_Z1(j), z2 = TY_Z2(j);
                const real r1 = z1 * r0;
                const real r2 = fma(real(0.5) * z1, r1, z2 * r0);
                const real r3q = real(0.25) * fma(real(1.0 / 3.0) * z1, r2, fma(real(2.0 / 3.0) * z2, r1, z[j] * r0));   // r_[3] / 4
                real c[10];
                lds9(sc.woutT[j], c, r3q);
#pragma unroll
                for (int k = 0; k < NS; k++) acc[k] = fma(c[k], r3q, acc[k]);
            }   // acc = y_[4]

            // ---- the step ---------------------------------------------------------------------------------------------------
            if (fresh) {   // Hairer-style first step from |y0| and |f0|
                real d0 = real(0), d1 = real(0);
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    const real isk = rcp_norm(atol + rtol * m_abs(y[k]));
                    const real f0 = TY_Y1(k);
                    d0 = fma(y[k] * isk, y[k] * isk, d0);
                    d1 = fma(f0 * isk, f0 * isk, d1);
                }
                d0 = m_sqrt<real>(d0 / real(NS));
                d1 = m_sqrt<real>(d1 / real(NS));
                const double h0 = (d0 < real(1e-5) || d1 < real(1e-5)) ? 1e-6 : 0.01 * (double)d0 / (double)d1;
                hprop = fmin(100.0 * h0, t_final - t);
                fresh = false;
            }
            const double dist = tk1 - t;
            bool clip = hprop * 1.01 >= dist;
            double hs = clip ? dist : hprop;
            bool event = false;
            int kb = -1;          // the species whose crossing of the lower clamp cuts this step
            real w[NS];
            real err = real(0);
            bool finite = true;
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                const real h = real(hs);
                const real h4 = (h * h) * (h * h);
                real e2 = real(0);
                bool crossed = false;
                finite = true;
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    w[k] = fma(h, fma(h, fma(h, fma(h, acc[k], TY_Y3(k)), TY_Y2(k)), TY_Y1(k)), y[k]);
                    const real ek = acc[k] * h4;
                    const real isk = rcp_norm(atol + rtol * m_max(m_abs(y[k]), m_abs(w[k])));
                    e2 = fma(ek * isk, ek * isk, e2);
                    finite = finite && (m_abs(w[k]) < real(1e30));
                    crossed = crossed || ((y[k] < p.lb) != (w[k] < p.lb));
                }
                err = m_sqrt<real>(e2 / real(NS));
                if (pass == 1 || !crossed || !finite || nevents >= TAYLOR_MAX_EVENTS) break;
                // a species changes sides of the lower clamp during this step: cut the step at the earliest crossing
                real best = h;
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    if ((y[k] < p.lb) != (w[k] < p.lb)) {
                        const real tl = h * (p.lb - y[k]) / (w[k] - y[k]);   // chord
                        if (tl <= best) { best = tl; kb = k; }
                    }
                }
                real c0 = real(0), c1 = real(0), c2 = real(0), c3 = real(0), c4 = real(0);
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    if (k == kb) { c0 = y[k] - p.lb; c1 = TY_Y1(k); c2 = TY_Y2(k); c3 = TY_Y3(k); c4 = acc[k]; }
                }
                real tc = best;
#pragma unroll 1
                for (int it = 0; it < 4; it++) {   // Newton on the quartic, kept inside (0, h]
                    const real P = fma(tc, fma(tc, fma(tc, fma(tc, c4, c3), c2), c1), c0);
                    const real dP = fma(tc, fma(tc, fma(tc, real(4) * c4, real(3) * c3), real(2) * c2), c1);
                    const real tn = tc - P / dP;
                    tc = (tn > real(0) && tn <= h) ? tn : tc;
                }
                hs = (double)tc;
                if (!(hs > 0.0)) hs = (double)h * 0x1p-20;   // (a crossing at the very start of the step: move on by a sliver)
                clip = false;
                event = true;
                nevents++;
            }
            if (event) {   // whichever way the rounding falls, the crossing species ends this step ON the new side
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    if (k == kb) w[k] = (y[k] < p.lb) ? m_max(w[k], p.lb) : m_min(w[k], below<real>(p.lb));
                }
            }
            finite = finite && (err == err) && (err < real(1e30));
            // 0.9 err^(-1/4)
            const float fac = 0.9f * rsqrtf(sqrtf(fmaxf((float)err, 1e-30f)));
            if (clamped_out) {
                status = PFR_ST_STIFF_;   // right-hand side at its output clamp: not a case for a series expansion
                done = true;
            } else if (finite && err <= real(1)) {
                const double f = fmin(6.0, fmax(0.2, (double)fac));
                hprop = (clip || event) ? fmax(hprop, hs * f) : hs * f;   // a step cut short by a knot or a kink does not shrink the proposal
                nacc++;
#pragma unroll
                for (int k = 0; k < NS; k++) y[k] = w[k];
                if (clip) {
                    t = tk1;
                    kc++;
                    if (dense) {
#pragma unroll
                        for (int k = 0; k < NS; k++) y_dense[((size_t)kc * NS + k) * n + i] = raw ? y[k] : m_min(m_max(y[k], p.lb), p.ub);
                    }
                    if (kc >= kend) {
                        done = true;
                    } else {
                        tk = tk1;
                        tk1 = (double)t_ahead;
                        const size_t kk = (size_t)(kc + 2 < NTOT ? kc + 2 : NTOT - 1);
                        t_ahead = a.tgrid[kk * n + i];
                        if (kRamp) {
                            Tk = Tk1;
                            Tk1 = real(T_ahead);
                            slope = (Tk1 - Tk) / real(tk1 - tk);
                            T_ahead = a.Tprof[kk * n + i];
                        }
                    }
                } else {
                    t += hs;
                }
            } else {
                nrej++;
                const double f = finite ? fmax(0.2, (double)fac) : 0.2;
                hprop = hs * fmin(f, 0.9);
                if (!(t + hprop > t) || hprop < 1e-300) { status = finite ? PFR_ST_UNDERFLOW_ : PFR_ST_NONFINITE_; done = true; }
            }
            if (!done && nacc + nrej > stiff_cap) { status = PFR_ST_STIFF_; done = true; }
            if (!done && nacc + nrej >= a.max_steps) { status = PFR_ST_MAXSTEPS_; done = true; }
        }
        if (done) {
            if (kc < 0) kc = 0;
            const int io = a.out_index ? a.out_index[i] : i;   // column of the results (caller's order; pfr_sweep_run)
            real yf[NS];
#pragma unroll
            for (int k = 0; k < NS; k++) {
                yf[k] = m_min(m_max(y[k], p.lb), p.ub);
                y_out[(size_t)k * n + io] = yf[k];
            }
            a.status[io] = status;
            if (a.stats) {
                a.stats[io] = nacc;
                a.stats[n + io] = nrej;
                a.stats[2 * n + io] = nrhs;
            }
            if (dense && kc < NTOT - 1) {
                for (int kk = kc + 1; kk < NTOT; kk++)
#pragma unroll
                    for (int k = 0; k < NS; k++) y_dense[((size_t)kk * NS + k) * n + i] = raw ? y[k] : yf[k];
            }
            have = false;
        }
    }
}
#undef TY_Y1
#undef TY_Y2
#undef TY_Y3
#undef TY_Q
#undef TY_R0
#undef TY_Z1
#undef TY_Z2

}  // namespace pfr
