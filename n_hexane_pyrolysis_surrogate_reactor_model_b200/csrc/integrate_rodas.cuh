// Adaptive Rosenbrock integrator (RODAS4, Hairer & Wanner: 6 stages, order 4(3), stiffly accurate),
// one PFR condition per thread.
//
// Replaces, for a whole batch at once, the reference's per-condition
//   torchdiffeq.odeint(CRNNFunc, u0, t, 'dopri5', atol, rtol)
//   (SURROGATE_MODEL/surrogate_model_Eoff_single_model.py:175-186, ...Eon_single_model.py:153-156).
//
// Layout: all per-condition arrays are SoA with the condition index fastest ([knot][n], [species][n]) so
// that a warp's loads and stores are contiguous.  The 9x9 matrix E = I/(h*gamma) - J is built and
// factored (LU, no pivoting: E is diagonally dominated by 1/(h*gamma) and J_kk <= 0 by construction of
// the CRNN) in registers, then parked in shared memory ([entry][thread], conflict-free)
// for the six triangular solves of the step.
//
// Temperature: constant (Eoff) or the piecewise-linear MLP profile (Eon).  When a profile or dense
// output is requested the stepper never crosses a knot of the time grid: the right-hand side is smooth
// inside every knot interval, which is what makes a tight match with the converged solution possible
// (the reference's dopri5 steps across the kinks and pays with rejected steps and ~1e-4 errors).
#pragma once
#include "crnn_device.cuh"
#include "fastmath.cuh"

namespace pfr {

constexpr int RODAS_BLOCK = 128;
constexpr int PFR_ST_MAXSTEPS_ = 1, PFR_ST_NONFINITE_ = 2, PFR_ST_UNDERFLOW_ = 3;

struct RodasArgs {
    int n;
    const float* T0;       // [n] inlet temperature (K)
    const float* c0;       // [n] inlet n-hexane concentration (mol/m3); all other species start at 0
    const float* tgrid;    // [801][n] strictly increasing knots, or nullptr (then t_end must be given)
    const float* Tprof;    // [801][n] temperature at the knots, or nullptr: T(t) = T0
    const float* t_end;    // [n] final time when tgrid == nullptr
    const int* idx_end;    // [n] last knot to integrate to (nullptr: 800)
    const int* perm;       // [n] thread j integrates condition perm[j] (cost-sorted order), or nullptr
    double rtol, atol;
    void* y_out;           // [9][n] real, clamped to [lb, ub]
    void* y_dense;         // [801][9][n] real clamped, or nullptr
    int* status;           // [n] 0 ok, 1 max steps, 2 non-finite, 3 step underflow
    int* stats;            // [3][n] accepted, rejected, rhs evaluations; or nullptr
    int max_steps;
    const FastTables* tables;  // device copy of the log / exp tables (fastmath.cuh)
    int flags;                 // bit 0: y_dense holds the raw (unclamped) knot states
    int* work_counter = nullptr;  // bs23_kernel: zero-initialised device counter of its lane-level work queue
    // Sweep pipeline (pfr_sweep_run): the per-condition INPUT arrays above are held in the pipeline's visiting order, results go
    // back to the caller's order: y_out / status / stats of work item i are written at column out_index[i] (nullptr: i).
    const int* out_index = nullptr;
    // Number of work items taken from device memory (the stiff-fallback launch: its list is built on the device and nobody
    // waits for the host to learn its length); nullptr: n.  Array strides stay n.
    const int* n_work = nullptr;
};

namespace rodas4 {
constexpr double gamma = 0.25;
constexpr double a21 = 0.1544000000000000e+01;
constexpr double a31 = 0.9466785280815826e+00, a32 = 0.2557011698983284e+00;
constexpr double a41 = 0.3314825187068521e+01, a42 = 0.2896124015972201e+01, a43 = 0.9986419139977817e+00;
constexpr double a51 = 0.1221224509226641e+01, a52 = 0.6019134481288629e+01, a53 = 0.1253708332932087e+02,
                 a54 = -0.6878860361058950e+00;
constexpr double C21 = -0.5668800000000000e+01;
constexpr double C31 = -0.2430093356833875e+01, C32 = -0.2063599157091915e+00;
constexpr double C41 = -0.1073529058151375e+00, C42 = -0.9594562251023355e+01, C43 = -0.2047028614809616e+02;
constexpr double C51 = 0.7496443313967647e+01, C52 = -0.1024680431464352e+02, C53 = -0.3399990352819905e+02,
                 C54 = 0.1170890893206160e+02;
constexpr double C61 = 0.8083246795921522e+01, C62 = -0.7981132988064893e+01, C63 = -0.3152159432874371e+02,
                 C64 = 0.1631930543123136e+02, C65 = -0.6058818238834054e+01;
constexpr double c2 = 0.386, c3 = 0.21, c4 = 0.63;
constexpr double d1 = 0.25, d2 = -0.1043, d3 = 0.1035, d4 = -0.3620000000000023e-01;
}  // namespace rodas4

// ---- 9x9 LU without pivoting, in registers; reciprocal pivots stored on the diagonal -------------
template <typename real>
__device__ __forceinline__ bool lu_factor(real (&A)[NS * NS]) {
    bool ok = true;
#pragma unroll
    for (int c = 0; c < NS; c++) {
        const real piv = A[c * NS + c];
        ok = ok && (m_abs(piv) > real(1e-30));
        const real ip = real(1) / piv;
        A[c * NS + c] = ip;
#pragma unroll
        for (int r = c + 1; r < NS; r++) {
            const real l = A[r * NS + c] * ip;
            A[r * NS + c] = l;
#pragma unroll
            for (int cc = c + 1; cc < NS; cc++) A[r * NS + cc] = fma(-l, A[c * NS + cc], A[r * NS + cc]);
        }
    }
    return ok;
}

// Per-thread shared-memory scratch, entry e of thread tid at sm[e * RODAS_BLOCK + tid] (conflict-free):
//   [0, 81)  LU factors of E (q_k, g_j park in its head while the Jacobian is formed)
//   [81, 90) y at the step start      [90, 99) work vector: stage argument, then its exponents z_j
//   [99,108) temperature ramp: h-independent df/dt; isothermal: the constant exponent part kT_j
constexpr int SM_LU = 0, SM_Y = NS * NS, SM_W = SM_Y + NS, SM_FX = SM_W + NS, SM_KT = SM_FX;
template <bool kRamp> constexpr int sm_entries() { return SM_FX + NS; }
#ifndef PFR_RHS_UNROLL
#define PFR_RHS_UNROLL 3
#endif

// x <- E^{-1} x with the factors parked in shared memory
template <typename real>
__device__ __forceinline__ void lu_solve(const real* __restrict__ sm, real (&x)[NS]) {
#pragma unroll
    for (int i = 1; i < NS; i++) {
#pragma unroll
        for (int k = 0; k < i; k++) x[i] = fma(-sm[(SM_LU + i * NS + k) * RODAS_BLOCK], x[k], x[i]);
    }
#pragma unroll
    for (int i = NS - 1; i >= 0; i--) {
#pragma unroll
        for (int k = i + 1; k < NS; k++) x[i] = fma(-sm[(SM_LU + i * NS + k) * RODAS_BLOCK], x[k], x[i]);
        x[i] *= sm[(SM_LU + i * NS + i) * RODAS_BLOCK];
    }
}

// kRamp : T(t) follows Tprof (piecewise linear); otherwise T = T0
// kKnots: steps stop at every knot of tgrid (required by kRamp and by dense output)
template <typename real, bool kRamp, bool kKnots>
__global__ void __launch_bounds__(RODAS_BLOCK, 2)
rodas4_kernel(const __grid_constant__ CrnnParams<real> p, const RodasArgs a) {
    using namespace rodas4;
    static_assert(!kRamp || kKnots, "a temperature ramp needs knot-limited stepping");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real* sm = reinterpret_cast<real*>(smem_raw) + threadIdx.x;

    const int slot = blockIdx.x * RODAS_BLOCK + threadIdx.x;
    if (slot >= a.n) return;
    const int i = a.perm ? a.perm[slot] : slot;
    const size_t n = (size_t)a.n;
    real* __restrict__ y_out = static_cast<real*>(a.y_out);
    real* __restrict__ y_dense = static_cast<real*>(a.y_dense);
    const bool dense = kKnots && (y_dense != nullptr);

#pragma unroll
    for (int k = 0; k < NS; k++) sm[(SM_Y + k) * RODAS_BLOCK] = (k == NS - 3) ? real(a.c0[i]) : real(0);

    const int kend = (kKnots && a.idx_end) ? a.idx_end[i] : NTOT - 1;
    double t = kKnots ? (double)a.tgrid[i] : 0.0;
    const double t_final = kKnots ? (double)a.tgrid[(size_t)kend * n + i]
                                  : (a.t_end ? (double)a.t_end[i] : (double)a.tgrid[(size_t)(NTOT - 1) * n + i]);
    if (dense) {
#pragma unroll
        for (int k = 0; k < NS; k++)
            y_dense[(size_t)k * n + i] = m_min(m_max(sm[(SM_Y + k) * RODAS_BLOCK], p.lb), p.ub);
    }

    // current knot interval [tk, tk1] and its temperature ramp
    int kc = 0;
    double tk = t, tk1 = kKnots ? (double)a.tgrid[n + i] : t_final;
    real Tk = real(a.T0[i]), slope = real(0);
    if (kRamp) {
        Tk = real(a.Tprof[i]);
        slope = (real(a.Tprof[n + i]) - Tk) / real(tk1 - tk);
    }

    if (!kRamp) {
        real kT0[NR], dk0[NR];
        arrhenius_T<real, false>(p, Tk, kT0, dk0);
#pragma unroll
        for (int j = 0; j < NR; j++) sm[(SM_KT + j) * RODAS_BLOCK] = kT0[j];
    }
    // exponent part that depends on T only: recomputed per stage on a ramp, re-read from shared memory otherwise
    auto load_kT = [&](double tq, real (&kT)[NR]) {
        if (kRamp) {
            real dk[NR];
            arrhenius_T<real, false>(p, Tk + slope * real(tq - tk), kT, dk);
        } else {
#pragma unroll
            for (int j = 0; j < NR; j++) kT[j] = sm[(SM_KT + j) * RODAS_BLOCK];
        }
    };

    const real rtol = real(a.rtol), atol = real(a.atol);
    int nacc = 0, nrej = 0, nrhs = 0, status = 0;
    double hprop = 0.0;
    bool first = true;
    bool done = (kend == 0) || !(t_final > t);

    while (!done) {
        asm volatile("" ::: "memory");  // y, kT, df/dt live in shared memory across iterations, not in registers
        real ak1[NS];  // starts as f0 (+ h d1 df/dt), becomes stage 1
        bool lu_ok;
        double hs;
        bool clip;
        {
            // ---------------- f0, df/dt and the Jacobian at (t, y) ----------------
            real dkT[NR];
            {
                real kT[NR];
                if (kRamp) arrhenius_T<real, true>(p, Tk + slope * real(t - tk), kT, dkT);
                else load_kT(t, kT);
                crnn_rhs_sm<real, RODAS_BLOCK, PFR_RHS_UNROLL, true>(p, kT, sm, SM_Y, SM_W, SM_LU, ak1);
            }
            nrhs++;
            if (first) {
                // Hairer's first guess without the second evaluation: h = 0.01 |y| / |f| in the error norm
                real d0 = real(0), d1n = real(0);
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    const real yk = sm[(SM_Y + k) * RODAS_BLOCK];
                    const real sk = atol + rtol * m_abs(yk);
                    d0 = fma(yk / sk, yk / sk, d0);
                    d1n = fma(ak1[k] / sk, ak1[k] / sk, d1n);
                }
                d0 = m_sqrt<real>(d0 / real(NS));
                d1n = m_sqrt<real>(d1n / real(NS));
                const double h0 = (d0 < real(1e-5) || d1n < real(1e-5)) ? 1e-6 : 0.01 * (double)d0 / (double)d1n;
                hprop = fmin(100.0 * h0, t_final - t);
                first = false;
            }
            // ---------------- step size: never cross the next stop ----------------
            const double dist = tk1 - t;
            clip = hprop * 1.01 >= dist;
            hs = clip ? dist : hprop;
            real g[NR];
#pragma unroll
            for (int j = 0; j < NR; j++) g[j] = sm[(SM_LU + NS + j) * RODAS_BLOCK];
            if (kRamp) {
                // df/dt = (df/dT) dT/dt ; stage i adds h d_i df/dt
                const real hd1 = real(hs) * real(d1);
#pragma unroll
                for (int r = 0; r < NS; r++) {
                    real s = real(0);
#pragma unroll
                    for (int j = 0; j < NR; j++) s = fma(p.wout[r][j], g[j] * dkT[j], s);
                    s *= slope;
                    sm[(SM_FX + r) * RODAS_BLOCK] = s;
                    ak1[r] = fma(hd1, s, ak1[r]);
                }
            }
            // f0 (+ h d1 df/dt) waits in the work vector while the matrix occupies the registers
#pragma unroll
            for (int k = 0; k < NS; k++) sm[(SM_W + k) * RODAS_BLOCK] = ak1[k];
            // J[r][k] = sum_j wout[r][j] g_j nu[k][j] q_k ; A = I/(h gamma) - J
            // (the +-1e5 clamp on du is not differentiated: it is never active on a physical trajectory)
            real A[NS * NS];
            const real fac = real(1.0 / gamma) / real(hs);
#pragma unroll
            for (int k = 0; k < NS; k++) {
                const real qk = sm[(SM_LU + k) * RODAS_BLOCK];
                real aj[NR];
#pragma unroll
                for (int j = 0; j < NR; j++) aj[j] = g[j] * p.nu[k][j] * qk;
#pragma unroll
                for (int r = 0; r < NS; r++) {
                    real s = (r == k) ? fac : real(0);
#pragma unroll
                    for (int j = 0; j < NR; j++) s = fma(-p.wout[r][j], aj[j], s);
                    A[r * NS + k] = s;
                }
            }
            lu_ok = lu_factor<real>(A);
#pragma unroll
            for (int e = 0; e < NS * NS; e++) sm[(SM_LU + e) * RODAS_BLOCK] = A[e];
        }
        // the factors must really leave the register file here: stop the compiler from forwarding the
        // stores above into the loads of the six triangular solves
        asm volatile("" ::: "memory");
        const real h = real(hs);
        const real ih = real(1) / h;

        // ---------------- six stages ----------------
        real ak2[NS], ak3[NS], ak4[NS], ak5[NS], dy[NS], kT[NR];
#pragma unroll
        for (int k = 0; k < NS; k++) ak1[k] = sm[(SM_W + k) * RODAS_BLOCK];
        lu_solve<real>(sm, ak1);

#pragma unroll
        for (int k = 0; k < NS; k++) sm[(SM_W + k) * RODAS_BLOCK] = fma(real(a21), ak1[k], sm[(SM_Y + k) * RODAS_BLOCK]);
        load_kT(t + c2 * hs, kT);
        crnn_rhs_sm<real, RODAS_BLOCK, PFR_RHS_UNROLL, false>(p, kT, sm, SM_W, SM_W, SM_LU, dy);
#pragma unroll
        for (int k = 0; k < NS; k++) {
            const real s = fma(real(C21) * ih, ak1[k], dy[k]);
            ak2[k] = kRamp ? fma(h * real(d2), sm[(SM_FX + k) * RODAS_BLOCK], s) : s;
        }
        lu_solve<real>(sm, ak2);

#pragma unroll
        for (int k = 0; k < NS; k++)
            sm[(SM_W + k) * RODAS_BLOCK] = fma(real(a32), ak2[k], fma(real(a31), ak1[k], sm[(SM_Y + k) * RODAS_BLOCK]));
        load_kT(t + c3 * hs, kT);
        crnn_rhs_sm<real, RODAS_BLOCK, PFR_RHS_UNROLL, false>(p, kT, sm, SM_W, SM_W, SM_LU, dy);
#pragma unroll
        for (int k = 0; k < NS; k++) {
            const real s = fma(real(C31) * ih, ak1[k], fma(real(C32) * ih, ak2[k], dy[k]));
            ak3[k] = kRamp ? fma(h * real(d3), sm[(SM_FX + k) * RODAS_BLOCK], s) : s;
        }
        lu_solve<real>(sm, ak3);

#pragma unroll
        for (int k = 0; k < NS; k++)
            sm[(SM_W + k) * RODAS_BLOCK] =
                fma(real(a43), ak3[k], fma(real(a42), ak2[k], fma(real(a41), ak1[k], sm[(SM_Y + k) * RODAS_BLOCK])));
        load_kT(t + c4 * hs, kT);
        crnn_rhs_sm<real, RODAS_BLOCK, PFR_RHS_UNROLL, false>(p, kT, sm, SM_W, SM_W, SM_LU, dy);
#pragma unroll
        for (int k = 0; k < NS; k++) {
            const real s = fma(real(C41) * ih, ak1[k], fma(real(C42) * ih, ak2[k], fma(real(C43) * ih, ak3[k], dy[k])));
            ak4[k] = kRamp ? fma(h * real(d4), sm[(SM_FX + k) * RODAS_BLOCK], s) : s;
        }
        lu_solve<real>(sm, ak4);

        // stage-5 argument; kept in registers as well because stages 5, 6 and the new state build on it
        real ynew[NS];
#pragma unroll
        for (int k = 0; k < NS; k++) {
            ynew[k] = fma(real(a54), ak4[k], fma(real(a53), ak3[k], fma(real(a52), ak2[k],
                      fma(real(a51), ak1[k], sm[(SM_Y + k) * RODAS_BLOCK]))));
            sm[(SM_W + k) * RODAS_BLOCK] = ynew[k];
        }
        load_kT(t + hs, kT);
        crnn_rhs_sm<real, RODAS_BLOCK, PFR_RHS_UNROLL, false>(p, kT, sm, SM_W, SM_W, SM_LU, dy);
#pragma unroll
        for (int k = 0; k < NS; k++)
            ak5[k] = fma(real(C51) * ih, ak1[k], fma(real(C52) * ih, ak2[k],
                     fma(real(C53) * ih, ak3[k], fma(real(C54) * ih, ak4[k], dy[k]))));
        lu_solve<real>(sm, ak5);

#pragma unroll
        for (int k = 0; k < NS; k++) {
            ynew[k] += ak5[k];  // embedded 3rd-order solution
            sm[(SM_W + k) * RODAS_BLOCK] = ynew[k];
        }
        crnn_rhs_sm<real, RODAS_BLOCK, PFR_RHS_UNROLL, false>(p, kT, sm, SM_W, SM_W, SM_LU, dy);
        real er[NS];
#pragma unroll
        for (int k = 0; k < NS; k++)
            er[k] = fma(real(C61) * ih, ak1[k], fma(real(C62) * ih, ak2[k], fma(real(C63) * ih, ak3[k],
                    fma(real(C64) * ih, ak4[k], fma(real(C65) * ih, ak5[k], dy[k])))));
        lu_solve<real>(sm, er);
        nrhs += 5;

        // ---------------- error estimate and step-size control ----------------
        real e2 = real(0);
        bool finite = lu_ok;
#pragma unroll
        for (int k = 0; k < NS; k++) {
            ynew[k] += er[k];
            const real sk = atol + rtol * m_max(m_abs(sm[(SM_Y + k) * RODAS_BLOCK]), m_abs(ynew[k]));
            const real w = er[k] / sk;
            e2 = fma(w, w, e2);
            finite = finite && (m_abs(ynew[k]) < real(1e30));
        }
        const real err = m_sqrt<real>(e2 / real(NS));
        finite = finite && (err == err) && (err < real(1e30));

        if (finite && err <= real(1)) {
            // accepted: hnew = h * min(6, max(0.2, 0.9 err^(-1/4)))
            double f = err > real(0) ? 0.9 / sqrt(sqrt((double)err)) : 6.0;
            f = fmin(6.0, fmax(0.2, f));
            hprop = clip ? fmax(hprop, hs * f) : hs * f;
            nacc++;
#pragma unroll
            for (int k = 0; k < NS; k++) sm[(SM_Y + k) * RODAS_BLOCK] = ynew[k];
            if (clip) {
                t = tk1;
                if (kKnots) {
                    kc++;
                    if (dense) {
#pragma unroll
                        for (int k = 0; k < NS; k++)
                            y_dense[((size_t)kc * NS + k) * n + i] = (a.flags & 1) ? ynew[k] : m_min(m_max(ynew[k], p.lb), p.ub);
                    }
                    if (kc >= kend) {
                        done = true;
                    } else {
                        tk = tk1;
                        tk1 = (double)a.tgrid[(size_t)(kc + 1) * n + i];
                        if (kRamp) {
                            Tk = real(a.Tprof[(size_t)kc * n + i]);
                            slope = (real(a.Tprof[(size_t)(kc + 1) * n + i]) - Tk) / real(tk1 - tk);
                        }
                    }
                } else {
                    done = true;
                }
            } else {
                t += hs;
            }
        } else {
            nrej++;
            const double f = finite ? fmax(0.2, 0.9 / sqrt(sqrt((double)err))) : 0.2;
            hprop = hs * fmin(f, 0.9);
            if (!(t + hprop > t) || hprop < 1e-300) { status = finite ? PFR_ST_UNDERFLOW_ : PFR_ST_NONFINITE_; done = true; }
        }
        if (!done && nacc + nrej >= a.max_steps) { status = PFR_ST_MAXSTEPS_; done = true; }
    }

    real yf[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) {
        yf[k] = m_min(m_max(sm[(SM_Y + k) * RODAS_BLOCK], p.lb), p.ub);
        y_out[(size_t)k * n + i] = yf[k];
    }
    a.status[i] = status;
    if (a.stats) {
        a.stats[i] = nacc;
        a.stats[n + i] = nrej;
        a.stats[2 * n + i] = nrhs;
    }
    if (dense && kc < NTOT - 1) {
        // rows past the last knot reached (idx_end < 800, or a failed trajectory) repeat the final state
        for (int kk = kc + 1; kk < NTOT; kk++)
#pragma unroll
            for (int k = 0; k < NS; k++) y_dense[((size_t)kk * NS + k) * n + i] = (a.flags & 1) ? sm[(SM_Y + k) * RODAS_BLOCK] : yf[k];
    }
}

}  // namespace pfr
