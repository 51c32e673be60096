// Reference-behaviour integrator: adaptive Dormand-Prince 5(4) with torchdiffeq 0.2.x semantics, one
// condition per thread.  This is the "same digits as the reference" mode of the drop-in:
//   torchdiffeq.odeint(CRNNFunc, u0, t, method='dopri5', atol, rtol)
//   (call sites SURROGATE_MODEL/surrogate_model_Eoff_single_model.py:185, ...Eon_single_model.py:154-155)
// restated from the package's published algorithm (not vendored by the reference):
// state, tableau and stage times in the state dtype `real`, step clock in double, Hairer initial step
// (order 4), RMS error ratio against atol + rtol*max(|y0|,|y1|), factor clip(0.9 ratio^-1/5, 0.2|1, 10),
// FSAL, alpha == 1 stages evaluated at nextafter(t1, -inf), no clipping at the final time (the last
// step overshoots and T(t) is linearly extrapolated past the last knot, exactly as the reference's
// linear_interpolation does), quartic dense output through (y0, y_mid, y1, f0, f1).
#pragma once
#include "crnn_device.cuh"

namespace pfr {

constexpr int DOPRI_BLOCK = 128;

struct Dopri5Args {
    int n;
    const float* T0;     // [n]
    const float* c0;     // [n]
    const float* tgrid;  // [801][n] (needed for a temperature profile or dense output), else nullptr
    const float* Tprof;  // [801][n] or nullptr: T(t) = T0
    const float* t_end;  // [n] output time when tgrid == nullptr
    const int* idx_end;  // [n] knot whose state is reported in y_out (nullptr: 800)
    const int* perm;     // [n] thread j integrates condition perm[j], or nullptr
    double rtol, atol;
    void* y_out;         // [9][n] real, clamped
    void* y_dense;       // [801][9][n] real, clamped, or nullptr
    int* status;         // [n] 0 ok, 1 max steps, 2 non-finite state, 3 dt underflow (torchdiffeq's asserts)
    int* stats;          // [3][n] accepted, rejected, rhs evaluations
    int max_steps;
};

template <typename real> __device__ __forceinline__ real prev_float(real x);
template <> __device__ __forceinline__ float prev_float<float>(float x) { return nextafterf(x, x - 1.0f); }
template <> __device__ __forceinline__ double prev_float<double>(double x) { return nextafter(x, x - 1.0); }
template <typename real> __device__ __forceinline__ real m_pow(real x, real y);
template <> __device__ __forceinline__ float m_pow<float>(float x, float y) { return powf(x, y); }
template <> __device__ __forceinline__ double m_pow<double>(double x, double y) { return pow(x, y); }

// linear_interpolation(ts, Ts)(t): idx = clamp(searchsorted(ts, t, right=True), 1, 800), extrapolating
// linearly outside the grid (...Eoff_single_model.py:106-115).  Column i of the [801][n] arrays.
template <typename real>
__device__ __forceinline__ real interp_T(const float* __restrict__ tg, const float* __restrict__ Tp, size_t n, int i, real t) {
    int lo = 0, hi = NTOT;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (real(tg[(size_t)mid * n + i]) > t) hi = mid; else lo = mid + 1;
    }
    const int idx = lo < 1 ? 1 : (lo > NTOT - 1 ? NTOT - 1 : lo);
    const real x0 = real(tg[(size_t)(idx - 1) * n + i]), x1 = real(tg[(size_t)idx * n + i]);
    const real y0 = real(Tp[(size_t)(idx - 1) * n + i]), y1 = real(Tp[(size_t)idx * n + i]);
    const real slope = (y1 - y0) / (x1 - x0);
    return y0 + slope * (t - x0);
}

template <typename real>
__device__ __forceinline__ real rms9(const real (&x)[NS]) {
    real s = real(0);
#pragma unroll
    for (int k = 0; k < NS; k++) s += x[k] * x[k];
    return m_sqrt<real>(s / real(NS));
}

template <typename real, bool kRamp>
__global__ void __launch_bounds__(DOPRI_BLOCK)
dopri5_kernel(const __grid_constant__ CrnnParams<real> p, const Dopri5Args a) {
    const int slot = blockIdx.x * DOPRI_BLOCK + threadIdx.x;
    if (slot >= a.n) return;
    const int i = a.perm ? a.perm[slot] : slot;
    const size_t n = (size_t)a.n;
    real* __restrict__ y_out = static_cast<real*>(a.y_out);
    real* __restrict__ y_dense = static_cast<real*>(a.y_dense);
    const bool kDense = y_dense != nullptr;

    // Butcher tableau in the state dtype
    const real al[6] = {real(1.0 / 5), real(3.0 / 10), real(4.0 / 5), real(8.0 / 9), real(1.0), real(1.0)};
    const real be[6][6] = {
        {real(1.0 / 5)},
        {real(3.0 / 40), real(9.0 / 40)},
        {real(44.0 / 45), real(-56.0 / 15), real(32.0 / 9)},
        {real(19372.0 / 6561), real(-25360.0 / 2187), real(64448.0 / 6561), real(-212.0 / 729)},
        {real(9017.0 / 3168), real(-355.0 / 33), real(46732.0 / 5247), real(49.0 / 176), real(-5103.0 / 18656)},
        {real(35.0 / 384), real(0), real(500.0 / 1113), real(125.0 / 192), real(-2187.0 / 6784), real(11.0 / 84)}};
    const real ce[7] = {real(35.0 / 384 - 1951.0 / 21600), real(0), real(500.0 / 1113 - 22642.0 / 50085),
                        real(125.0 / 192 - 451.0 / 720), real(-2187.0 / 6784 - -12231.0 / 42400),
                        real(11.0 / 84 - 649.0 / 6300), real(-1.0 / 60.0)};
    const real cm[7] = {real(6025192743.0 / 30085553152.0 / 2), real(0), real(51252292925.0 / 65400821598.0 / 2),
                        real(-2691868925.0 / 45128329728.0 / 2), real(187940372067.0 / 1594534317056.0 / 2),
                        real(-1776094331.0 / 19743644256.0 / 2), real(11237099.0 / 235043384.0 / 2)};

    const real rtol = real(a.rtol), atol = real(a.atol);
    const int kend = a.idx_end ? a.idx_end[i] : NTOT - 1;
    const real T0 = real(a.T0[i]);
    real kT[NR], dkT[NR], g_[NR], q_[NS], md_[NS];
    if (!kRamp) arrhenius_T<real, false>(p, T0, kT, dkT);
    int nacc = 0, nrej = 0, nrhs = 0, status = 0;

    auto rhs = [&](real tq, const real (&u)[NS], real (&du)[NS]) {
        if (kRamp) arrhenius_T<real, false>(p, interp_T<real>(a.tgrid, a.Tprof, n, i, tq), kT, dkT);
        crnn_rhs<real, false>(p, kT, u, du, g_, q_, md_);
        nrhs++;
    };

    real y[NS], f[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) y[k] = real(0);
    y[NS - 3] = real(a.c0[i]);
    const double t0d = a.tgrid ? (double)a.tgrid[i] : 0.0;

    // _select_initial_step(order = 4)
    rhs(real(t0d), y, f);
    double dt;
    {
        real sc[NS], tmp[NS], y1[NS], f1[NS];
#pragma unroll
        for (int k = 0; k < NS; k++) { sc[k] = atol + m_abs(y[k]) * rtol; tmp[k] = y[k] / sc[k]; }
        const real d0 = rms9<real>(tmp);
#pragma unroll
        for (int k = 0; k < NS; k++) tmp[k] = f[k] / sc[k];
        const real d1 = rms9<real>(tmp);
        real h0 = (d0 < real(1e-5) || d1 < real(1e-5)) ? real(1e-6) : real(0.01) * d0 / d1;
        h0 = m_abs(h0);
#pragma unroll
        for (int k = 0; k < NS; k++) y1[k] = y[k] + h0 * f[k];
        rhs(real(t0d + (double)h0), y1, f1);
#pragma unroll
        for (int k = 0; k < NS; k++) tmp[k] = (f1[k] - f[k]) / sc[k];
        const real d2 = m_abs(rms9<real>(tmp) / h0);
        real h1;
        if (d1 <= real(1e-15) && d2 <= real(1e-15)) h1 = m_max(real(1e-6), h0 * real(1e-3));
        else h1 = m_pow<real>(real(0.01) / m_max(d1, d2), real(1.0 / 5.0));
        h1 = m_abs(h1);
        dt = (double)m_min(real(100) * h0, h1);
    }

    double rk_t0 = t0d, rk_t1 = t0d;
    real ic[5][NS];  // dense-output polynomial e, d, c, b, a
#pragma unroll
    for (int m = 0; m < 5; m++)
#pragma unroll
        for (int k = 0; k < NS; k++) ic[m][k] = y[k];
    real yrep[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) yrep[k] = y[k];
    if (kDense) {
#pragma unroll
        for (int k = 0; k < NS; k++) y_dense[(size_t)k * n + i] = m_min(m_max(y[k], p.lb), p.ub);
    }

    int io = kDense ? 1 : kend;
    const int io_last = kDense ? NTOT - 1 : kend;
    for (; io <= io_last && io >= 1 && status == 0; io++) {
        const double next_t = a.tgrid ? (double)a.tgrid[(size_t)io * n + i] : (double)a.t_end[i];
        while (next_t > rk_t1) {
            const double ta = rk_t1, tb = ta + dt;
            if (!(ta + dt > ta)) { status = 3; break; }
            bool fin = true;
#pragma unroll
            for (int k = 0; k < NS; k++) fin = fin && (m_abs(y[k]) <= real(3.0e38)) ;
            if (!fin) { status = 2; break; }
            if (nacc + nrej >= a.max_steps) { status = 1; break; }
            const real t0s = real(ta), dts = real(dt), t1s = real(tb);
            real kk[7][NS], yi[NS];
#pragma unroll
            for (int k = 0; k < NS; k++) kk[0][k] = f[k];
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const real ti = (j >= 4) ? prev_float<real>(t1s) : t0s + al[j] * dts;
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    real s = real(0);
#pragma unroll
                    for (int m = 0; m <= j; m++) s += kk[m][k] * (be[j][m] * dts);
                    yi[k] = y[k] + s;
                }
                rhs(ti, yi, kk[j + 1]);
            }
            real er[NS];
#pragma unroll
            for (int k = 0; k < NS; k++) {
                real s = real(0);
#pragma unroll
                for (int m = 0; m < 7; m++) s += kk[m][k] * (dts * ce[m]);
                er[k] = s / (atol + rtol * m_max(m_abs(y[k]), m_abs(yi[k])));
            }
            const real ratio = m_abs(rms9<real>(er));
            if (ratio <= real(1)) {
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    real s = real(0);
#pragma unroll
                    for (int m = 0; m < 7; m++) s += kk[m][k] * (dts * cm[m]);
                    const real ymid = y[k] + s, f0 = kk[0][k], f1 = kk[6][k], y0 = y[k], y1 = yi[k];
                    ic[4][k] = real(2) * dts * (f1 - f0) - real(8) * (y1 + y0) + real(16) * ymid;
                    ic[3][k] = dts * (real(5) * f0 - real(3) * f1) + real(18) * y0 + real(14) * y1 - real(32) * ymid;
                    ic[2][k] = dts * (f1 - real(4) * f0) - real(11) * y0 - real(5) * y1 + real(16) * ymid;
                    ic[1][k] = dts * f0;
                    ic[0][k] = y0;
                    y[k] = y1;
                    f[k] = f1;
                }
                rk_t0 = ta;
                rk_t1 = tb;
                nacc++;
            } else {
                rk_t0 = ta;
                nrej++;
            }
            if (ratio == real(0)) {
                dt = dt * 10.0;
            } else {
                const double dfac = ratio < real(1) ? 1.0 : 0.2;
                double fac = 0.9 / pow((double)ratio, 0.2);
                fac = fac < dfac ? dfac : fac;   // NaN-propagating like torch.max/min: comparisons false keep NaN
                fac = fac > 10.0 ? 10.0 : fac;
                dt = dt * fac;
            }
        }
        if (status != 0) break;
        // _interp_evaluate at next_t
        const real x = real((next_t - rk_t0) / (rk_t1 - rk_t0));
        real v[NS], xp = x;
#pragma unroll
        for (int k = 0; k < NS; k++) v[k] = ic[0][k] + x * ic[1][k];
#pragma unroll
        for (int m = 2; m < 5; m++) {
            xp = xp * x;
#pragma unroll
            for (int k = 0; k < NS; k++) v[k] = v[k] + xp * ic[m][k];
        }
        if (io == kend) {
#pragma unroll
            for (int k = 0; k < NS; k++) yrep[k] = v[k];
        }
        if (kDense) {
#pragma unroll
            for (int k = 0; k < NS; k++) y_dense[((size_t)io * NS + k) * n + i] = m_min(m_max(v[k], p.lb), p.ub);
        }
    }
    if (kDense && status != 0) {
        for (int kq = io; kq < NTOT; kq++)
#pragma unroll
            for (int k = 0; k < NS; k++) y_dense[((size_t)kq * NS + k) * n + i] = m_min(m_max(y[k], p.lb), p.ub);
    }
#pragma unroll
    for (int k = 0; k < NS; k++) y_out[(size_t)k * n + i] = m_min(m_max(yrep[k], p.lb), p.ub);
    a.status[i] = status;
    if (a.stats) {
        a.stats[i] = nacc;
        a.stats[n + i] = nrej;
        a.stats[2 * n + i] = nrhs;
    }
}

}  // namespace pfr
