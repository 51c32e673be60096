// CRNN right-hand side and analytic Jacobian, one condition per thread, everything in registers.
//
// Computes what the reference's CRNNFunc.forward computes
//   (SURROGATE_MODEL/surrogate_model_Eoff_single_model.py:135-153, ...Eon_single_model.py:142-151):
//     Y  = clamp(u, lb, ub)
//     z  = clamp(w_in^T [ln Y ; -1/(R_kcal T) ; ln T] + w_b, zlo, zhi)
//     du = clamp(w_out exp(z), dulo, duhi)
// re-associated for a batched integrator: the temperature part of z
//     kT_j = lnA_j - Ea_j/(R_kcal T) + b_j ln T
// is hoisted (once per trajectory when T is constant, once per stage otherwise).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pfr {

constexpr int NS = 9;      // species
constexpr int NR = 9;      // pseudo-reactions
constexpr int NTOT = 801;  // knots of the MLP-predicted grids

// Passed by value as a __grid_constant__ kernel argument: lives in the constant bank, every access is
// warp-uniform.  nu[k][j]: reaction order of species k in reaction j (w_in rows 0..8).
template <typename real>
struct CrnnParams {
    real nu[NS][NR];
    real Ea[NR], b[NR], lnA[NR];
    real wout[NS][NR];
    real lb, ub, zlo, zhi, dulo, duhi;
    real inv_R;  // 1 / R_kcal
};

template <typename real> __device__ __forceinline__ real m_log(real x);
template <> __device__ __forceinline__ double m_log<double>(double x) { return log(x); }
template <> __device__ __forceinline__ float m_log<float>(float x) { return logf(x); }
template <typename real> __device__ __forceinline__ real m_exp(real x);
template <> __device__ __forceinline__ double m_exp<double>(double x) { return exp(x); }
template <> __device__ __forceinline__ float m_exp<float>(float x) { return expf(x); }
template <typename real> __device__ __forceinline__ real m_sqrt(real x);
template <> __device__ __forceinline__ double m_sqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float m_sqrt<float>(float x) { return sqrtf(x); }
template <typename real> __device__ __forceinline__ real m_abs(real x) { return x < real(0) ? -x : x; }
template <typename real> __device__ __forceinline__ real m_max(real a, real b) { return a > b ? a : b; }
template <typename real> __device__ __forceinline__ real m_min(real a, real b) { return a < b ? a : b; }

// Temperature part of the exponent.  Also returns d kT_j / dT when WithDeriv (for df/dt on a T ramp).
template <typename real, bool WithDeriv>
__device__ __forceinline__ void arrhenius_T(const CrnnParams<real>& p, real T, real (&kT)[NR], real (&dkT)[NR]) {
    const real invT = real(1) / T;
    const real mE = -p.inv_R * invT;  // -1 / (R T)
    const real lnT = m_log<real>(T);
#pragma unroll
    for (int j = 0; j < NR; j++) {
        kT[j] = fma(p.Ea[j], mE, fma(p.b[j], lnT, p.lnA[j]));
        if (WithDeriv) dkT[j] = (p.b[j] - p.Ea[j] * mE) * invT;  // Ea/(R T^2) + b/T
    }
}

// du = f(y) at fixed kT.  Optionally keeps what the Jacobian needs: g_j = r_j * [z_j not clamped],
// q_k = [y_k not clamped] / Y_k, md_i = [du_i not clamped].
template <typename real, bool KeepJac>
__device__ __forceinline__ void crnn_rhs(const CrnnParams<real>& p, const real (&kT)[NR], const real (&y)[NS],
                                         real (&du)[NS], real (&g)[NR], real (&q)[NS], real (&md)[NS]) {
    real lnY[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) {
        const real Y = m_min(m_max(y[k], p.lb), p.ub);
        lnY[k] = m_log<real>(Y);
        if (KeepJac) q[k] = (y[k] >= p.lb && y[k] <= p.ub) ? real(1) / Y : real(0);
    }
    real r[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) {
        real z = kT[j];
#pragma unroll
        for (int k = 0; k < NS; k++) z = fma(p.nu[k][j], lnY[k], z);
        const real zc = m_min(m_max(z, p.zlo), p.zhi);
        r[j] = m_exp<real>(zc);
        if (KeepJac) g[j] = (z >= p.zlo && z <= p.zhi) ? r[j] : real(0);
    }
#pragma unroll
    for (int i = 0; i < NS; i++) {
        real s = real(0);
#pragma unroll
        for (int j = 0; j < NR; j++) s = fma(p.wout[i][j], r[j], s);
        du[i] = m_min(m_max(s, p.dulo), p.duhi);
        if (KeepJac) md[i] = (s >= p.dulo && s <= p.duhi) ? real(1) : real(0);
    }
}

// Streaming variant for the Rosenbrock kernel: vectors live in the thread's shared-memory scratch
// (entry e at sm[e * STRIDE]) so that the species / reaction loops can stay rolled (UNROLL < 9), which
// bounds how many software log/exp expansions the compiler keeps in flight and hence the register count.
//   in : y at entries [in_off, in_off+9)
//   out: du (registers); entries [w_off, w_off+9) are overwritten with the un-clamped exponents z_j
//   KeepJac: q_k -> entries [jac_off, jac_off+9), g_j -> entries [jac_off+9, jac_off+18)
template <typename real, int STRIDE, int UNROLL, bool KeepJac>
__device__ __forceinline__ void crnn_rhs_sm(const CrnnParams<real>& p, const real (&kT)[NR], real* __restrict__ sm,
                                            int in_off, int w_off, int jac_off, real (&du)[NS]) {
    real z[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) z[j] = kT[j];
#pragma unroll UNROLL
    for (int k = 0; k < NS; k++) {
        const real yk = sm[(in_off + k) * STRIDE];
        const real Y = m_min(m_max(yk, p.lb), p.ub);
        const real l = m_log<real>(Y);
        if (KeepJac) sm[(jac_off + k) * STRIDE] = (yk >= p.lb && yk <= p.ub) ? real(1) / Y : real(0);
#pragma unroll
        for (int j = 0; j < NR; j++) z[j] = fma(p.nu[k][j], l, z[j]);
    }
#pragma unroll
    for (int j = 0; j < NR; j++) sm[(w_off + j) * STRIDE] = z[j];
#pragma unroll
    for (int i = 0; i < NS; i++) du[i] = real(0);
#pragma unroll UNROLL
    for (int j = 0; j < NR; j++) {
        const real zj = sm[(w_off + j) * STRIDE];
        const real r = m_exp<real>(m_min(m_max(zj, p.zlo), p.zhi));
        if (KeepJac) sm[(jac_off + NS + j) * STRIDE] = (zj >= p.zlo && zj <= p.zhi) ? r : real(0);
#pragma unroll
        for (int i = 0; i < NS; i++) du[i] = fma(p.wout[i][j], r, du[i]);
    }
#pragma unroll
    for (int i = 0; i < NS; i++) du[i] = m_min(m_max(du[i], p.dulo), p.duhi);
}

}  // namespace pfr
