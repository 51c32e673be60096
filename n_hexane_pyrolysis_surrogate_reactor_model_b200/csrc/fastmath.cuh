// Table-driven FP64 log / exp for the integrator's hot loop.
//
// CUDA's libdevice log()/exp() cost ~24 FP64-pipe instructions each plus ~20 moves / branches (64-bit
// polynomial coefficients materialised through uniform registers, special-case slow paths).  The CRNN
// right-hand side only ever takes log of clamped positive normal numbers (Y in [1e-6, 60], T > 0) and
// exp of exponents clamped to [-30, 30] (any finite |x| < 700 is handled), so:
//   log x = e ln2 + log c_i + log1p(r),  r = m / c_i - 1,  c_i = 1 + (i + 1/2)/128,  |r| <= 2^-8, degree-6 series
//   exp x = 2^q T_j (1 + p(r)),  x = (64 q + j) ln2/64 + r,  |r| <= ln2/128, degree-6 series
// 11 / 10 FP64 instructions, coefficients as immediate constant-bank operands, one shared-memory table read.
// Measured accuracy (tests/test_gpu_parity.py::test_fast_log_exp): <= 1 ulp of the result for exp, <= 2e-15
// absolute (1 ulp at |log x| ~ 14) for log.
#pragma once
#include <cuda_runtime.h>

namespace pfr {

constexpr int LOGTAB_N = 128, EXPTAB_N = 64;
struct FastTables {
    double2 logtab[LOGTAB_N];  // (1/c_i rounded, -log(that))
    double exptab[EXPTAB_N];   // 2^(j/64)
};

constexpr double LN2_HI = 6.93147180369123816490e-01;  // fdlibm split: the high part has 21 trailing zero bits
constexpr double LN2_LO = 1.90821492927058770002e-10;

#ifndef PFR_FAST_MAGIC
#define PFR_FAST_MAGIC 1   // integer <-> double conversions of log / exp as "magic number" FP64 adds (2^52 + 2^51 shifts the integer
                           // into the low mantissa word) instead of I2F.F64 / F2I.F64 on the XU pipe: one DADD more per call on the
                           // FP64 pipe, three conversions fewer per log + exp pair on the (16 lanes / clk) XU pipe
#endif
constexpr double MAGIC_52_51 = 6755399441055744.0;   // 2^52 + 2^51: (x + MAGIC) holds rn(x) in its low word for |x| < 2^31

__device__ __forceinline__ double fast_log(double x, const double2* __restrict__ tab) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);  // [1, 2)
    const double2 t = tab[(hi >> 13) & (LOGTAB_N - 1)];
    const double r = fma(m, t.x, -1.0);
    double p = fma(r, -1.0 / 6.0, 0.2);
    p = fma(r, p, -0.25);
    p = fma(r, p, 1.0 / 3.0);
    p = fma(r, p, -0.5);
    p = fma(r * r, p, r);
#if PFR_FAST_MAGIC
    // the biased exponent dropped into the low mantissa word of 2^52 is the double 2^52 + (e + 1023), exactly
    const double ed = __hiloint2double(0x43300000, hi >> 20) - (4503599627370496.0 + 1023.0);
#else
    const double ed = (double)((hi >> 20) - 1023);
#endif
    return fma(ed, LN2_HI, t.y) + fma(ed, LN2_LO, p);
}

__device__ __forceinline__ double fast_exp(double x, const double* __restrict__ tab) {
#if PFR_FAST_MAGIC
    const double sh = fma(x, 92.33248261689366, MAGIC_52_51);   // 64 / ln 2; the sum is rounded to an integer (ties to even)
    const int k = __double2loint(sh);
    const double kd = sh - MAGIC_52_51;
#else
    const int k = __double2int_rn(x * 92.33248261689366);  // 64 / ln 2
    const double kd = (double)k;
#endif
    double r = fma(kd, -LN2_HI / 64.0, x);
    r = fma(kd, -LN2_LO / 64.0, r);
    double p = fma(r, 1.0 / 720.0, 1.0 / 120.0);
    p = fma(r, p, 1.0 / 24.0);
    p = fma(r, p, 1.0 / 6.0);
    p = fma(r, p, 0.5);
    p = fma(r * r, p, r);
    const double T = tab[k & (EXPTAB_N - 1)];
    const double res = fma(T, p, T);
    return __hiloint2double(__double2hiint(res) + ((k >> 6) << 20), __double2loint(res));
}

// 1/x to ~1e-10 (MUFU.RCP64H seed + one Newton step): enough for the weights of the error norm
__device__ __forceinline__ double rcp_norm(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return fma(r, fma(-x, r, 1.0), r);
}
__device__ __forceinline__ float rcp_norm(float x) { return __frcp_rn(x); }
__device__ __forceinline__ double rcp_full(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ float rcp_full(float x) { return __frcp_rn(x); }

// Latency-oriented variants for the kernels that run one warp per condition with no other warp to hide a dependent chain
// behind (the training step: a few hundred conditions on 148 SMs).  Same tables, same range reduction, same polynomial; the
// polynomial is evaluated Estrin-style (depth 4 instead of 6) and the two-term reconstruction is split so that the exponent part
// does not wait for the polynomial.  Results differ from fast_log / fast_exp by rounding only (<= 1 ulp).
__device__ __forceinline__ double fast_log_ilp(double x, const double2* __restrict__ tab) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 t = tab[(hi >> 13) & (LOGTAB_N - 1)];
    const double ed = __hiloint2double(0x43300000, hi >> 20) - (4503599627370496.0 + 1023.0);
    const double base = fma(ed, LN2_HI, t.y);                 // independent of the polynomial
    const double r = fma(m, t.x, -1.0);
    const double r2 = r * r;
    const double a = fma(r, 1.0 / 3.0, -0.5), b = fma(r, 0.2, -0.25);
    const double q = fma(r2, fma(r2, -1.0 / 6.0, b), a);      // -1/2 + r/3 + r^2 (-1/4 + r/5 - r^2/6)
    return base + fma(ed, LN2_LO, fma(r2, q, r));
}

__device__ __forceinline__ double fast_exp_ilp(double x, const double* __restrict__ tab) {
    const double sh = fma(x, 92.33248261689366, MAGIC_52_51);
    const int k = __double2loint(sh);
    const double kd = sh - MAGIC_52_51;
    const double T = tab[k & (EXPTAB_N - 1)];
    double r = fma(kd, -LN2_HI / 64.0, x);
    r = fma(kd, -LN2_LO / 64.0, r);
    const double r2 = r * r;
    const double a = fma(r, 1.0 / 6.0, 0.5), b = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    const double q = fma(r2, fma(r2, 1.0 / 720.0, b), a);     // 1/2 + r/6 + r^2 (1/24 + r/120 + r^2/720)
    const double res = fma(T, fma(r2, q, r), T);
    return __hiloint2double(__double2hiint(res) + ((k >> 6) << 20), __double2loint(res));
}

}  // namespace pfr
