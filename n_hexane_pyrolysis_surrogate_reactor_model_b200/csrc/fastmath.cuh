// Table-driven FP64 log / exp for the integrator's hot loop.
//
// CUDA's libdevice log()/exp() cost ~24 FP64-pipe instructions each plus ~20 moves / branches (64-bit
// polynomial coefficients materialised through uniform registers, special-case slow paths).  The CRNN
// right-hand side only ever takes log of clamped positive normal numbers (Y in [1e-6, 60], T > 0) and
// exp of exponents clamped to [-30, 30] (any finite |x| < 700 is handled), so:
//   log x = e ln2 + log c_i + log1p(r),  r = m / c_i - 1,  c_i = 1 + (i + 1/2)/256,  |r| <= 2^-9, degree-4 minimax
//   exp x = 2^q T_j (1 + p(r)),  x = (256 q + j) ln2/256 + r,  |r| <= ln2/512, degree-4 series (r^5/120 < 4e-17)
// 8 / 9 FP64 instructions (round 1: 128 / 64-entry tables, degree-6 series, 11 / 11), coefficients as immediate
// constant-bank operands, one shared-memory table read; fast_exp_scaled takes its argument in units of ln2/256 (the
// explicit integrators pre-scale the exponent coefficients of the CRNN, so the range reduction is one exact subtraction): 8.
// Measured accuracy (tests/test_gpu_parity.py::test_fast_log_exp): <= 1 ulp of the result for exp, <= 3e-15
// absolute (1.5 ulp at |log x| ~ 14) for log.
#pragma once
#include <cuda_runtime.h>

namespace pfr {

constexpr int LOGTAB_N = 256, EXPTAB_N = 256, LOGTAB_SHIFT = 12, EXPTAB_BITS = 8;
struct FastTables {
    double2 logtab[LOGTAB_N];  // (1/c_i rounded, -log(that))
    double exptab[EXPTAB_N];   // 2^(j/256)
};

constexpr double LN2 = 6.93147180559945286227e-01;     // (nearest double: 2.3e-17 below ln 2)
constexpr double LN2_HI = 6.93147180369123816490e-01;  // fdlibm split: the high part has 21 trailing zero bits
constexpr double LN2_LO = 1.90821492927058770002e-10;
constexpr double EXP_ARG_SCALE = 369.329930467574632284;   // 256 / ln 2: fast_exp_scaled(x * EXP_ARG_SCALE) = exp(x)

// log1p(r) = r + r^2 (L2 + r (L3 + r L4)) on |r| <= 2^-9: minimax fit of (log1p(r) - r) / r^2, max error 7.4e-16 (the series
// truncated after r^4 / 4: 5.7e-15)
constexpr double LOG_L2 = -0.49999999999873539862, LOG_L3 = 0.33333399640540911563, LOG_L4 = -0.25000096727866272713;

#ifndef PFR_FAST_MAGIC
#define PFR_FAST_MAGIC 3   // bit 0: log, bit 1: exp -- integer <-> double conversions as "magic number" FP64 adds (2^52 + 2^51 shifts the integer
                           // into the low mantissa word) instead of I2F.F64 / F2I.F64 on the XU pipe: one DADD more per call on the
                           // FP64 pipe, three conversions fewer per log + exp pair on the (16 lanes / clk) XU pipe
#endif
constexpr double MAGIC_52_51 = 6755399441055744.0;   // 2^52 + 2^51: (x + MAGIC) holds rn(x) in its low word for |x| < 2^31

__device__ __forceinline__ double fast_log(double x, const double2* __restrict__ tab) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);  // [1, 2)
    const double2 t = tab[(hi >> LOGTAB_SHIFT) & (LOGTAB_N - 1)];
    const double r = fma(m, t.x, -1.0);
    double p = fma(r, LOG_L4, LOG_L3);
    p = fma(r, p, LOG_L2);
    p = fma(r * r, p, r);
#if PFR_FAST_MAGIC & 1
    // the biased exponent dropped into the low mantissa word of 2^52 is the double 2^52 + (e + 1023), exactly
    const double ed = __hiloint2double(0x43300000, hi >> 20) - (4503599627370496.0 + 1023.0);
#else
    const double ed = (double)((hi >> 20) - 1023);
#endif
    // |e| <= 1023 here but <= 20 for everything the right-hand side feeds in: e * (ln 2 - LN2) <= 5e-16 there, the rounding of
    // the final sum (half an ulp of a result of size ~14) is the larger term
    return fma(ed, LN2, t.y + p);
}

// exponent adjustment 2^q applied to the high word
__device__ __forceinline__ double exp_finish(double T, double p, int k) {
    const double res = fma(T, p, T);
    // 2^q with q = k >> 8 into the exponent field: (k with its table index cleared) << 12, one mask and one shift-add (LEA)
    // (written as a multiply-add so that it stays ONE integer instruction behind the mask: the compiler's own form of
    // ((k >> 8) << 20) + hi is shift, mask, add)
    int hi;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(hi) : "r"(k & ~(EXPTAB_N - 1)), "n"(1 << (20 - EXPTAB_BITS)), "r"(__double2hiint(res)));
    return __hiloint2double(hi, __double2loint(res));
}

__device__ __forceinline__ double fast_exp(double x, const double* __restrict__ tab) {
#if PFR_FAST_MAGIC & 2
    const double sh = fma(x, EXP_ARG_SCALE, MAGIC_52_51);   // the sum is rounded to an integer (ties to even)
    const int k = __double2loint(sh);
    const double kd = sh - MAGIC_52_51;
#else
    const int k = __double2int_rn(x * EXP_ARG_SCALE);
    const double kd = (double)k;
#endif
    double r = fma(kd, -LN2_HI / EXPTAB_N, x);
    r = fma(kd, -LN2_LO / EXPTAB_N, r);
    double p = fma(r, 1.0 / 24.0, 1.0 / 6.0);
    p = fma(r, p, 0.5);
    p = fma(r * r, p, r);
    return exp_finish(tab[k & (EXPTAB_N - 1)], p, k);
}

// exp(xs ln2 / 256): the argument arrives in units of ln2 / 256, so the reduced argument d = xs - rn(xs) is exact and the series
// is taken in d with the powers of ln2 / 256 folded into its coefficients
__device__ __forceinline__ double fast_exp_scaled(double xs, const double* __restrict__ tab) {
    constexpr double C1 = LN2 / EXPTAB_N, C2 = C1 * C1 / 2.0, C3 = C1 * C1 * C1 / 6.0, C4 = C1 * C1 * C1 * C1 / 24.0;
#if PFR_FAST_MAGIC & 2
    const double sh = xs + MAGIC_52_51;
    const int k = __double2loint(sh);
    const double d = xs - (sh - MAGIC_52_51);
#else
    const int k = __double2int_rn(xs);          // F2I + I2F on the XU pipe: two issue slots instead of the six of three FP64 adds
    const double d = xs - (double)k;
#endif
    double p = fma(d, C4, C3);
    p = fma(d, p, C2);
    p = fma(d, p, C1);
    return exp_finish(tab[k & (EXPTAB_N - 1)], d * p, k);
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
// polynomial is evaluated Estrin-style and the two-term reconstruction is split so that the exponent part
// does not wait for the polynomial.  Results differ from fast_log / fast_exp by rounding only (<= 1 ulp).
__device__ __forceinline__ double fast_log_ilp(double x, const double2* __restrict__ tab) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 t = tab[(hi >> LOGTAB_SHIFT) & (LOGTAB_N - 1)];
    const double ed = __hiloint2double(0x43300000, hi >> 20) - (4503599627370496.0 + 1023.0);
    const double base = fma(ed, LN2_HI, t.y);                 // independent of the polynomial
    const double r = fma(m, t.x, -1.0);
    const double r2 = r * r;
    const double a = fma(r, LOG_L3, LOG_L2);
    const double q = fma(r2, LOG_L4, a);                      // L2 + L3 r + L4 r^2
    return base + fma(ed, LN2_LO, fma(r2, q, r));
}

__device__ __forceinline__ double fast_exp_ilp(double x, const double* __restrict__ tab) {
    const double sh = fma(x, EXP_ARG_SCALE, MAGIC_52_51);
    const int k = __double2loint(sh);
    const double kd = sh - MAGIC_52_51;
    const double T = tab[k & (EXPTAB_N - 1)];
    double r = fma(kd, -LN2_HI / EXPTAB_N, x);
    r = fma(kd, -LN2_LO / EXPTAB_N, r);
    const double r2 = r * r;
    const double a = fma(r, 1.0 / 6.0, 0.5);
    const double q = fma(r2, 1.0 / 24.0, a);                  // 1/2 + r/6 + r^2/24
    return exp_finish(T, fma(r2, q, r), k);
}

}  // namespace pfr
