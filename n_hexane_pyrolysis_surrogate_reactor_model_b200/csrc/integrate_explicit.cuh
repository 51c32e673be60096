// Explicit integrators for the NON-STIFF regime of the CRNN plug-flow problem, one PFR condition per thread:
//   bs23_kernel  Bogacki-Shampine 3(2) with FSAL, knot-limited steps      (the coupled Eon sweep; described first)
//   dp54_kernel  Dormand-Prince 5(4) with FSAL, free stepping to t_end    (the isothermal Eoff sweep; second half of the file)
//
// Why it exists.  On the coupled (Eon) path the temperature is the MLP's piecewise-linear profile and a step never
// crosses one of its 801 knots (integrate_rodas.cuh explains why).  The knot spacing (median 5e-4 s) is then 10-50x
// SMALLER than the step an explicit method could take stably: the reference's own dopri5 needs 8-55 accepted steps
// for a whole trajectory (tests/golden/reference_vectors.npz, Eon/dopri5_stats), so h * rho(J) << 1 on every knot
// interval.  A Rosenbrock step pays for a 9x9 Jacobian and its inverse (half of its FP64 work) that buy nothing
// there.  One BS23 step per knot interval is three right-hand sides and nothing else: 2.4x less FP64 work per
// trajectory than ROS3 at the same accuracy (work-precision table).
//
// Why one thread per condition here (the Rosenbrock kernels use three lanes).  Without the Jacobian the live state is
// four 9-vectors; it fits in registers, there are no shuffles at all, every thread has nine independent log / exp /
// dot-product chains in flight, and a warp instruction serves 32 conditions instead of 10.  The CRNN coefficients are
// broadcast reads from a block-shared copy (two per 16-byte load).
//
// Stiffness guard.  The embedded error estimate of an explicit pair blows up when h * lambda leaves the stability
// region, so a stiff interval shows up as a cascade of rejected sub-steps.  A condition that needs more than
// 8 attempts per knot interval on average (+ 512) stops with status PFR_ST_STIFF; the host (Surrogate) then integrates
// exactly those conditions with the Rosenbrock kernel.  (An accuracy-limited explicit run takes 1.0-1.6 attempts per
// interval at rtol = atol = 1e-5 ... 1e-9 and ~3.5 at 1e-10; error control holds either way, the guard only bounds the
// wasted work: at most ~3x the cost of the Rosenbrock integration it falls back to.)  The
// product stays an adaptive stiff-capable integrator; this is its fast path.
//
// Same controller conventions as the Rosenbrock kernels (RMS norm against atol + rtol max(|y0|, |y1|) -- bs23 weighs with |y1|
// alone, PFR_NORM_NEW_ONLY --, Hairer-style first step, factor clip(0.9 err^(-1/3), 0.2, 6)); a knot-clipped accepted step does
// not shrink the proposal.
#pragma once
#include "crnn_device.cuh"
#include "fastmath.cuh"
#include "integrate_rodas.cuh"

namespace pfr {

#ifndef PFR_BS23_SMEM_STATE
#define PFR_BS23_SMEM_STATE 1   // FSAL slope and the two running combinations of a step in shared memory (27 KB per CTA) instead of
                                // registers: 252 -> 168 registers without spills, i.e. three CTAs per SM; measured 84.5 -> 78.4 ms
                                // for 2^20 conditions at 1e-8 (with two CTAs the shared-memory version is 4 % slower than registers)
#endif
#ifndef PFR_BS23_BLOCK
#define PFR_BS23_BLOCK 128
#endif
template <typename real> constexpr size_t bs23_smem_bytes() { return PFR_BS23_SMEM_STATE ? (size_t)3 * NS * PFR_BS23_BLOCK * sizeof(real) : 0; }
#ifndef PFR_BS23_KINK
#define PFR_BS23_KINK 0     // error margin demanded of a step that straddles the kink of the lower state clamp; 0 = off.
                            // Measured with 100 (2^20 LHS conditions, 1e-8): median outlet error 2.1e-7 -> 6.1e-8, p99 3.2e-6 ->
                            // 1.2e-6, max unchanged, kernel +7 % (LLNL) ... +16 % of the step (JetSurf): the knot-limited steps
                            // are short enough without it, so it is left off; the free-stepping dp54_kernel needs it.
#endif
#ifndef PFR_NORM_NEW_ONLY
#define PFR_NORM_NEW_ONLY 1   // bs23: error weights atol + rtol |y1| instead of atol + rtol max(|y0|, |y1|) -- never looser, so the error control
                              // only tightens; saves DSETP + 2 FSEL per species and step: 63.7 -> 62.7 ms at 2^20 conditions (r02m / r02n A/B)
#endif
constexpr int BS23_BLOCK = PFR_BS23_BLOCK;
#ifndef PFR_BS23_MINB
#define PFR_BS23_MINB (PFR_BS23_SMEM_STATE ? 3 : 2)   // CTAs per SM: 168 registers / 12 warps with the shared-memory state, else 252 / 8
#endif
constexpr int BS23_CTAS_PER_SM = PFR_BS23_MINB;   // persistent grid: as many CTAs as fit
constexpr int PFR_ST_STIFF_ = 4;

template <typename real> __device__ __forceinline__ real t_log(real x, const FastTables& ft);
template <> __device__ __forceinline__ double t_log<double>(double x, const FastTables& ft) { return fast_log(x, ft.logtab); }
template <> __device__ __forceinline__ float t_log<float>(float x, const FastTables&) { return logf(x); }
template <typename real> __device__ __forceinline__ real t_exp(real x, const FastTables& ft);
template <> __device__ __forceinline__ double t_exp<double>(double x, const FastTables& ft) { return fast_exp(x, ft.exptab); }
template <> __device__ __forceinline__ float t_exp<float>(float x, const FastTables&) { return expf(x); }
// The explicit kernels carry the exponents z_j in units of ln2 / 256 (double; plain units in float): nu, Ea, b, lnA and the exponent
// clamp are pre-scaled once (fill_coef_dup / load_tpc_coef with kScaled), which turns the range reduction of the exponential into
// one exact subtraction (fast_exp_scaled: 8 FP64 instructions instead of 9).
template <typename real> __host__ __device__ constexpr real z_scale() { return sizeof(real) == 8 ? real(EXP_ARG_SCALE) : real(1); }
template <typename real> __device__ __forceinline__ real t_exp_scaled(real x, const FastTables& ft);
template <> __device__ __forceinline__ double t_exp_scaled<double>(double x, const FastTables& ft) { return fast_exp_scaled(x, ft.exptab); }
template <> __device__ __forceinline__ float t_exp_scaled<float>(float x, const FastTables&) { return expf(x); }

// Block-shared copy of the CRNN coefficients, laid out in the order the right-hand side consumes them.  Every thread
// reads the same address (a broadcast), two coefficients per 16-byte load.  (As FMA operands straight from the
// constant bank they would have to pass through uniform registers on sm_100: 189 64-bit values do not fit, and the
// compiler spills uniform registers into vector registers and back -- measured: 25 % of the instruction stream.)
template <typename real>
struct TpcCoef {
    real arr[NR][4];      // Ea_j, b_j, lnA_j, 0
    real nu[NS][10];      // nu[k][j], j = 0..8, 0
    real woutT[NR][10];   // wout[i][j] stored [j][i], i = 0..8, 0
    FastTables ft;
};

// Second feed of the same coefficients: the kernel-parameter constant bank, read through UNIFORM registers (LDCU.64, then
// DFMA R, R, UR, R).  Measured on this part (tools/micro/coef_paths.cu, dfma_regs.cu): a 9x9 mat-vec whose coefficients all arrive
// by broadcast LDS.128 runs at 45 % of the DFMA rate and one fed entirely by LDCU.64 at 50 %, at any occupancy, while the same
// mat-vec with its coefficients in registers runs at 93 %: the feed costs as much as the arithmetic, whichever of the two paths
// delivers it (a 16-byte shared-memory load returns 512 bytes per warp through the register file's write port, which the
// 256-byte results of the DFMAs already keep busy; the constant cache serves one uniform load per clock and SM).  The two paths
// are independent, so the rows of a mat-vec are SPLIT between them (PFR_UNIFORM_ROWS: bit k = row k comes through uniform
// registers).  The block is stored twice and the copy toggles with every right-hand side -- a loop-carried, warp-uniform index --
// because loads from a fixed parameter address are loop-invariant and both nvcc and ptxas would hoist all of them out of the
// step loop (and spill them).
#ifndef PFR_UNIFORM_ROWS
#define PFR_UNIFORM_ROWS 0x1ff   // all nine rows of both mat-vecs (measured: 82.4 ms with none, 75.7 with three, 71.5 with six, 70.2 ms with all nine)
#endif
template <typename real>
struct __align__(16) CoefDup {
    real nu[2][NS][10];      // [copy][k][j]
    real woutT[2][NR][10];   // [copy][j][i]
    real arr[2][NR][4];      // [copy][j]: Ea_j, b_j, lnA_j, 0
};
template <typename real>
static inline void fill_coef_dup(CoefDup<real>& d, const CrnnParams<real>& p) {
    const real zs = z_scale<real>();   // (nu and the Arrhenius coefficients in units of ln2 / 256, see t_exp_scaled)
    for (int c = 0; c < 2; c++)
        for (int r = 0; r < NS; r++)
            for (int e = 0; e < 10; e++) {
                d.nu[c][r][e] = e < NR ? zs * p.nu[r][e] : real(0);
                d.woutT[c][r][e] = e < NS ? p.wout[e][r] : real(0);
            }
    for (int c = 0; c < 2; c++)
        for (int j = 0; j < NR; j++) {
            d.arr[c][j][0] = zs * p.Ea[j];
            d.arr[c][j][1] = zs * p.b[j];
            d.arr[c][j][2] = zs * p.lnA[j];
            d.arr[c][j][3] = real(0);
        }
}

// Third feed, an experiment that is OFF by default (PFR_CONST_COEF=1 builds it): the same block in __constant__ memory at FIXED
// addresses, read two coefficients per uniform 16-byte load (LDCU.128).  ptxas only ever emits LDCU.64 when the address carries an
// index (the copy toggle above), so this path uses immediate offsets, and the loads are volatile inline PTX so that nvcc cannot
// hoist 162 loop-invariant values out of the step loop.  In isolation it is the fastest feed (tools/micro/coef_paths.cu,
// profiles/r02o_coef_feed_microbench.txt: a 9x9 mat-vec pair at 21.4 TFLOP/s against 18.3 with LDCU.64 and 16.6 with broadcast
// LDS.128).  In the kernels it is not: ptxas parks the first ~23 coefficients in uniform registers for the whole loop (they ARE
// loop-invariant) and streams all the others through the four registers that are left -- every LDCU.128 waits for the two
// DFMAs that read its predecessor, the load latency is exposed 70 times per right-hand side: bs23_kernel 62.0 -> 76.9 ms,
// dp54_kernel 6.5 -> 6.25 ms at 2^20 conditions (profiles/r02s_const_coef_ab.jsonl), results bit-identical.  The block is one
// per device: the host side (stage_const_coef in capi.cu) re-uploads it in stream order when the model changes and chains the
// launches that read it.
#ifndef PFR_CONST_COEF
#define PFR_CONST_COEF 0
#endif
struct __align__(16) ConstCoef {
    double nu[NS][10];      // [k][j], in units of ln2 / 256 (z_scale), j = 9: 0
    double woutT[NR][10];   // [j][i], i = 9: 0
};
extern "C" { __constant__ ConstCoef pfr_const_coef; }
static inline void fill_const_coef(ConstCoef& d, const CrnnParams<double>& p) {
    for (int r = 0; r < NS; r++)
        for (int e = 0; e < 10; e++) {
            d.nu[r][e] = e < NR ? z_scale<double>() * p.nu[r][e] : 0.0;
            d.woutT[r][e] = e < NS ? p.wout[e][r] : 0.0;
        }
}
template <int kOffset>
__device__ __forceinline__ void ldc2(double& a, double& b) {
    asm volatile("ld.const.v2.f64 {%0, %1}, [pfr_const_coef+%2];" : "=d"(a), "=d"(b) : "n"(kOffset));
}
// acc[j] += C[kRow][j] * x, j = 0..8, for the matrix that starts kBase bytes into the block
template <int kBase, int kRow, int kJ = 0>
__device__ __forceinline__ void row_fma_const(double x, double (&acc)[9]) {
    double a, b;
    ldc2<kBase + (kRow * 10 + kJ) * 8>(a, b);
    acc[kJ] = fma(a, x, acc[kJ]);
    if constexpr (kJ + 1 < 9) acc[kJ + 1] = fma(b, x, acc[kJ + 1]);
    if constexpr (kJ + 2 < 9) row_fma_const<kBase, kRow, kJ + 2>(x, acc);
}

// volatile: the values are loop-invariant; a plain load lets the compiler hoist all of them out of the step loop.  `after`
// is an artificial input: the load may not be scheduled before that value exists, which keeps the nine coefficient rows
// of a mat-vec from being fetched (and held in 180 registers) ahead of the logarithms / exponentials they multiply.
__device__ __forceinline__ void lds2(const double* q, double& a, double& b, double after) {
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"((unsigned)__cvta_generic_to_shared(q)), "d"(after));
}
__device__ __forceinline__ void lds2(const float* q, float& a, float& b, float after) {
    asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"((unsigned)__cvta_generic_to_shared(q)), "f"(after));
}
template <typename real>
__device__ __forceinline__ void lds9(const real* q, real (&c)[10], real after) {
#pragma unroll
    for (int e = 0; e < 5; e++) lds2(q + 2 * e, c[2 * e], c[2 * e + 1], after);
}

// conservative |x| >= bound test on the high word (double) so that it runs on the integer pipe; `thr` = bound_key(bound)
__device__ __forceinline__ int bound_key(double b) { return __double2hiint(fabs(b)); }
__device__ __forceinline__ int bound_key(float b) { return __float_as_int(fabsf(b)); }
__device__ __forceinline__ bool maybe_outside(double x, int thr) { return (__double2hiint(x) & 0x7fffffff) >= thr; }
__device__ __forceinline__ bool maybe_outside(float x, int thr) { return (__float_as_int(x) & 0x7fffffff) >= thr; }
// the same test over a whole 9-vector: "some |x_j| >= bound (or NaN)".  Positive entries are caught by the SIGNED maximum of the high
// words (negative doubles are negative integers there), negative ones by the UNSIGNED maximum (their sign bit makes them the largest
// unsigned values, ordered by magnitude): two 3-input integer max chains (VIMNMX3) and two compares -- 10 instructions instead of
// the 18 of nine masked compares.
__device__ __forceinline__ int key_word(double x) { return __double2hiint(x); }
__device__ __forceinline__ int key_word(float x) { return __float_as_int(x); }
template <typename real, int N>
__device__ __forceinline__ bool any_outside(const real (&x)[N], int thr) {
    static_assert(N % 2 == 1, "pairs after the first entry");
    int ms = key_word(x[0]);
    unsigned mu = (unsigned)ms;
#pragma unroll
    for (int j = 1; j < N; j += 2) {
        const int a = key_word(x[j]), b = key_word(x[j + 1]);
        ms = max(ms, max(a, b));
        mu = max(mu, max((unsigned)a, (unsigned)b));
    }
    return ms >= thr || mu >= ((unsigned)thr | 0x80000000u);
}

// Temperature part of the exponents: kT_j = lnA_j - Ea_j / (R T) + b_j ln T
template <typename real>
__device__ __forceinline__ void arrhenius_tpc(const CrnnParams<real>& p, const TpcCoef<real>& sc, real T, real (&kT)[NR]) {
    const real invT = rcp_full(T);
    const real mE = -p.inv_R * invT;
    const real lnT = t_log<real>(T, sc.ft);
#pragma unroll
    for (int j = 0; j < NR; j++) {
        real Ea, b, lnA, pad;
        lds2(&sc.arr[j][0], Ea, b, lnT);
        lds2(&sc.arr[j][2], lnA, pad, lnT);
        kT[j] = fma(Ea, mE, fma(b, lnT, lnA));
    }
}

// the same with the coefficients from uniform registers (CoefDup)
template <typename real>
__device__ __forceinline__ void arrhenius_uni(const CrnnParams<real>& p, const TpcCoef<real>& sc, const CoefDup<real>& cd, int copy, real T, real (&kT)[NR]) {
    const real invT = rcp_full(T);
    const real mE = -p.inv_R * invT;
    const real lnT = t_log<real>(T, sc.ft);
#pragma unroll
    for (int j = 0; j < NR; j++) kT[j] = fma(cd.arr[copy][j][0], mE, fma(cd.arr[copy][j][1], lnT, cd.arr[copy][j][2]));
}

// z_j += sum_k nu[k][j] ln clamp(y_k): the first mat-vec of the right-hand side, streaming over the species (kUpper: with the upper
// state clamp; see rhs_tpc_kT)
template <bool kUpper, int kK = 0>
__device__ __forceinline__ void exponents_const(const CrnnParams<double>& p, const FastTables& ft, const double (&y)[NS], double (&z)[NR]) {
    const double yc = m_max(y[kK], p.lb);
    row_fma_const<0, kK>(fast_log(kUpper ? m_min(yc, p.ub) : yc, ft.logtab), z);
    if constexpr (kK + 1 < NS) exponents_const<kUpper, kK + 1>(p, ft, y, z);
}
template <int kJ = 0>
__device__ __forceinline__ void rates_const(const FastTables& ft, const double (&z)[NR], double (&du)[NS]) {
    row_fma_const<(int)sizeof(double) * NS * 10, kJ>(fast_exp_scaled(z[kJ], ft.exptab), du);
    if constexpr (kJ + 1 < NR) rates_const<kJ + 1>(ft, z, du);
}

template <typename real, int kUniRows, bool kUpper>
__device__ __forceinline__ void exponents_tpc(const CrnnParams<real>& p, const TpcCoef<real>& sc, const CoefDup<real>& cd, int copy,
                                              const real (&y)[NS], real (&z)[NR]) {
    if constexpr (sizeof(real) == 8 && PFR_CONST_COEF) {
        exponents_const<kUpper>(p, sc.ft, y, z);
    } else {
#pragma unroll
        for (int k = 0; k < NS; k++) {
            const real yc = m_max(y[k], p.lb);
            const real l = t_log<real>(kUpper ? m_min(yc, p.ub) : yc, sc.ft);
            if ((kUniRows >> k) & 1) {
#pragma unroll
                for (int j = 0; j < NR; j++) z[j] = fma(cd.nu[copy][k][j], l, z[j]);
            } else {
                real c[10];
                lds9(sc.nu[k], c, l);
#pragma unroll
                for (int j = 0; j < NR; j++) z[j] = fma(c[j], l, z[j]);
            }
        }
    }
}

// du = f(y) at given kT: streaming form, 18 live values (the exponents z_j, then the sums du_i).  zthr / dthr: bound_key of
// min(|zlo|, |zhi|) and min(|dulo|, |duhi|): the exponent and output clamps are skipped when no entry comes near them.
template <typename real, int kUniRows>
__device__ __forceinline__ void rhs_tpc_kT(const CrnnParams<real>& p, const TpcCoef<real>& sc, const CoefDup<real>& cd, int copy, int zthr, int dthr,
                                           const real (&kT)[NR], const real (&y)[NS], real (&du)[NS]) {
    // (kT, nu and therefore z are in units of ln2 / 256 -- z_scale; zthr is the key of the scaled clamp bound)
    real z[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) z[j] = kT[j];
    // (The state clamp stays 2 DSETP + 4 FSEL per species.  Measured alternatives, both SLOWER: deciding it on the integer pipe with the
    // exact clamp in a branch taken when some species is not strictly inside, -6 %; handling only the upper bound that way -- a signed
    // maximum over the high words, REDUX over the warp so that the branch is uniform, a second copy of this loop with the two-sided
    // clamp behind it -- 64.0 -> 69.4 ms: the branch keeps the scheduler from overlapping the first logarithms with the tail of the
    // previous stage.)
    exponents_tpc<real, kUniRows, true>(p, sc, cd, copy, y, z);
    if (any_outside(z, zthr)) {
#pragma unroll
        for (int j = 0; j < NR; j++) z[j] = m_min(m_max(z[j], z_scale<real>() * p.zlo), z_scale<real>() * p.zhi);
    }
#pragma unroll
    for (int i = 0; i < NS; i++) du[i] = real(0);
    if constexpr (sizeof(real) == 8 && PFR_CONST_COEF) {
        rates_const(sc.ft, z, du);
    } else {
#pragma unroll
        for (int j = 0; j < NR; j++) {
            const real r = t_exp_scaled<real>(z[j], sc.ft);
            if ((kUniRows >> j) & 1) {
#pragma unroll
                for (int i = 0; i < NS; i++) du[i] = fma(cd.woutT[copy][j][i], r, du[i]);
            } else {
                real c[10];
                lds9(sc.woutT[j], c, r);
#pragma unroll
                for (int i = 0; i < NS; i++) du[i] = fma(c[i], r, du[i]);
            }
        }
    }
    if (any_outside(du, dthr)) {
#pragma unroll
        for (int i = 0; i < NS; i++) du[i] = m_min(m_max(du[i], p.dulo), p.duhi);
    }
}

// du = f(T, y)
template <typename real, int kUniRows>
__device__ __forceinline__ void rhs_tpc(const CrnnParams<real>& p, const TpcCoef<real>& sc, const CoefDup<real>& cd, int copy, int zthr, int dthr, real T,
                                        const real (&y)[NS], real (&du)[NS]) {
    real kT[NR];
    arrhenius_tpc<real>(p, sc, T, kT);   // (its 27 coefficients stay on the shared-memory path: with them on the uniform path as well,
                                         // ptxas 12.9 keeps the copy index in a vector register and every uniform load of the
                                         // right-hand side turns into a per-thread constant load)
    rhs_tpc_kT<real, kUniRows>(p, sc, cd, copy, zthr, dthr, kT, y, du);
}

// block-shared copy of the coefficients and tables (every kernel of this file starts with it)
template <typename real, int kBlock, bool kScaled = false>
__device__ __forceinline__ void load_tpc_coef(TpcCoef<real>& sc, const CrnnParams<real>& p, const FastTables* tables) {
    const real zs = kScaled ? z_scale<real>() : real(1);   // kScaled: exponent coefficients in units of ln2 / 256 (t_exp_scaled)
    for (int e = threadIdx.x; e < NS * 10; e += kBlock) {
        const int r = e / 10, c = e % 10;
        sc.nu[r][c] = c < NR ? zs * p.nu[r][c] : real(0);
        sc.woutT[r][c] = c < NS ? p.wout[c][r] : real(0);
    }
    if (threadIdx.x < NR) {
        sc.arr[threadIdx.x][0] = zs * p.Ea[threadIdx.x];
        sc.arr[threadIdx.x][1] = zs * p.b[threadIdx.x];
        sc.arr[threadIdx.x][2] = zs * p.lnA[threadIdx.x];
        sc.arr[threadIdx.x][3] = real(0);
    }
    if (sizeof(real) == 8) {
        for (int e = threadIdx.x; e < LOGTAB_N; e += kBlock) sc.ft.logtab[e] = tables->logtab[e];
        for (int e = threadIdx.x; e < EXPTAB_N; e += kBlock) sc.ft.exptab[e] = tables->exptab[e];
    }
    __syncthreads();
}

template <typename real, bool kRamp>
__global__ void __launch_bounds__(BS23_BLOCK, PFR_BS23_MINB)
bs23_kernel(const __grid_constant__ CrnnParams<real> p, const __grid_constant__ CoefDup<real> cd, const RodasArgs a) {
    __shared__ __align__(16) TpcCoef<real> sc;
    load_tpc_coef<real, BS23_BLOCK, true>(sc, p, a.tables);
    const int zthr = bound_key(z_scale<real>() * m_min(m_abs(p.zlo), m_abs(p.zhi))), dthr = bound_key(m_min(m_abs(p.dulo), m_abs(p.duhi)));
    const size_t n = (size_t)a.n;
    real* __restrict__ y_out = static_cast<real*>(a.y_out);
    real* __restrict__ y_dense = static_cast<real*>(a.y_dense);
    const bool dense = y_dense != nullptr;
    const bool raw = (a.flags & 1) != 0;
    const real rtol = real(a.rtol), atol = real(a.atol);

    // Lane-level work queue.  Conditions take different numbers of steps (sub-steps in the fast phases, different outlet
    // knots), so with a fixed condition per lane a warp idles a quarter of its lanes on average (measured: 23.6 of 32
    // active).  Instead every lane that finishes draws the next condition from a global counter (one atomic per warp and
    // round) and carries on inside the same loop; the grid is persistent (BS23_CTAS_PER_SM CTAs per SM).  A new
    // condition enters with a zero-length step: with h = 0 all stages evaluate f(t0, y0), so the ordinary step code
    // produces the FSAL slope k1 and no second copy of the right-hand side is needed.
    bool have = false, fresh = false, exhausted = false;
    int i = 0, kend = 0, kc = 0, nacc = 0, nrej = 0, nrhs = 0, status = 0, stiff_cap = 0;
    int ucopy = 0;   // which copy of the coefficient block the next right-hand side reads (CoefDup): warp-uniform, toggles
    double t = 0.0, t_final = 0.0, tk = 0.0, tk1 = 0.0, hprop = 0.0;
    real Tk = real(0), Tk1 = real(0), slope = real(0);
    float t_ahead = 0.f, T_ahead = 0.f;
    real y[NS];
#if PFR_BS23_SMEM_STATE
    // FSAL slope and the two running combinations of a step live in shared memory ([vector][species][thread], conflict-free):
    // 54 registers fewer, which is what lets a third CTA onto the SM
    extern __shared__ __align__(16) unsigned char bs_dyn[];
    real* const stv = reinterpret_cast<real*>(bs_dyn) + threadIdx.x;
#define K1(k) stv[(0 * NS + (k)) * BS23_BLOCK]
#define ACC(k) stv[(1 * NS + (k)) * BS23_BLOCK]
#define ER(k) stv[(2 * NS + (k)) * BS23_BLOCK]
#else
    real k1[NS], acc[NS], er[NS];
#define K1(k) k1[k]
#define ACC(k) acc[k]
#define ER(k) er[k]
#endif
#pragma unroll
    for (int k = 0; k < NS; k++) { y[k] = real(0); K1(k) = real(0); }

    while (true) {
        {
            // (a plain atomicAdd on a warp-uniform address: ptxas aggregates it itself -- one REDUX + one atomic per warp and
            // round, consecutive slots to the requesting lanes in lane order.  Written out by hand with __ballot_sync /
            // __shfl_sync, the compiler has to allow for a diverged warp at each of those, and with that every uniform-register
            // load in the rest of the loop is lost.)
            if (!have && !exhausted) {
                const int slot = atomicAdd(a.work_counter, 1);
                if (slot >= a.n) {
                    exhausted = true;
                } else {
                    i = a.perm ? a.perm[slot] : slot;
                    kend = a.idx_end ? a.idx_end[i] : NTOT - 1;
#pragma unroll
                    for (int k = 0; k < NS; k++) { y[k] = real(0); K1(k) = real(0); }   // (k1 is multiplied by h = 0 in the entry step)
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
                    // knot k + 2 is fetched on arrival at knot k and used one interval later (hides the gathered-load latency)
                    const size_t k2i = NTOT > 2 ? 2 : NTOT - 1;
                    t_ahead = a.tgrid[k2i * n + i];
                    T_ahead = kRamp ? a.Tprof[k2i * n + i] : 0.f;
                    nacc = nrej = nrhs = status = 0;
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
        if (__reduce_or_sync(0xffffffffu, (unsigned)have) == 0u) break;   // (REDUX writes a uniform register: a uniform loop exit)
        bool done = have && !fresh && kc < 0;
        const bool active = have && !done;
        // The stage loop is NOT under `if (active)`: uniform-register loads (the second coefficient feed, CoefDup) can only be issued
        // from code in which the whole warp is known to be converged.  A lane without work evaluates the right-hand side on whatever
        // its registers hold -- the warp executes those instructions anyway -- and takes no part in what follows the loop.
        {
            const double dist = tk1 - t;
            const bool clip = !fresh && hprop * 1.01 >= dist;
            const double hs = fresh ? 0.0 : (clip ? dist : hprop);
            const real h = real(hs);
            const real tau = real(t - tk);   // time since the knot: T(t + c h) = Tk + slope (tau + c h)
            // The three stages run as a rolled loop so that the right-hand side exists once in the instruction stream (three
            // inlined copies made the loop body ~59 KB and instruction fetch the largest stall, 22 % of all stall cycles).
            // The stage index is warp-uniform.  acc / er accumulate the solution and error combinations as the slopes arrive:
            //   y1 = y + h (2/9 k1 + 1/3 k2 + 4/9 k3),   err = h (-5/72 k1 + 1/12 k2 + 1/9 k3 - 1/8 k4)
            real k2[NS], w[NS];
#pragma unroll
            for (int k = 0; k < NS; k++) w[k] = fma(real(0.5) * h, K1(k), y[k]);
            const real T_last = clip ? Tk1 : fma(slope, h + tau, Tk);   // the stage at t + h: exactly the knot value on a clipped step
#pragma unroll 1
            for (int stage = 0; stage < 3; stage++) {
                const real cs = stage == 0 ? real(0.5) : real(0.75);
                const real Ts = kRamp ? (stage == 2 ? T_last : fma(slope, fma(cs, h, tau), Tk)) : Tk;
                // `ucopy` is its own loop-carried variable on purpose: derived from `stage` it would share that variable's vector
                // register (stage is compared against per-thread values), and an index in a vector register turns every
                // uniform load into a per-thread constant load
                rhs_tpc<real, PFR_UNIFORM_ROWS>(p, sc, cd, ucopy, zthr, dthr, Ts, w, k2);
                ucopy ^= 1;
                if (stage == 0) {
#pragma unroll
                    for (int k = 0; k < NS; k++) {
                        const real k1k = K1(k);
                        ACC(k) = fma(real(1.0 / 3.0), k2[k], real(2.0 / 9.0) * k1k);
                        ER(k) = fma(real(1.0 / 12.0), k2[k], real(-5.0 / 72.0) * k1k);
                        w[k] = fma(real(0.75) * h, k2[k], y[k]);
                    }
                } else if (stage == 1) {
#pragma unroll
                    for (int k = 0; k < NS; k++) {
                        w[k] = fma(h, fma(real(4.0 / 9.0), k2[k], ACC(k)), y[k]);   // third-order solution y1
                        ER(k) = fma(real(1.0 / 9.0), k2[k], ER(k));
                    }
                }
            }   // k2 now holds k4 = f(t + h, y1): the next step's k1
          if (active) {
            nrhs += 3;
            real e2 = real(0);
            // |w_k| < 1e30 and not NaN for every species, on the integer pipe (any_outside: two 3-input max chains over the high words)
            bool finite = !any_outside(w, bound_key(real(1e30)));
#pragma unroll
            for (int k = 0; k < NS; k++) {
                const real ek = h * fma(real(-1.0 / 8.0), k2[k], ER(k));
#if PFR_NORM_NEW_ONLY
                const real isk = rcp_norm(atol + rtol * m_abs(w[k]));
#else
                const real isk = rcp_norm(atol + rtol * m_max(m_abs(y[k]), m_abs(w[k])));
#endif
                e2 = fma(ek * isk, ek * isk, e2);
            }
            real err = m_sqrt<real>(e2 / real(NS));
#if PFR_BS23_KINK
            {   // A step during which a species crosses the lower state clamp straddles a kink of the right-hand side
                // (ln max(y, lb)); the embedded estimate, built for smooth f, under-reports the error there.  Such a step has
                // to meet the tolerance with a margin, which shrinks it until the kink is crossed with a small h.
                bool crossed = false;
#pragma unroll
                for (int k = 0; k < NS; k++) crossed = crossed || ((y[k] < p.lb) != (w[k] < p.lb));
                if (crossed) err *= real(PFR_BS23_KINK);
            }
#endif
            finite = finite && (err == err) && (err < real(1e30));
            const float fac = 0.9f / cbrtf(fmaxf((float)err, 1e-30f));
            if (fresh) {
                // the zero-length step left y unchanged and k2 = f(t0, y0); Hairer-style first step from |y0| and |f0|
#pragma unroll
                for (int k = 0; k < NS; k++) K1(k) = k2[k];
                real d0 = real(0), d1 = real(0);   // (once per trajectory: kept out of the per-step norm, 36 FP64 instructions)
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    const real isk = rcp_norm(atol + rtol * m_abs(y[k]));   // (w = y after the zero-length step)
                    d0 = fma(y[k] * isk, y[k] * isk, d0);
                    d1 = fma(k2[k] * isk, k2[k] * isk, d1);
                }
                d0 = m_sqrt<real>(d0 / real(NS));
                d1 = m_sqrt<real>(d1 / real(NS));
                const double h0 = (d0 < real(1e-5) || d1 < real(1e-5)) ? 1e-6 : 0.01 * (double)d0 / (double)d1;
                hprop = fmin(100.0 * h0, t_final - t);
                fresh = false;
            } else if (finite && err <= real(1)) {
                const double f = fmin(6.0, fmax(0.2, (double)fac));
                hprop = clip ? fmax(hprop, hs * f) : hs * f;
                nacc++;
#pragma unroll
                for (int k = 0; k < NS; k++) { y[k] = w[k]; K1(k) = k2[k]; }
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
#undef K1
#undef ACC
#undef ER

// ------------------------------------------------------------------------------------------------------------------------
// Free-stepping explicit integrator for the ISOTHERMAL sweep (Eoff: T = T0, integrate to t_end, no knots): Dormand-Prince 5(4)
// with FSAL, one condition per thread, same work queue / coefficient broadcast / clamp tests as bs23_kernel.
//
// The isothermal problem is as non-stiff as the coupled one (the reference's dopri5 takes 3-36 accepted steps for it), and
// here the error -- not a knot grid -- sets the step, so a 5th-order pair pays: ~100 right-hand sides per trajectory
// against RODAS4's ~180 plus 30 Jacobians and inversions, and a warp instruction serves 32 conditions instead of 10.
// The seven stage slopes live in shared memory ([stage][species][thread], conflict-free), which lets the stages run as a
// ROLLED loop over the tableau (one copy of the right-hand side in the instruction stream) without indexing registers
// dynamically.  kT_j is constant along a trajectory and is hoisted.  The controller follows the Rosenbrock kernels (RMS
// norm against atol + rtol max(|y0|, |y1|), factor clip(0.9 err^(-1/5), 0.2, 6), last step clipped to t_end); a condition
// that needs more than DP54_MAX_ATTEMPTS attempts stops with PFR_ST_STIFF and is handed to RODAS4 by the host.
#ifndef PFR_DP54_KINK
#define PFR_DP54_KINK 1000   // as PFR_BS23_KINK; free steps are long, the margin matters more: at 1e-7 the p99 error falls 500x
#endif
#ifndef PFR_DP54_MINB
#define PFR_DP54_MINB 3   // CTAs per SM: 168 registers, no spills (the slopes live in shared memory), 12 warps / SM; 2 CTAs: 11 % slower
#endif
constexpr int DP54_BLOCK = 128, DP54_CTAS_PER_SM = PFR_DP54_MINB, DP54_STAGES = 7, DP54_MAX_ATTEMPTS = 2000;
template <typename real> constexpr size_t dp54_smem_bytes() { return (size_t)DP54_STAGES * NS * DP54_BLOCK * sizeof(real); }

struct Dp54Tableau {
    double a[DP54_STAGES][DP54_STAGES - 1];   // row s: coefficients of k_0 .. k_{s-1}; row 6 = the 5th-order weights (FSAL)
    double e[DP54_STAGES];                    // error weights b - b*
};
__constant__ Dp54Tableau c_dp54 = {
    {{0, 0, 0, 0, 0, 0},
     {1.0 / 5, 0, 0, 0, 0, 0},
     {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
     {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
     {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
     {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
     {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}},
    {35.0 / 384 - 1951.0 / 21600, 0, 500.0 / 1113 - 22642.0 / 50085, 125.0 / 192 - 451.0 / 720, -2187.0 / 6784 + 12231.0 / 42400,
     11.0 / 84 - 649.0 / 6300, -1.0 / 60}};

template <typename real>
__global__ void __launch_bounds__(DP54_BLOCK, DP54_CTAS_PER_SM)
dp54_kernel(const __grid_constant__ CrnnParams<real> p, const __grid_constant__ CoefDup<real> cd, const RodasArgs a) {
    __shared__ __align__(16) TpcCoef<real> sc;
    extern __shared__ __align__(16) unsigned char dp_dyn[];
    real* const ks = reinterpret_cast<real*>(dp_dyn) + threadIdx.x;   // slope k_s of species i: ks[(s * NS + i) * DP54_BLOCK]
    load_tpc_coef<real, DP54_BLOCK, true>(sc, p, a.tables);
    const int zthr = bound_key(z_scale<real>() * m_min(m_abs(p.zlo), m_abs(p.zhi))), dthr = bound_key(m_min(m_abs(p.dulo), m_abs(p.duhi)));
    const size_t n = (size_t)a.n;
    real* __restrict__ y_out = static_cast<real*>(a.y_out);
    const real rtol = real(a.rtol), atol = real(a.atol);

    bool have = false, fresh = false, exhausted = false;
    int i = 0, nacc = 0, nrej = 0, nrhs = 0, status = 0;
    int ucopy = 0;   // copy of the coefficient block the next right-hand side reads (CoefDup): warp-uniform, toggles
    double t = 0.0, t_final = 0.0, hprop = 0.0;
    real y[NS], kT[NR];
#pragma unroll
    for (int k = 0; k < NS; k++) y[k] = real(0);
#pragma unroll
    for (int j = 0; j < NR; j++) kT[j] = real(0);

    while (true) {
        {
            // (a plain atomicAdd on a warp-uniform address: ptxas aggregates it itself -- one REDUX + one atomic per warp and
            // round, consecutive slots to the requesting lanes in lane order.  Written out by hand with __ballot_sync /
            // __shfl_sync, the compiler has to allow for a diverged warp at each of those, and with that every uniform-register
            // load in the rest of the loop is lost.)
            if (!have && !exhausted) {
                const int slot = atomicAdd(a.work_counter, 1);
                if (slot >= a.n) {
                    exhausted = true;
                } else {
                    i = a.perm ? a.perm[slot] : slot;
#pragma unroll
                    for (int k = 0; k < NS; k++) y[k] = real(0);
                    y[NS - 3] = real(a.c0[i]);
                    t = 0.0;
                    t_final = (double)a.t_end[i];
                    nacc = nrej = nrhs = status = 0;
                    hprop = 0.0;
                    have = true;
                    fresh = t_final > t;
                    if (fresh) arrhenius_tpc<real>(p, sc, real(a.T0[i]), kT);
                }
            }
        }
        if (__reduce_or_sync(0xffffffffu, (unsigned)have) == 0u) break;   // (REDUX writes a uniform register: a uniform loop exit)
        bool done = have && !fresh && !(t_final > t);
        const bool active = have && !done;
        // The stage loop has the same six rounds for every lane and is not under `if (active)`: uniform-register loads need a warp
        // that is known to be converged (see bs23_kernel).  A new condition enters with a zero-length step -- every stage then
        // evaluates f(y0), the last one leaves it where the first-step guess and the FSAL slot expect it (the warp runs six
        // rounds for its other lanes anyway); a lane without work computes on whatever its registers hold.
        {
            const double dist = t_final - t;
            const bool clip = !fresh && hprop * 1.01 >= dist;
            const double hs = fresh ? 0.0 : (clip ? dist : hprop);
            const real h = real(hs);
            real w[NS], f[NS];
#pragma unroll 1
            for (int s = 1; s < DP54_STAGES; s++) {
#pragma unroll
                for (int k = 0; k < NS; k++) w[k] = real(0);
                for (int j = 0; j < s; j++) {
                    const real c = real(c_dp54.a[s][j]);
#pragma unroll
                    for (int k = 0; k < NS; k++) w[k] = fma(c, ks[(j * NS + k) * DP54_BLOCK], w[k]);
                }
#pragma unroll
                for (int k = 0; k < NS; k++) w[k] = fresh ? y[k] : fma(h, w[k], y[k]);   // (not h * stale slopes: they may be non-finite)
                rhs_tpc_kT<real, PFR_UNIFORM_ROWS>(p, sc, cd, ucopy, zthr, dthr, kT, w, f);
                ucopy ^= 1;
#pragma unroll
                for (int k = 0; k < NS; k++) ks[(s * NS + k) * DP54_BLOCK] = f[k];
            }
          if (active) {
            nrhs += fresh ? 1 : DP54_STAGES - 1;
            // w = y1 (row 6 of the tableau = the 5th-order weights), f = f(y1)   [entry step: w = y0, f = f(y0), h = 0]
            real e2 = real(0);
            bool finite = !any_outside(w, bound_key(real(1e30)));   // (as in bs23_kernel)
#pragma unroll
            for (int k = 0; k < NS; k++) {
                real ek = real(0);
#pragma unroll
                for (int j = 0; j < DP54_STAGES; j++) ek = fma(real(c_dp54.e[j]), ks[(j * NS + k) * DP54_BLOCK], ek);
                ek *= h;
                const real isk = rcp_norm(atol + rtol * m_max(m_abs(y[k]), m_abs(w[k])));
                e2 = fma(ek * isk, ek * isk, e2);
            }
            real err = m_sqrt<real>(e2 / real(NS));
            // A step during which a species crosses the lower state clamp straddles a kink of the right-hand side
            // (ln max(y, lb)); the embedded estimate, built for smooth f, under-reports the error there.  Such a step has to
            // meet the tolerance with a margin of PFR_DP54_KINK, which shrinks it until the kink is crossed with a small h.
            bool crossed = false;
#pragma unroll
            for (int k = 0; k < NS; k++) crossed = crossed || ((y[k] < p.lb) != (w[k] < p.lb));
            if (crossed) err *= real(PFR_DP54_KINK);
            finite = finite && (err == err) && (err < real(1e30));
            const float fac = 0.9f * __powf(fmaxf((float)err, 1e-30f), -0.2f);
            if (fresh) {
                real d0 = real(0), d1 = real(0);   // (once per trajectory: kept out of the per-step norm)
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    const real isk = rcp_norm(atol + rtol * m_abs(y[k]));   // (w = y on the entry step)
                    d0 = fma(y[k] * isk, y[k] * isk, d0);
                    d1 = fma(f[k] * isk, f[k] * isk, d1);
                }
                d0 = m_sqrt<real>(d0 / real(NS));
                d1 = m_sqrt<real>(d1 / real(NS));
                const double h0 = (d0 < real(1e-5) || d1 < real(1e-5)) ? 1e-6 : 0.01 * (double)d0 / (double)d1;
                hprop = fmin(100.0 * h0, t_final - t);
                fresh = false;
#pragma unroll
                for (int k = 0; k < NS; k++) ks[k * DP54_BLOCK] = f[k];
            } else if (finite && err <= real(1)) {
                const double g = fmin(6.0, fmax(0.2, (double)fac));
                hprop = clip ? fmax(hprop, hs * g) : hs * g;
                nacc++;
#pragma unroll
                for (int k = 0; k < NS; k++) { y[k] = w[k]; ks[k * DP54_BLOCK] = f[k]; }
                if (clip) { t = t_final; done = true; } else { t += hs; }
            } else {
                nrej++;
                const double g = finite ? fmax(0.2, (double)fac) : 0.2;
                hprop = hs * fmin(g, 0.9);
                if (!(t + hprop > t) || hprop < 1e-300) { status = finite ? PFR_ST_UNDERFLOW_ : PFR_ST_NONFINITE_; done = true; }
            }
            if (!done && nacc + nrej > DP54_MAX_ATTEMPTS) { status = PFR_ST_STIFF_; done = true; }
            if (!done && nacc + nrej >= a.max_steps) { status = PFR_ST_MAXSTEPS_; done = true; }
          }
        }
        if (done) {
            const int io = a.out_index ? a.out_index[i] : i;   // column of the results (caller's order; pfr_sweep_run)
#pragma unroll
            for (int k = 0; k < NS; k++) y_out[(size_t)k * n + io] = m_min(m_max(y[k], p.lb), p.ub);
            a.status[io] = status;
            if (a.stats) {
                a.stats[io] = nacc;
                a.stats[n + io] = nrej;
                a.stats[2 * n + io] = nrhs;
            }
            have = false;
        }
    }
}

}  // namespace pfr
