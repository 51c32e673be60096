// RODAS4 with three cooperating lanes per PFR condition (ten conditions per warp).
//
// Same method, same controller and same results as rodas4_kernel (integrate_rodas.cuh); different mapping.
// Index i of every 9-vector (species, reactions, matrix rows) belongs to lane i % 3 of the group and sits in
// slot i / 3 of that lane's registers, so a lane holds 3 species, 3 reactions and 3 rows of E = I/(h gamma) - J.
// Nothing per-condition lives in shared memory:
//   * the three rows of E are inverted in place by a distributed Gauss-Jordan sweep (pivot row broadcast by
//     warp shuffles, no pivoting), after which each of the six stage solves is a 3x9 mat-vec
//     with short dependency chains instead of a serial triangular substitution;
//   * the right-hand side needs two 9-element all-gathers (ln Y and the rates r), also by shuffles.
// Per lane that is ~90 live doubles instead of ~250, so three to four times as many warps fit on an SM and each
// lane still has three independent log/exp chains in flight; the unrolled loop body is a third as long.
// Control state (t, h, knot cursor) is replicated in the three lanes and computed from bit-identical inputs.
#pragma once
#include "crnn_device.cuh"
#include "fastmath.cuh"
#include "integrate_rodas.cuh"

namespace pfr {

constexpr int COOP_BLOCK = 128;
constexpr int COOP_PER_WARP = 10;
constexpr int COOP_PER_BLOCK = COOP_PER_WARP * (COOP_BLOCK / 32);
#ifndef PFR_COOP_MINB
#define PFR_COOP_MINB 3
#endif
#ifndef PFR_DOT2
#define PFR_DOT2 0   // 1: dot products on two accumulators
#endif
#ifndef PFR_AINV_SMEM
#define PFR_AINV_SMEM 0   // 1: park the inverse in shared memory for the six solves
#endif
constexpr unsigned FULL = 0xffffffffu;

constexpr int PADN = 10;  // 9-vectors padded to 10 so that rows stay 16-byte aligned in shared memory
// Tuning switches, all measured on B200 at 2^18 LLNL Eon conditions (shipped setting first, 88.8 ms):
//   PFR_OPAQUE_PTR 0 volatile scalar parameter loads | 1 plain loads through an opaque pointer: the compiler merges /
//                    hoists them again and spills ~1 KB per thread
//   PFR_VECLD      0 coefficient read inside the FMA chain | 1 vector loads (spills) | 2 nine scalar loads first (no change)
//   PFR_DOT2       0 one accumulator per dot product | 1 two (no change)
//   PFR_AINV_SMEM  0 inverse stays in registers | 1 parked in shared memory (LSU-bound, 99 ms)
#ifndef PFR_OPAQUE_PTR
#define PFR_OPAQUE_PTR 0
#endif
#ifndef PFR_VECLD
#define PFR_VECLD 0
#endif

// Block-shared copy of the CRNN parameters, re-laid out per lane: lane l, slot m owns reaction / species 3m+l and
// reads its nine coefficients as one contiguous row.
template <typename real>
struct CoopParams {
    real nuL[3][3][PADN];    // nuL[l][m][k]   = nu[k][3m+l]
    real woutL[3][3][PADN];  // woutL[l][m][j] = wout[3m+l][j]
    real Ea[NR], b[NR], lnA[NR];
    FastTables ft;  // log / exp tables (used by the double instantiation)
};

// Nine consecutive reals (a 16-byte aligned row) with vector loads
__device__ __forceinline__ void lds9(const double* q, double (&v)[NS]) {
    const double2* q2 = reinterpret_cast<const double2*>(q);
#pragma unroll
    for (int e = 0; e < 4; e++) { const double2 t = q2[e]; v[2 * e] = t.x; v[2 * e + 1] = t.y; }
    v[8] = q[8];
}
__device__ __forceinline__ void lds9(const float* q, float (&v)[NS]) {
    const float2* q2 = reinterpret_cast<const float2*>(q);
#pragma unroll
    for (int e = 0; e < 4; e++) { const float2 t = q2[e]; v[2 * e] = t.x; v[2 * e + 1] = t.y; }
    v[8] = q[8];
}

// log / exp of the hot loop: table-driven in double (fastmath.cuh), CUDA's logf / expf in float
template <typename real> __device__ __forceinline__ real c_log(real x, const CoopParams<real>& sp);
template <> __device__ __forceinline__ double c_log<double>(double x, const CoopParams<double>& sp) { return fast_log(x, sp.ft.logtab); }
template <> __device__ __forceinline__ float c_log<float>(float x, const CoopParams<float>&) { return logf(x); }
template <typename real> __device__ __forceinline__ real c_exp(real x, const CoopParams<real>& sp);
template <> __device__ __forceinline__ double c_exp<double>(double x, const CoopParams<double>& sp) { return fast_exp(x, sp.ft.exptab); }
template <> __device__ __forceinline__ float c_exp<float>(float x, const CoopParams<float>&) { return expf(x); }

// Parameter reads from the block-shared copy.  volatile: the values are loop-invariant, and without it the
// compiler hoists all 54 of a lane's coefficients out of the step loop into registers (and then spills).
template <typename real>
__device__ __forceinline__ real ldp(const real* q) {
#if PFR_OPAQUE_PTR
    return *q;
#else
    return *reinterpret_cast<const volatile real*>(q);
#endif
}

template <typename real>
__device__ __forceinline__ void gather9(const real (&own)[3], real (&all)[NS], int base) {
#pragma unroll
    for (int m = 0; m < 3; m++)
#pragma unroll
        for (int s = 0; s < 3; s++) all[3 * m + s] = __shfl_sync(FULL, own[m], base + s);
}

// sum of one value per lane over the three lanes of a group, in a fixed order (bit-identical in every lane)
template <typename real>
__device__ __forceinline__ real group_sum(real v, int base) {
    const real a = __shfl_sync(FULL, v, base), b = __shfl_sync(FULL, v, base + 1), c = __shfl_sync(FULL, v, base + 2);
    return (a + b) + c;
}

template <typename real> __device__ __forceinline__ real fast_rcp(real x);
template <> __device__ __forceinline__ float fast_rcp<float>(float x) { return __frcp_rn(x); }
template <> __device__ __forceinline__ double fast_rcp<double>(double x) {
    // MUFU.RCP64H seed + two Newton steps: full double precision for the normal, well-scaled arguments met here
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// The parameter block is loop-invariant and read by every RHS evaluation.  Re-deriving its address from an opaque
// zero keeps the compiler from merging the reads of different evaluations (which would pin a lane's 54
// coefficients in registers for the whole step) while leaving it free to schedule / vectorise them inside one.
template <typename real>
__device__ __forceinline__ const CoopParams<real>& opaque(const CoopParams<real>& sp) {
#if PFR_OPAQUE_PTR
    unsigned z;
    asm volatile("mov.u32 %0, 0;" : "=r"(z));
    return *reinterpret_cast<const CoopParams<real>*>(reinterpret_cast<const char*>(&sp) + z);
#else
    return sp;
#endif
}

template <typename real, bool WithDeriv>
__device__ __forceinline__ void coop_arrhenius(const CoopParams<real>& sp_, real inv_R, int l, real T, real (&kT)[3], real& mE_out,
                                               real& invT_out) {
    const CoopParams<real>& sp = opaque(sp_);
    const real invT = fast_rcp<real>(T);
    const real mE = -inv_R * invT;
    const real lnT = c_log<real>(T, sp);
#pragma unroll
    for (int m = 0; m < 3; m++) {
        const int j = 3 * m + l;
        kT[m] = fma(ldp(&sp.Ea[j]), mE, fma(ldp(&sp.b[j]), lnT, ldp(&sp.lnA[j])));
    }
    mE_out = mE;
    invT_out = invT;
}

// s = init + sum_k a_k b_k, optionally on two accumulators (shorter dependent-FMA chain)
#define PFR_DOT9(s_, init_, A_, B_)                                                  \
    do {                                                                             \
        if (PFR_DOT2) {                                                              \
            real s0_ = fma(A_(0), B_(0), init_), s1_ = A_(1) * B_(1);                \
            s0_ = fma(A_(2), B_(2), s0_); s1_ = fma(A_(3), B_(3), s1_);              \
            s0_ = fma(A_(4), B_(4), s0_); s1_ = fma(A_(5), B_(5), s1_);              \
            s0_ = fma(A_(6), B_(6), s0_); s1_ = fma(A_(7), B_(7), s1_);              \
            s_ = fma(A_(8), B_(8), s0_) + s1_;                                       \
        } else {                                                                     \
            real s0_ = init_;                                                        \
            _Pragma("unroll") for (int q_ = 0; q_ < 9; q_++) s0_ = fma(A_(q_), B_(q_), s0_); \
            s_ = s0_;                                                                \
        }                                                                            \
    } while (0)

// du_own = f(y) for the lane's three species.  KeepJac also returns g (masked rates, all nine) and q for the
// lane's species.
template <typename real, bool KeepJac>
__device__ __forceinline__ void coop_rhs(const CoopParams<real>& sp_, const CrnnParams<real>& p, int l, int base,
                                         const real (&kT)[3], const real (&yin)[3], real (&du)[3], real (&g_all)[NS],
                                         real (&q_own)[3]) {
    const CoopParams<real>& sp = opaque(sp_);
    real lnY[3];
#pragma unroll
    for (int m = 0; m < 3; m++) {
        const real Y = m_min(m_max(yin[m], p.lb), p.ub);
        lnY[m] = c_log<real>(Y, sp);
        if (KeepJac) q_own[m] = (yin[m] >= p.lb && yin[m] <= p.ub) ? fast_rcp<real>(Y) : real(0);
    }
    real lnY_all[NS];
    gather9<real>(lnY, lnY_all, base);
    real r[3], gm[3];
#pragma unroll
    for (int m = 0; m < 3; m++) {
        real z;
#if PFR_VECLD == 2
        real c9[NS];
#pragma unroll
        for (int k = 0; k < NS; k++) c9[k] = *reinterpret_cast<const volatile real*>(&sp.nuL[l][m][k]);
#define A_(k) c9[k]
#elif PFR_VECLD
        real c9[NS];
        lds9(sp.nuL[l][m], c9);
#define A_(k) c9[k]
#else
#define A_(k) ldp(&sp.nuL[l][m][k])
#endif
#define B_(k) lnY_all[k]
        PFR_DOT9(z, kT[m], A_, B_);
#undef A_
#undef B_
        r[m] = c_exp<real>(m_min(m_max(z, p.zlo), p.zhi), sp);
        if (KeepJac) gm[m] = (z >= p.zlo && z <= p.zhi) ? r[m] : real(0);
    }
    real r_all[NS];
    gather9<real>(r, r_all, base);
    if (KeepJac) gather9<real>(gm, g_all, base);
#pragma unroll
    for (int m = 0; m < 3; m++) {
        real s;
#if PFR_VECLD == 2
        real c9[NS];
#pragma unroll
        for (int j = 0; j < NR; j++) c9[j] = *reinterpret_cast<const volatile real*>(&sp.woutL[l][m][j]);
#define A_(j) c9[j]
#elif PFR_VECLD
        real c9[NS];
        lds9(sp.woutL[l][m], c9);
#define A_(j) c9[j]
#else
#define A_(j) ldp(&sp.woutL[l][m][j])
#endif
#define B_(j) r_all[j]
        PFR_DOT9(s, real(0), A_, B_);
#undef A_
#undef B_
        du[m] = m_min(m_max(s, p.dulo), p.duhi);
    }
}

// x_own = A_own * b, with b distributed like x
template <typename real>
__device__ __forceinline__ void coop_solve(const real (&A)[3][NS], const real* Ainv, real (&x)[3], int base) {
    real b_all[NS];
    gather9<real>(x, b_all, base);
#pragma unroll
    for (int m = 0; m < 3; m++) {
        real s;
#if PFR_AINV_SMEM
#define A_(k) ldp(Ainv + (m * NS + (k)) * COOP_BLOCK)
#else
#define A_(k) A[m][k]
#endif
#define B_(k) b_all[k]
        PFR_DOT9(s, real(0), A_, B_);
#undef A_
#undef B_
        x[m] = s;
    }
}

// ROS3 (Sandu, Verwer, Blom, Spee, Carmichael, Potra 1997): 3 stages, 2 right-hand sides, order 3(2), L-stable; the third
// stage re-uses the second stage's right-hand side.  Written in the same transformed form as RODAS4 above
// (E k_i = f(y + sum a_ij k_j) + sum (C_ij / h) k_j + h d_i f_t, E = I/(h gamma) - J).  Coefficients verified by an
// order-of-convergence experiment on the CRNN itself (tools/proto/ros_proto.py: global 3.0, embedded 2.0).
namespace ros3 {
constexpr double gamma = 0.43586652150845899941601945119356;
constexpr double C21 = -1.0156171083877702091975600115545;
constexpr double C31 = 4.0759956452537699824805835358067, C32 = 9.2076794298330791242156818474003;
constexpr double m2 = 6.1697947043828245592553615689730, m3 = -0.4277225654321857332623837380651;  // m1 = a21 = a31 = 1
constexpr double e1 = 0.5, e2 = -2.9079558716805469821718236208017, e3 = 0.2235406989781156962736090927619;
constexpr double c2 = gamma;
constexpr double d1 = gamma, d2 = 0.24291996454816804366592249683314, d3 = 2.1851380027664058511513169485832;
}  // namespace ros3

constexpr int COOP_RODAS4 = 0, COOP_ROS3 = 1;

// 0.9 err^(-1/(q+1)) with q the order of the embedded solution (3 for RODAS4, 2 for ROS3; the cube root in float: it only
// scales the next step, and the three lanes of a group evaluate it on bit-identical inputs)
template <int kMethod> __device__ __forceinline__ double ctrl_factor(double err) {
    if (kMethod == COOP_ROS3) return (double)(0.9f / cbrtf(fmaxf((float)err, 1e-30f)));
    return 0.9 / sqrt(sqrt(err));
}

template <typename real, bool kRamp, bool kKnots, int kMethod = COOP_RODAS4>
__global__ void __launch_bounds__(COOP_BLOCK, PFR_COOP_MINB)
rodas4_coop_kernel(const __grid_constant__ CrnnParams<real> p, const RodasArgs a) {
    using namespace rodas4;
    constexpr double kGamma = kMethod == COOP_ROS3 ? ros3::gamma : rodas4::gamma;
    constexpr double kD1 = kMethod == COOP_ROS3 ? ros3::d1 : rodas4::d1;
    static_assert(!kRamp || kKnots, "a temperature ramp needs knot-limited stepping");
    // stiff-fallback launch of the sweep pipeline: sized for the whole batch, its (usually empty) list counted on the device --
    // blocks beyond the list leave before they touch anything
    if (a.n_work && (long long)blockIdx.x * COOP_PER_BLOCK >= (long long)*a.n_work) return;
    __shared__ __align__(16) CoopParams<real> sp_block;
    CoopParams<real>& sp0 = sp_block;
#if PFR_AINV_SMEM
    extern __shared__ __align__(16) unsigned char coop_dyn[];  // [27][COOP_BLOCK] reals: each lane's rows of E^-1
    real* const Ainv = reinterpret_cast<real*>(coop_dyn) + threadIdx.x;
#else
    real* const Ainv = nullptr;
#endif
    for (int e = threadIdx.x; e < 3 * 3 * PADN; e += COOP_BLOCK) {
        const int ll = e / (3 * PADN), mm = (e / PADN) % 3, kk = e % PADN;
        sp0.nuL[ll][mm][kk] = kk < NS ? p.nu[kk][3 * mm + ll] : real(0);
        sp0.woutL[ll][mm][kk] = kk < NS ? p.wout[3 * mm + ll][kk] : real(0);
    }
    if (threadIdx.x < NR) {
        sp0.Ea[threadIdx.x] = p.Ea[threadIdx.x];
        sp0.b[threadIdx.x] = p.b[threadIdx.x];
        sp0.lnA[threadIdx.x] = p.lnA[threadIdx.x];
    }
    if (sizeof(real) == 8) {
        for (int e = threadIdx.x; e < LOGTAB_N; e += COOP_BLOCK) sp0.ft.logtab[e] = a.tables->logtab[e];
        for (int e = threadIdx.x; e < EXPTAB_N; e += COOP_BLOCK) sp0.ft.exptab[e] = a.tables->exptab[e];
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int grp = lane / 3, l = lane - 3 * (lane / 3);
    const bool mirror = grp == COOP_PER_WARP;  // lanes 30, 31 shadow lanes 27, 28 so that every shuffle has a full warp
    if (mirror) { grp = COOP_PER_WARP - 1; l = lane - 30; }
    const int base = 3 * grp;
    const int first = (blockIdx.x * (COOP_BLOCK / 32) + warp) * COOP_PER_WARP;
    const int nwork = a.n_work ? min(*a.n_work, a.n) : a.n;   // (device-side count: the stiff-fallback list of pfr_sweep_run)
    if (first >= nwork) return;  // warp-uniform
    const int slot = first + grp;
    const bool writer = !mirror && slot < nwork;
    const int i = a.perm ? a.perm[slot < nwork ? slot : nwork - 1] : (slot < nwork ? slot : nwork - 1);
    const int io = a.out_index ? a.out_index[i] : i;            // column of the results (caller's order)
    const size_t n = (size_t)a.n;
    real* __restrict__ y_out = static_cast<real*>(a.y_out);
    real* __restrict__ y_dense = static_cast<real*>(a.y_dense);
    const bool dense = kKnots && (y_dense != nullptr);

    real y[3];
#pragma unroll
    for (int m = 0; m < 3; m++) y[m] = (3 * m + l == NS - 3) ? real(a.c0[i]) : real(0);

    const int kend = (kKnots && a.idx_end) ? a.idx_end[i] : NTOT - 1;
    double t = kKnots ? (double)a.tgrid[i] : 0.0;
    const double t_final = kKnots ? (double)a.tgrid[(size_t)kend * n + i]
                                  : (a.t_end ? (double)a.t_end[i] : (double)a.tgrid[(size_t)(NTOT - 1) * n + i]);
    if (dense && writer) {
#pragma unroll
        for (int m = 0; m < 3; m++) y_dense[(size_t)(3 * m + l) * n + i] = (a.flags & 1) ? y[m] : m_min(m_max(y[m], p.lb), p.ub);
    }
    int kc = 0;
    double tk = t, tk1 = kKnots ? (double)a.tgrid[n + i] : t_final;
    real Tk = real(a.T0[i]), Tk1 = Tk, slope = real(0);
    if (kRamp) {
        Tk = real(a.Tprof[i]);
        Tk1 = real(a.Tprof[n + i]);
        slope = (Tk1 - Tk) / real(tk1 - tk);
    }
    // knot k + 2 is fetched when the lane arrives at knot k and used one knot interval later: the (gathered, L2 / DRAM)
    // load latency hides behind a whole step instead of stalling the group at every knot
    // (RODAS4 on a ramp is at its register limit and measured 2 % slower with the two extra live values: it loads at the knot)
    constexpr bool kAhead = kKnots && !(kMethod == COOP_RODAS4 && kRamp && sizeof(real) == 8);
    float t_ahead = 0.f, T_ahead = 0.f;
    if (kAhead) {
        const size_t k2 = NTOT > 2 ? 2 : NTOT - 1;
        t_ahead = a.tgrid[k2 * n + i];
        if (kRamp) T_ahead = a.Tprof[k2 * n + i];
    }
    real kT[3], mE, invT;
    if (!kRamp) coop_arrhenius<real, false>(sp0, p.inv_R, l, Tk, kT, mE, invT);

    const real rtol = real(a.rtol), atol = real(a.atol);
    int nacc = 0, nrej = 0, nrhs = 0, status = 0;
    double hprop = 0.0;
    bool first_step = true;
    bool done = (kend == 0) || !(t_final > t);

    while (__any_sync(FULL, !done)) {
        const CoopParams<real>& sp = opaque(sp_block);
        // ---------------- f0, df/dt, Jacobian rows at (t, y) ----------------
        real ak1[3], fx[3], A[3][NS];
        real g_all[NS], q_own[3];
        if (kRamp) coop_arrhenius<real, true>(sp, p.inv_R, l, Tk + slope * real(t - tk), kT, mE, invT);
        coop_rhs<real, true>(sp, p, l, base, kT, y, ak1, g_all, q_own);
        if (first_step) {
            real d0 = real(0), d1n = real(0);
#pragma unroll
            for (int m = 0; m < 3; m++) {
                const real isk = fast_rcp<real>(atol + rtol * m_abs(y[m]));
                d0 = fma(y[m] * isk, y[m] * isk, d0);
                d1n = fma(ak1[m] * isk, ak1[m] * isk, d1n);
            }
            d0 = m_sqrt<real>(group_sum<real>(d0, base) / real(NS));
            d1n = m_sqrt<real>(group_sum<real>(d1n, base) / real(NS));
            const double h0 = (d0 < real(1e-5) || d1n < real(1e-5)) ? 1e-6 : 0.01 * (double)d0 / (double)d1n;
            hprop = fmin(100.0 * h0, t_final - t);
            first_step = false;
        }
        const double dist = tk1 - t;
        const bool clip = hprop * 1.01 >= dist;
        double hs = clip ? dist : hprop;
        if (done) hs = 1.0;  // finished groups keep executing the warp's instructions on harmless numbers
        const real h = real(hs);
        const real ih = fast_rcp<real>(h);
        if (kRamp) {
            // df/dt = W_out (g o dkT/dT) dT/dt for the lane's species
            real gd[NR];
#pragma unroll
            for (int j = 0; j < NR; j++) gd[j] = g_all[j] * ((p.b[j] - p.Ea[j] * mE) * invT);
#pragma unroll
            for (int m = 0; m < 3; m++) {
                real s = real(0);
#pragma unroll
                for (int j = 0; j < NR; j++) s = fma(ldp(&sp.woutL[l][m][j]), gd[j], s);
                fx[m] = s * slope;
                ak1[m] = fma(h * real(kD1), fx[m], ak1[m]);
            }
        }
        {
            // rows of E for the lane's species: E[i][k] = delta_ik / (h gamma) - q_k sum_j (wout[i][j] g_j) nu[k][j]
            const real fac = real(1.0 / kGamma) * ih;
#pragma unroll
            for (int m = 0; m < 3; m++) {
                real G[NR];
#pragma unroll
                for (int j = 0; j < NR; j++) G[j] = ldp(&sp.woutL[l][m][j]) * g_all[j];
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    real s = real(0);
#pragma unroll
                    for (int j = 0; j < NR; j++) s = fma(G[j], p.nu[k][j], s);
                    A[m][k] = s;
                }
            }
#pragma unroll
            for (int k = 0; k < NS; k++) {
                const real qk = __shfl_sync(FULL, q_own[k / 3], base + k % 3);
#pragma unroll
                for (int m = 0; m < 3; m++) A[m][k] = fma(-qk, A[m][k], (3 * m + l == k) ? fac : real(0));
            }
        }
        // ---------------- in-place Gauss-Jordan inverse of E over the group's nine rows ----------------
        bool lu_ok = true;
#pragma unroll
        for (int c = 0; c < NS; c++) {
            const int oc = c % 3, mc = c / 3;
            real prow[NS];
#pragma unroll
            for (int k = 0; k < NS; k++) prow[k] = __shfl_sync(FULL, A[mc][k], base + oc);
            lu_ok = lu_ok && (m_abs(prow[c]) > real(1e-30));
            const real ip = fast_rcp<real>(prow[c]);
#pragma unroll
            for (int k = 0; k < NS; k++) prow[k] = (k == c) ? ip : prow[k] * ip;
            const bool own_pivot = (l == oc);
#pragma unroll
            for (int m = 0; m < 3; m++) {
                const real f = A[m][c];
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    const real upd = (k == c) ? -f * ip : fma(-f, prow[k], A[m][k]);
                    A[m][k] = (m == mc && own_pivot) ? prow[k] : upd;
                }
            }
        }

#if PFR_AINV_SMEM
#pragma unroll
        for (int m = 0; m < 3; m++)
#pragma unroll
            for (int k = 0; k < NS; k++) Ainv[(m * NS + k) * COOP_BLOCK] = A[m][k];
        asm volatile("" ::: "memory");  // the inverse leaves the register file here
#endif
        // ---------------- stages ----------------
        real er[3], ynew[3], dy[3], g_[NS], q_[3];
        coop_solve<real>(A, Ainv, ak1, base);
        if constexpr (kMethod == COOP_ROS3) {
            real ak2[3], ak3[3];
#pragma unroll
            for (int m = 0; m < 3; m++) ynew[m] = y[m] + ak1[m];
            if (kRamp) coop_arrhenius<real, false>(sp, p.inv_R, l, Tk + slope * real(t + ros3::c2 * hs - tk), kT, mE, invT);
            coop_rhs<real, false>(sp, p, l, base, kT, ynew, dy, g_, q_);
#pragma unroll
            for (int m = 0; m < 3; m++) {
                const real s2 = fma(real(ros3::C21) * ih, ak1[m], dy[m]);
                const real s3 = fma(real(ros3::C31) * ih, ak1[m], dy[m]);
                ak2[m] = kRamp ? fma(h * real(ros3::d2), fx[m], s2) : s2;
                ak3[m] = kRamp ? fma(h * real(ros3::d3), fx[m], s3) : s3;
            }
            coop_solve<real>(A, Ainv, ak2, base);
#pragma unroll
            for (int m = 0; m < 3; m++) ak3[m] = fma(real(ros3::C32) * ih, ak2[m], ak3[m]);
            coop_solve<real>(A, Ainv, ak3, base);
#pragma unroll
            for (int m = 0; m < 3; m++) {
                ynew[m] = fma(real(ros3::m3), ak3[m], fma(real(ros3::m2), ak2[m], ynew[m]));
                er[m] = fma(real(ros3::e3), ak3[m], fma(real(ros3::e2), ak2[m], real(ros3::e1) * ak1[m]));
            }
        } else {
            real ak2[3], ak3[3], ak4[3], ak5[3];

#pragma unroll
            for (int m = 0; m < 3; m++) ynew[m] = fma(real(a21), ak1[m], y[m]);
            if (kRamp) coop_arrhenius<real, false>(sp, p.inv_R, l, Tk + slope * real(t + c2 * hs - tk), kT, mE, invT);
            coop_rhs<real, false>(sp, p, l, base, kT, ynew, dy, g_, q_);
#pragma unroll
            for (int m = 0; m < 3; m++) {
                const real s = fma(real(C21) * ih, ak1[m], dy[m]);
                ak2[m] = kRamp ? fma(h * real(d2), fx[m], s) : s;
            }
            coop_solve<real>(A, Ainv, ak2, base);

#pragma unroll
            for (int m = 0; m < 3; m++) ynew[m] = fma(real(a32), ak2[m], fma(real(a31), ak1[m], y[m]));
            if (kRamp) coop_arrhenius<real, false>(sp, p.inv_R, l, Tk + slope * real(t + c3 * hs - tk), kT, mE, invT);
            coop_rhs<real, false>(sp, p, l, base, kT, ynew, dy, g_, q_);
#pragma unroll
            for (int m = 0; m < 3; m++) {
                const real s = fma(real(C31) * ih, ak1[m], fma(real(C32) * ih, ak2[m], dy[m]));
                ak3[m] = kRamp ? fma(h * real(d3), fx[m], s) : s;
            }
            coop_solve<real>(A, Ainv, ak3, base);

#pragma unroll
            for (int m = 0; m < 3; m++) ynew[m] = fma(real(a43), ak3[m], fma(real(a42), ak2[m], fma(real(a41), ak1[m], y[m])));
            if (kRamp) coop_arrhenius<real, false>(sp, p.inv_R, l, Tk + slope * real(t + c4 * hs - tk), kT, mE, invT);
            coop_rhs<real, false>(sp, p, l, base, kT, ynew, dy, g_, q_);
#pragma unroll
            for (int m = 0; m < 3; m++) {
                const real s = fma(real(C41) * ih, ak1[m], fma(real(C42) * ih, ak2[m], fma(real(C43) * ih, ak3[m], dy[m])));
                ak4[m] = kRamp ? fma(h * real(d4), fx[m], s) : s;
            }
            coop_solve<real>(A, Ainv, ak4, base);

#pragma unroll
            for (int m = 0; m < 3; m++)
                ynew[m] = fma(real(a54), ak4[m], fma(real(a53), ak3[m], fma(real(a52), ak2[m], fma(real(a51), ak1[m], y[m]))));
            if (kRamp) coop_arrhenius<real, false>(sp, p.inv_R, l, Tk + slope * real(t + hs - tk), kT, mE, invT);
            coop_rhs<real, false>(sp, p, l, base, kT, ynew, dy, g_, q_);
#pragma unroll
            for (int m = 0; m < 3; m++)
                ak5[m] = fma(real(C51) * ih, ak1[m], fma(real(C52) * ih, ak2[m], fma(real(C53) * ih, ak3[m],
                         fma(real(C54) * ih, ak4[m], dy[m]))));
            coop_solve<real>(A, Ainv, ak5, base);

#pragma unroll
            for (int m = 0; m < 3; m++) ynew[m] += ak5[m];  // embedded 3rd-order solution
            coop_rhs<real, false>(sp, p, l, base, kT, ynew, dy, g_, q_);
#pragma unroll
            for (int m = 0; m < 3; m++)
                er[m] = fma(real(C61) * ih, ak1[m], fma(real(C62) * ih, ak2[m], fma(real(C63) * ih, ak3[m],
                        fma(real(C64) * ih, ak4[m], fma(real(C65) * ih, ak5[m], dy[m])))));
            coop_solve<real>(A, Ainv, er, base);
#pragma unroll
            for (int m = 0; m < 3; m++) ynew[m] += er[m];
        }

        // ---------------- error estimate and step-size control (replicated, bit-identical in the 3 lanes) ----
        real e2 = real(0);
        bool fin_own = true;
#pragma unroll
        for (int m = 0; m < 3; m++) {
            const real sk = atol + rtol * m_max(m_abs(y[m]), m_abs(ynew[m]));
            const real w = er[m] * fast_rcp<real>(sk);
            e2 = fma(w, w, e2);
            fin_own = fin_own && (m_abs(ynew[m]) < real(1e30));
        }
        const real err = m_sqrt<real>(group_sum<real>(e2, base) / real(NS));
        const unsigned okmask = __ballot_sync(FULL, fin_own && lu_ok);
        bool finite = ((okmask >> base) & 7u) == 7u;
        finite = finite && (err == err) && (err < real(1e30));

        if (!done) {
            nrhs += kMethod == COOP_ROS3 ? 2 : 6;
            if (finite && err <= real(1)) {
                double f = err > real(0) ? ctrl_factor<kMethod>((double)err) : 6.0;
                f = fmin(6.0, fmax(0.2, f));
                hprop = clip ? fmax(hprop, hs * f) : hs * f;
                nacc++;
#pragma unroll
                for (int m = 0; m < 3; m++) y[m] = ynew[m];
                if (clip) {
                    t = tk1;
                    if (kKnots) {
                        kc++;
                        if (dense && writer) {
#pragma unroll
                            for (int m = 0; m < 3; m++)
                                y_dense[((size_t)kc * NS + 3 * m + l) * n + i] = (a.flags & 1) ? y[m] : m_min(m_max(y[m], p.lb), p.ub);
                        }
                        if (kc >= kend) {
                            done = true;
                        } else {
                            tk = tk1;
                            if (kAhead) {
                                tk1 = (double)t_ahead;
                                const size_t k2 = (size_t)(kc + 2 < NTOT ? kc + 2 : NTOT - 1);
                                t_ahead = a.tgrid[k2 * n + i];
                                if (kRamp) {
                                    Tk = Tk1;
                                    Tk1 = real(T_ahead);
                                    slope = (Tk1 - Tk) / real(tk1 - tk);
                                    T_ahead = a.Tprof[k2 * n + i];
                                }
                            } else {
                                tk1 = (double)a.tgrid[(size_t)(kc + 1) * n + i];
                                if (kRamp) {
                                    Tk = real(a.Tprof[(size_t)kc * n + i]);
                                    slope = (real(a.Tprof[(size_t)(kc + 1) * n + i]) - Tk) / real(tk1 - tk);
                                }
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
                const double f = finite ? fmax(0.2, ctrl_factor<kMethod>((double)err)) : 0.2;
                hprop = hs * fmin(f, 0.9);
                if (!(t + hprop > t) || hprop < 1e-300) { status = finite ? PFR_ST_UNDERFLOW_ : PFR_ST_NONFINITE_; done = true; }
            }
            if (!done && nacc + nrej >= a.max_steps) { status = PFR_ST_MAXSTEPS_; done = true; }
        }
    }

    if (!writer) return;
    real yf[3];
#pragma unroll
    for (int m = 0; m < 3; m++) {
        yf[m] = m_min(m_max(y[m], p.lb), p.ub);
        y_out[(size_t)(3 * m + l) * n + io] = yf[m];
    }
    if (l == 0) {
        a.status[io] = status;
        if (a.stats) {
            a.stats[io] = nacc;
            a.stats[n + io] = nrej;
            a.stats[2 * n + io] = nrhs;
        }
    }
    if (dense && kc < NTOT - 1) {
        for (int kk = kc + 1; kk < NTOT; kk++)
#pragma unroll
            for (int m = 0; m < 3; m++) y_dense[((size_t)kk * NS + 3 * m + l) * n + i] = (a.flags & 1) ? y[m] : yf[m];
    }
}

}  // namespace pfr
