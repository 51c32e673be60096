// bs23_lanes_kernel: the explicit knot-limited integrator of integrate_explicit.cuh (Bogacki-Shampine 3(2), FSAL, same
// controller, same stiffness guard) mapped ONE CONDITION PER WARP, lane k = species k = reaction k.
//
// Why a second mapping.  bs23_kernel (one condition per thread) is built for throughput: a million conditions, three warps per
// scheduler hiding each other's latencies.  The training step integrates a few hundred conditions (640 in
// SURROGATE_MODEL_TRAINING/WIDE_Eoff_surrogate_model_training.py:398-422) with dense output at all 801 knots: one thread then
// walks ~2 800 instructions per step with nobody to overlap them with, and the forward pass is pure latency (3.5 ms for 640
// conditions, 4.3 us per step).  Here the nine logarithms, the nine exponentials and the rows of both mat-vecs of a right-hand
// side run side by side in nine lanes (vectors travel by warp shuffle, dot products as three partial sums), every stage
// combination is one FMA per lane, and the error norm is a five-step butterfly: the dependent chain of a step is ~7x shorter.
//
// Same arithmetic per component as bs23_kernel up to the summation order of the dot products (results agree to ~1e-13, step
// sequences are identical in practice); stats[2] counts 3 right-hand sides per attempt + 1 for the inlet slope.
#pragma once
#include "crnn_device.cuh"
#include "fastmath.cuh"
#include "integrate_explicit.cuh"

namespace pfr {

constexpr int LANES_WARPS = 4;   // conditions (warps) per block

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool kRamp>
__global__ void __launch_bounds__(32 * LANES_WARPS)
bs23_lanes_kernel(const __grid_constant__ CrnnParams<double> p, const RodasArgs a) {
    __shared__ __align__(16) FastTables ft;
    __shared__ CrnnParams<double> sp;   // block-shared copy: the lanes read their own rows / columns from it once
    for (int e = threadIdx.x; e < LOGTAB_N; e += 32 * LANES_WARPS) ft.logtab[e] = a.tables->logtab[e];
    for (int e = threadIdx.x; e < EXPTAB_N; e += 32 * LANES_WARPS) ft.exptab[e] = a.tables->exptab[e];
    for (int e = threadIdx.x; e < NS * NR; e += 32 * LANES_WARPS) {
        sp.nu[e / NR][e % NR] = p.nu[e / NR][e % NR];
        sp.wout[e / NR][e % NR] = p.wout[e / NR][e % NR];
    }
    if (threadIdx.x < NR) { sp.Ea[threadIdx.x] = p.Ea[threadIdx.x]; sp.b[threadIdx.x] = p.b[threadIdx.x]; sp.lnA[threadIdx.x] = p.lnA[threadIdx.x]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * LANES_WARPS + warp;
    if (slot >= a.n) return;   // warp-uniform
    const int i = a.perm ? a.perm[slot] : slot;
    const size_t n = (size_t)a.n;
    const bool sp_lane = lane < NS;
    const int k = sp_lane ? lane : 0;   // lanes 9..31 shadow lane 0; they never write and add zero to the norms
    double nu_col[NS], wout_row[NR];
#pragma unroll
    for (int r = 0; r < NS; r++) { nu_col[r] = sp.nu[r][k]; wout_row[r] = sp.wout[k][r]; }
    const double Ea = sp.Ea[k], bb = sp.b[k], lnA = sp.lnA[k];
    double* __restrict__ y_out = static_cast<double*>(a.y_out);
    double* __restrict__ y_dense = static_cast<double*>(a.y_dense);
    const bool dense = y_dense != nullptr, raw = (a.flags & 1) != 0;
    const double rtol = a.rtol, atol = a.atol;

    // f_k(T, w): every lane passes its own component of the stage argument and receives its own component of the slope
    auto arrhenius = [&](double T) { return fma(Ea, -p.inv_R * rcp_full(T), fma(bb, fast_log_ilp(T, ft.logtab), lnA)); };
    auto rhs = [&](double kT, double w) -> double {
        const double l = fast_log_ilp(m_min(m_max(w, p.lb), p.ub), ft.logtab);
        double z0 = kT, z1 = 0.0, z2 = 0.0;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            z0 = fma(nu_col[r], __shfl_sync(0xffffffffu, l, r), z0);
            z1 = fma(nu_col[r + 3], __shfl_sync(0xffffffffu, l, r + 3), z1);
            z2 = fma(nu_col[r + 6], __shfl_sync(0xffffffffu, l, r + 6), z2);
        }
        const double rr = fast_exp_ilp(m_min(m_max((z0 + z1) + z2, p.zlo), p.zhi), ft.exptab);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            s0 = fma(wout_row[j], __shfl_sync(0xffffffffu, rr, j), s0);
            s1 = fma(wout_row[j + 3], __shfl_sync(0xffffffffu, rr, j + 3), s1);
            s2 = fma(wout_row[j + 6], __shfl_sync(0xffffffffu, rr, j + 6), s2);
        }
        return m_min(m_max((s0 + s1) + s2, p.dulo), p.duhi);
    };

    const int kend = a.idx_end ? a.idx_end[i] : NTOT - 1;
    double y = (lane == NS - 3) ? (double)a.c0[i] : 0.0;
    double t = (double)a.tgrid[i];
    const double t_final = (double)a.tgrid[(size_t)kend * n + i];
    int kc = 0, nacc = 0, nrej = 0, nrhs = 0, status = 0;
    double tk = t, tk1 = (double)a.tgrid[n + i];
    double Tk = (double)a.T0[i], Tk1 = Tk, slope = 0.0;
    if (kRamp) {
        Tk = (double)a.Tprof[i];
        Tk1 = (double)a.Tprof[n + i];
        slope = (Tk1 - Tk) * rcp_full(tk1 - tk);
    }
    const size_t k2i = 2;
    float t_ahead = a.tgrid[k2i * n + i], T_ahead = kRamp ? a.Tprof[k2i * n + i] : 0.f;
    if (dense && sp_lane) y_dense[(size_t)k * n + i] = raw ? y : m_min(m_max(y, p.lb), p.ub);
    const double kT_const = kRamp ? 0.0 : arrhenius(Tk);
    bool done = !(kend != 0 && t_final > t);
    double k1 = 0.0, hprop = 0.0;
    if (!done) {
        // inlet slope and Hairer-style first step from |y0| and |f0| (the same guess bs23_kernel makes with its zero-length step)
        k1 = rhs(kRamp ? arrhenius(Tk) : kT_const, y);
        nrhs = 1;
        const double isk = rcp_norm(atol + rtol * m_abs(y));
        const double d0 = sqrt(warp_sum(sp_lane ? (y * isk) * (y * isk) : 0.0) / NS), d1 = sqrt(warp_sum(sp_lane ? (k1 * isk) * (k1 * isk) : 0.0) / NS);
        const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        hprop = fmin(100.0 * h0, t_final - t);
    }
    const int stiff_cap = 8 * kend + 512;
    while (!done) {
        const double dist = tk1 - t;
        const bool clip = hprop * 1.01 >= dist;
        const double h = clip ? dist : hprop;
        const double tau = t - tk;
        // y1 = y + h (2/9 k1 + 1/3 k2 + 4/9 k3),   err = h (-5/72 k1 + 1/12 k2 + 1/9 k3 - 1/8 k4)
        const double k2 = rhs(kRamp ? arrhenius(fma(slope, fma(0.5, h, tau), Tk)) : kT_const, fma(0.5 * h, k1, y));
        const double k3 = rhs(kRamp ? arrhenius(fma(slope, fma(0.75, h, tau), Tk)) : kT_const, fma(0.75 * h, k2, y));
        const double y1 = sp_lane ? fma(h, fma(4.0 / 9.0, k3, fma(1.0 / 3.0, k2, (2.0 / 9.0) * k1)), y) : 0.0;   // (shadow lanes carry nothing)
        const double k4 = rhs(kRamp ? arrhenius(clip ? Tk1 : fma(slope, tau + h, Tk)) : kT_const, y1);
        nrhs += 3;
        const double ek = h * fma(-1.0 / 8.0, k4, fma(1.0 / 9.0, k3, fma(1.0 / 12.0, k2, (-5.0 / 72.0) * k1)));
        const double isk = rcp_norm(atol + rtol * m_abs(y1));   // (weights from |y1| alone, as bs23_kernel: PFR_NORM_NEW_ONLY)
        const double err = sqrt(warp_sum(sp_lane ? (ek * isk) * (ek * isk) : 0.0) / NS);
        const bool finite = __all_sync(0xffffffffu, m_abs(y1) < 1e30) && (err == err) && (err < 1e30);
        const float fac = 0.9f / cbrtf(fmaxf((float)err, 1e-30f));
        if (finite && err <= 1.0) {
            const double f = fmin(6.0, fmax(0.2, (double)fac));
            hprop = clip ? fmax(hprop, h * f) : h * f;
            nacc++;
            y = y1;
            k1 = k4;
            if (clip) {
                t = tk1;
                kc++;
                if (dense && sp_lane) y_dense[((size_t)kc * NS + k) * n + i] = raw ? y : m_min(m_max(y, p.lb), p.ub);
                if (kc >= kend) {
                    done = true;
                } else {
                    tk = tk1;
                    tk1 = (double)t_ahead;
                    const size_t kk = (size_t)(kc + 2 < NTOT ? kc + 2 : NTOT - 1);
                    t_ahead = a.tgrid[kk * n + i];
                    if (kRamp) {
                        Tk = Tk1;
                        Tk1 = (double)T_ahead;
                        slope = (Tk1 - Tk) * rcp_full(tk1 - tk);
                        T_ahead = a.Tprof[kk * n + i];
                    }
                }
            } else {
                t += h;
            }
        } else {
            nrej++;
            const double f = finite ? fmax(0.2, (double)fac) : 0.2;
            hprop = h * fmin(f, 0.9);
            if (!(t + hprop > t) || hprop < 1e-300) { status = finite ? PFR_ST_UNDERFLOW_ : PFR_ST_NONFINITE_; done = true; }
        }
        if (!done && nacc + nrej > stiff_cap) { status = PFR_ST_STIFF_; done = true; }
        if (!done && nacc + nrej >= a.max_steps) { status = PFR_ST_MAXSTEPS_; done = true; }
    }
    const int io = a.out_index ? a.out_index[i] : i;
    const double yf = m_min(m_max(y, p.lb), p.ub);
    if (sp_lane) y_out[(size_t)k * n + io] = yf;
    if (lane == 0) {
        a.status[io] = status;
        if (a.stats) {
            a.stats[io] = nacc;
            a.stats[n + io] = nrej;
            a.stats[2 * n + io] = nrhs;
        }
    }
    if (dense && sp_lane && kc < NTOT - 1)
        for (int kk = kc + 1; kk < NTOT; kk++) y_dense[((size_t)kk * NS + k) * n + i] = raw ? y : yf;
}

// ------------------------------------------------------------------------------------------------------------------------
// dp54_lanes_kernel: the ISOTHERMAL training forward pass (WIDE_Eoff / narrow Eoff trainers: T = T0, labels at the 801 knots of
// tgrid).  Nothing forces a step to end on a knot when the temperature is constant -- the reference's own dopri5 takes 3-36 steps
// for such a trajectory and INTERPOLATES its 801 outputs (torchdiffeq's dense output; ...training.py:383) -- so this kernel does
// the same: Dormand-Prince 5(4) with FSAL, free stepping to the last knot, and the knot states from the method's 4th-order
// continuous extension (Shampine's coefficients, the ones scipy's RK45 and torchdiffeq use).  ~25-60 steps of 6 right-hand sides
// instead of bs23_lanes_kernel's 800 steps of 3: the pass is a chain of dependent right-hand sides, so its time falls with their
// number.  Mapping as above: one condition per warp, lane k = species k = reaction k.  Controller and kink margin of dp54_kernel
// (RMS norm against atol + rtol max(|y0|, |y1|), a step that carries a species across the lower state clamp has to meet the
// tolerance with a margin of PFR_DP54_KINK); stats[2] counts right-hand sides (1 + 6 per attempt).
// The knot times of a condition are read 32 at a time (lane l reads knot base + l) and handed round by shuffle.
struct Dp54Dense {
    double p[DP54_STAGES][4];   // b_i(theta) = sum_j p[i][j] theta^j;  y(t + theta h) = y + h theta sum_i b_i(theta) k_i
};
__constant__ Dp54Dense c_dp54_dense = {
    {{1.0, -8048581381.0 / 2820520608.0, 8663915743.0 / 2820520608.0, -12715105075.0 / 11282082432.0},
     {0.0, 0.0, 0.0, 0.0},
     {0.0, 131558114200.0 / 32700410799.0, -68118460800.0 / 10900136933.0, 87487479700.0 / 32700410799.0},
     {0.0, -1754552775.0 / 470086768.0, 14199869525.0 / 1410260304.0, -10690763975.0 / 1880347072.0},
     {0.0, 127303824393.0 / 49829197408.0, -318862633887.0 / 49829197408.0, 701980252875.0 / 199316789632.0},
     {0.0, -282668133.0 / 205662961.0, 2019193451.0 / 616988883.0, -1453857185.0 / 822651844.0},
     {0.0, 40617522.0 / 29380423.0, -110615467.0 / 29380423.0, 69997945.0 / 29380423.0}}};

__global__ void __launch_bounds__(32 * LANES_WARPS)
dp54_lanes_kernel(const __grid_constant__ CrnnParams<double> p, const RodasArgs a) {
    __shared__ __align__(16) FastTables ft;
    __shared__ CrnnParams<double> sp;
    for (int e = threadIdx.x; e < LOGTAB_N; e += 32 * LANES_WARPS) ft.logtab[e] = a.tables->logtab[e];
    for (int e = threadIdx.x; e < EXPTAB_N; e += 32 * LANES_WARPS) ft.exptab[e] = a.tables->exptab[e];
    for (int e = threadIdx.x; e < NS * NR; e += 32 * LANES_WARPS) {
        sp.nu[e / NR][e % NR] = p.nu[e / NR][e % NR];
        sp.wout[e / NR][e % NR] = p.wout[e / NR][e % NR];
    }
    if (threadIdx.x < NR) { sp.Ea[threadIdx.x] = p.Ea[threadIdx.x]; sp.b[threadIdx.x] = p.b[threadIdx.x]; sp.lnA[threadIdx.x] = p.lnA[threadIdx.x]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * LANES_WARPS + warp;
    if (slot >= a.n) return;   // warp-uniform
    const int i = a.perm ? a.perm[slot] : slot;
    const size_t n = (size_t)a.n;
    const bool sp_lane = lane < NS;
    const int k = sp_lane ? lane : 0;
    double nu_col[NS], wout_row[NR];
#pragma unroll
    for (int r = 0; r < NS; r++) { nu_col[r] = sp.nu[r][k]; wout_row[r] = sp.wout[k][r]; }
    double* __restrict__ y_out = static_cast<double*>(a.y_out);
    double* __restrict__ y_dense = static_cast<double*>(a.y_dense);
    const bool dense = y_dense != nullptr, raw = (a.flags & 1) != 0;
    const double rtol = a.rtol, atol = a.atol;
    const double T0 = (double)a.T0[i];
    const double kT = fma(sp.Ea[k], -p.inv_R * rcp_full(T0), fma(sp.b[k], fast_log_ilp(T0, ft.logtab), sp.lnA[k]));

    auto rhs = [&](double w) -> double {
        const double l = fast_log_ilp(m_min(m_max(w, p.lb), p.ub), ft.logtab);
        double z0 = kT, z1 = 0.0, z2 = 0.0;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            z0 = fma(nu_col[r], __shfl_sync(0xffffffffu, l, r), z0);
            z1 = fma(nu_col[r + 3], __shfl_sync(0xffffffffu, l, r + 3), z1);
            z2 = fma(nu_col[r + 6], __shfl_sync(0xffffffffu, l, r + 6), z2);
        }
        const double rr = fast_exp_ilp(m_min(m_max((z0 + z1) + z2, p.zlo), p.zhi), ft.exptab);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            s0 = fma(wout_row[j], __shfl_sync(0xffffffffu, rr, j), s0);
            s1 = fma(wout_row[j + 3], __shfl_sync(0xffffffffu, rr, j + 3), s1);
            s2 = fma(wout_row[j + 6], __shfl_sync(0xffffffffu, rr, j + 6), s2);
        }
        return m_min(m_max((s0 + s1) + s2, p.dulo), p.duhi);
    };

    const int kend = a.idx_end ? a.idx_end[i] : NTOT - 1;
    double y = (lane == NS - 3) ? (double)a.c0[i] : 0.0;
    double t = (double)a.tgrid[i];
    const double t_final = (double)a.tgrid[(size_t)kend * n + i];
    int kc = 0, nacc = 0, nrej = 0, nrhs = 0, status = 0;
    if (dense && sp_lane) y_dense[(size_t)k * n + i] = raw ? y : m_min(m_max(y, p.lb), p.ub);
    // knot times, 32 at a time: tk_chunk of lane l = time of knot kbase + l (the next knot to be written is kc + 1)
    int kbase = 1;
    float tk_chunk = a.tgrid[(size_t)min(kbase + lane, NTOT - 1) * n + i];
    bool done = !(kend != 0 && t_final > t);
    double ks[DP54_STAGES];
#pragma unroll
    for (int s = 0; s < DP54_STAGES; s++) ks[s] = 0.0;
    double hprop = 0.0;
    if (!done) {
        ks[0] = rhs(y);
        nrhs = 1;
        const double isk = rcp_norm(atol + rtol * m_abs(y));
        const double d0 = sqrt(warp_sum(sp_lane ? (y * isk) * (y * isk) : 0.0) / NS), d1 = sqrt(warp_sum(sp_lane ? (ks[0] * isk) * (ks[0] * isk) : 0.0) / NS);
        const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        hprop = fmin(100.0 * h0, t_final - t);
    }
    while (!done) {
        const double dist = t_final - t;
        const bool clip = hprop * 1.01 >= dist;
        const double h = clip ? dist : hprop;
        double y1 = y;
#pragma unroll
        for (int s = 1; s < DP54_STAGES; s++) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < s; j++) acc = fma(c_dp54.a[s][j], ks[j], acc);
            y1 = sp_lane ? fma(h, acc, y) : 0.0;   // (shadow lanes carry nothing; row 6 = the 5th-order weights: y1 is the new state)
            ks[s] = rhs(y1);
        }
        nrhs += DP54_STAGES - 1;
        double ek = 0.0;
#pragma unroll
        for (int j = 0; j < DP54_STAGES; j++) ek = fma(c_dp54.e[j], ks[j], ek);
        ek *= h;
        const double isk = rcp_norm(atol + rtol * m_max(m_abs(y), m_abs(y1)));
        double err = sqrt(warp_sum(sp_lane ? (ek * isk) * (ek * isk) : 0.0) / NS);
        if (__any_sync(0xffffffffu, sp_lane && ((y < p.lb) != (y1 < p.lb)))) err *= (double)PFR_DP54_KINK;
        const bool finite = __all_sync(0xffffffffu, m_abs(y1) < 1e30) && (err == err) && (err < 1e30);
        const float fac = 0.9f * __powf(fmaxf((float)err, 1e-30f), -0.2f);
        if (finite && err <= 1.0) {
            const double g = fmin(6.0, fmax(0.2, (double)fac));
            hprop = clip ? fmax(hprop, h * g) : h * g;
            nacc++;
            const double t1 = clip ? t_final : t + h;
            if (dense) {
                // the knots in (t, t1]: y(t + theta h) = y + h theta (c0 + theta (c1 + theta (c2 + theta c3))), c_j = sum_i p[i][j] k_i
                double c[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    double v = 0.0;
#pragma unroll
                    for (int s = 0; s < DP54_STAGES; s++)
                        if (s != 1) v = fma(c_dp54_dense.p[s][j], ks[s], v);
                    c[j] = v;
                }
                const double rh = rcp_full(h);
                while (kc < kend) {
                    const double tk = (double)__shfl_sync(0xffffffffu, tk_chunk, kc + 1 - kbase);
                    if (tk > t1) break;
                    const double th = (tk - t) * rh;
                    const double yk = (tk == t1) ? y1 : fma(h * th, fma(th, fma(th, fma(th, c[3], c[2]), c[1]), c[0]), y);
                    kc++;
                    if (sp_lane) y_dense[((size_t)kc * NS + k) * n + i] = raw ? yk : m_min(m_max(yk, p.lb), p.ub);
                    if (kc + 1 - kbase == 32) {
                        kbase += 32;
                        tk_chunk = a.tgrid[(size_t)min(kbase + lane, NTOT - 1) * n + i];
                    }
                }
            }
            y = y1;
            ks[0] = ks[DP54_STAGES - 1];
            t = t1;
            if (clip) { kc = kend; done = true; }
        } else {
            nrej++;
            const double g = finite ? fmax(0.2, (double)fac) : 0.2;
            hprop = h * fmin(g, 0.9);
            if (!(t + hprop > t) || hprop < 1e-300) { status = finite ? PFR_ST_UNDERFLOW_ : PFR_ST_NONFINITE_; done = true; }
        }
        if (!done && nacc + nrej > DP54_MAX_ATTEMPTS) { status = PFR_ST_STIFF_; done = true; }
        if (!done && nacc + nrej >= a.max_steps) { status = PFR_ST_MAXSTEPS_; done = true; }
    }
    const int io = a.out_index ? a.out_index[i] : i;
    const double yf = m_min(m_max(y, p.lb), p.ub);
    if (sp_lane) y_out[(size_t)k * n + io] = yf;
    if (lane == 0) {
        a.status[io] = status;
        if (a.stats) {
            a.stats[io] = nacc;
            a.stats[n + io] = nrej;
            a.stats[2 * n + io] = nrhs;
        }
    }
    if (dense && sp_lane && kc < NTOT - 1)
        for (int kk = kc + 1; kk < NTOT; kk++) y_dense[((size_t)kk * NS + k) * n + i] = raw ? y : yf;
}

}  // namespace pfr
