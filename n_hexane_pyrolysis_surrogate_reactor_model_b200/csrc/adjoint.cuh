// Loss and parameter gradient of one CRNN training step, one condition per thread.
//
// Replaces  loss = Trainer.loss_n_ode(p, i_exp); loss.backward()
//   (SURROGATE_MODEL_TRAINING/WIDE_Eoff_surrogate_model_training.py:387-396, 414-416)
// for a whole batch of conditions.  The reference back-propagates through every operation of torchdiffeq's
// dopri5; here the gradient with respect to the CRNN parameters (w_in[11][9], w_b[9], w_out[9][9]) is the
// continuous adjoint of the same loss:
//     L      = mean_{i<7, k<=800} ((clamp(y_i(t_k)) - ref_ik) / yscale_i)^2
//     lam'   = -J(y,T)^T lam  between knots,   lam(t_k^-) = lam(t_k^+) + dL/dy(t_k),   lam(t_800^+) = 0
//     dL/dth = sum over knot intervals of  integral lam^T df/dth dt
// integrated backwards knot interval by knot interval with classical RK4 (`substeps` per interval); inside an
// interval y(t) is the cubic Hermite interpolant of the forward pass' knot states (the intervals are ~1/800 of
// the residence time, h*lambda <= 0.1) and T(t) is the same linear ramp the forward pass saw.
// df/dth contracts to outer products:  with lt_i = lam_i [du_i unclamped],  mu_j = [z_j unclamped] r_j sum_i lt_i wout_ij :
//     d/dwout_ij = lt_i r_j      d/dwb_j = mu_j      d/dwin_kj = mu_j wv_k      (J^T lam)_k = q_k sum_j nu_kj mu_j
// The 189 accumulators of a condition live in the thread's local memory (a training batch is hundreds of
// conditions, not millions; the arrays stay in L1).
#pragma once
#include "crnn_device.cuh"
#include "fastmath.cuh"

namespace pfr {

constexpr int NPAR = 11 * NR + NR + NS * NR;  // 189: w_in | w_b | w_out
constexpr int NOBS = 7;                       // observed species i_obs = 0..6
constexpr int ADJ_BLOCK = 64;

struct AdjointArgs {
    int n;
    const float* T0;        // [n]
    const float* tgrid;     // [801][n]
    const float* Tprof;     // [801][n] or nullptr (T = T0)
    const double* y_knots;  // [801][9][n] forward states at the knots, NOT clamped
    const float* ref;       // [801][7][n] labels (mol/m3)
    const float* yscale;    // [7][n]
    int substeps;
    double* loss;           // [n] per-condition MSE
    double* grad;           // [189][n] per-condition gradient
    const FastTables* tables = nullptr;   // adjoint_warp_kernel: device copy of the log / exp tables (fastmath.cuh)
};

struct AdjNode {
    double r[NR], mz[NR], wv[NS + 2], q[NS], md[NS];
};

// forward quantities of the RHS at (T, y): rates, masks, the input vector wv = [ln Y ; -1/(R T) ; ln T]
__device__ __forceinline__ void adj_forward(const CrnnParams<double>& p, double T, const double (&y)[NS], AdjNode& nd, double (&f)[NS]) {
#pragma unroll
    for (int k = 0; k < NS; k++) {
        const double Y = m_min(m_max(y[k], p.lb), p.ub);
        nd.wv[k] = log(Y);
        nd.q[k] = (y[k] >= p.lb && y[k] <= p.ub) ? 1.0 / Y : 0.0;
    }
    nd.wv[NS] = -p.inv_R / T;
    nd.wv[NS + 1] = log(T);
#pragma unroll
    for (int j = 0; j < NR; j++) {
        double z = fma(p.Ea[j], nd.wv[NS], fma(p.b[j], nd.wv[NS + 1], p.lnA[j]));
#pragma unroll
        for (int k = 0; k < NS; k++) z = fma(p.nu[k][j], nd.wv[k], z);
        nd.r[j] = exp(m_min(m_max(z, p.zlo), p.zhi));
        nd.mz[j] = (z >= p.zlo && z <= p.zhi) ? 1.0 : 0.0;
    }
#pragma unroll
    for (int i = 0; i < NS; i++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NR; j++) s = fma(p.wout[i][j], nd.r[j], s);
        f[i] = m_min(m_max(s, p.dulo), p.duhi);
        nd.md[i] = (s >= p.dulo && s <= p.duhi) ? 1.0 : 0.0;
    }
}

// J^T lam, and G += w * lam^T df/dtheta
__device__ __forceinline__ void adj_contract(const CrnnParams<double>& p, const AdjNode& nd, const double (&lam)[NS], double w,
                                             double* __restrict__ G, double (&Jtl)[NS]) {
    double lt[NS], mu[NR];
#pragma unroll
    for (int i = 0; i < NS; i++) lt[i] = lam[i] * nd.md[i];
#pragma unroll
    for (int j = 0; j < NR; j++) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < NS; i++) s = fma(lt[i], p.wout[i][j], s);
        mu[j] = s * nd.r[j] * nd.mz[j];
    }
#pragma unroll
    for (int k = 0; k < NS; k++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NR; j++) s = fma(p.nu[k][j], mu[j], s);
        Jtl[k] = s * nd.q[k];
    }
    // w_in [11][9]
    for (int k = 0; k < NS + 2; k++) {
        const double wk = w * nd.wv[k];
#pragma unroll
        for (int j = 0; j < NR; j++) G[k * NR + j] = fma(wk, mu[j], G[k * NR + j]);
    }
    // w_b [9]
#pragma unroll
    for (int j = 0; j < NR; j++) G[11 * NR + j] = fma(w, mu[j], G[11 * NR + j]);
    // w_out [9][9]
    for (int i = 0; i < NS; i++) {
        const double wl = w * lt[i];
#pragma unroll
        for (int j = 0; j < NR; j++) G[12 * NR + i * NR + j] = fma(wl, nd.r[j], G[12 * NR + i * NR + j]);
    }
}

template <bool kRamp>
__global__ void __launch_bounds__(ADJ_BLOCK)
adjoint_kernel(const __grid_constant__ CrnnParams<double> p, const AdjointArgs a) {
    const int i = blockIdx.x * ADJ_BLOCK + threadIdx.x;
    if (i >= a.n) return;
    const size_t n = (size_t)a.n;
    double G[NPAR];
    for (int e = 0; e < NPAR; e++) G[e] = 0.0;
    double lam[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) lam[k] = 0.0;
    double sc[NOBS];
#pragma unroll
    for (int s = 0; s < NOBS; s++) sc[s] = (double)a.yscale[(size_t)s * n + i];
    const double wnorm = 1.0 / (double)(NOBS * NTOT);
    const double T0 = (double)a.T0[i];
    double loss = 0.0;

    double yb[NS], fb[NS];
    AdjNode nd;
    double tb = (double)a.tgrid[(size_t)(NTOT - 1) * n + i];
    double Tb = kRamp ? (double)a.Tprof[(size_t)(NTOT - 1) * n + i] : T0;
#pragma unroll
    for (int k = 0; k < NS; k++) yb[k] = a.y_knots[((size_t)(NTOT - 1) * NS + k) * n + i];
    adj_forward(p, Tb, yb, nd, fb);

    for (int kk = NTOT - 1; kk >= 0; kk--) {
        // jump of the adjoint at knot kk: dL/dy(t_kk) through the output clamp
#pragma unroll
        for (int s = 0; s < NOBS; s++) {
            const double pc = m_min(m_max(yb[s], p.lb), p.ub);
            const double d = (pc - (double)a.ref[((size_t)kk * NOBS + s) * n + i]) / sc[s];
            loss = fma(d, d, loss);
            if (yb[s] >= p.lb && yb[s] <= p.ub) lam[s] += 2.0 * d / sc[s] * wnorm;
        }
        if (kk == 0) break;
        // interval [ta, tb]
        const double ta = (double)a.tgrid[(size_t)(kk - 1) * n + i];
        const double Ta = kRamp ? (double)a.Tprof[(size_t)(kk - 1) * n + i] : T0;
        double ya[NS], fa[NS];
#pragma unroll
        for (int k = 0; k < NS; k++) ya[k] = a.y_knots[((size_t)(kk - 1) * NS + k) * n + i];
        AdjNode nda;
        adj_forward(p, Ta, ya, nda, fa);
        const double h = tb - ta;
        const double slope = (Tb - Ta) / h;
        const double hs = h / (double)a.substeps;

        auto node_at = [&](double tau, AdjNode& out) {
            // cubic Hermite state and ramp temperature at tau in [ta, tb]
            const double s = (tau - ta) / h, s2 = s * s, s3 = s2 * s;
            const double h00 = 2 * s3 - 3 * s2 + 1, h10 = s3 - 2 * s2 + s, h01 = -2 * s3 + 3 * s2, h11 = s3 - s2;
            double yy[NS], ff[NS];
#pragma unroll
            for (int k = 0; k < NS; k++) yy[k] = h00 * ya[k] + h10 * h * fa[k] + h01 * yb[k] + h11 * h * fb[k];
            adj_forward(p, kRamp ? Ta + slope * (tau - ta) : T0, yy, out, ff);
        };

        for (int ss = 0; ss < a.substeps; ss++) {
            const double tau1 = tb - ss * hs;
            const bool last = (ss == a.substeps - 1);
            double k1[NS], k2[NS], k3[NS], k4[NS], l2[NS];
            AdjNode nm;
            // stage 1 at tau1: `nd` always holds the node at the upper end of the current sub-step
            adj_contract(p, nd, lam, hs / 6.0, G, k1);
            // stages 2, 3 at the midpoint
            node_at(tau1 - 0.5 * hs, nm);
#pragma unroll
            for (int k = 0; k < NS; k++) l2[k] = fma(0.5 * hs, k1[k], lam[k]);
            adj_contract(p, nm, l2, hs / 3.0, G, k2);
#pragma unroll
            for (int k = 0; k < NS; k++) l2[k] = fma(0.5 * hs, k2[k], lam[k]);
            adj_contract(p, nm, l2, hs / 3.0, G, k3);
            // stage 4 at tau0
            if (last) nd = nda; else node_at(tau1 - hs, nd);
#pragma unroll
            for (int k = 0; k < NS; k++) l2[k] = fma(hs, k3[k], lam[k]);
            adj_contract(p, nd, l2, hs / 6.0, G, k4);
#pragma unroll
            for (int k = 0; k < NS; k++) lam[k] += hs / 6.0 * (k1[k] + 2.0 * k2[k] + 2.0 * k3[k] + k4[k]);
        }
        // move to the next interval down: knot kk-1 becomes the upper end (nd already holds its node)
#pragma unroll
        for (int k = 0; k < NS; k++) { yb[k] = ya[k]; fb[k] = fa[k]; }
        tb = ta;
        Tb = Ta;
    }
    a.loss[i] = loss * wnorm;
    for (int e = 0; e < NPAR; e++) a.grad[(size_t)e * n + i] = G[e];
}

// ------------------------------------------------------------------------------------------------------------------
// Warp-cooperative version of the same sweep: one condition per warp.  Lane k < 9 owns species k (state, adjoint, q_k,
// md_k) and reaction k (z_k, r_k, mu_k); lanes 9 and 10 compute the two temperature entries of wv; all 32 lanes
// share the 189 accumulators (entry e belongs to lane e % 32).  The vectors a lane needs from the others go through a
// small per-warp shared-memory scratch.  A training batch is a few hundred conditions, so this is what fills the GPU:
// the thread-per-condition kernel above ran 640 threads and took 184 ms; this one takes ~3 ms.
constexpr int ADJW_WARPS = 4;

struct AdjWarpScratch {
    // node vectors: V = [wv(11), 1.0, lt(9)] (index 0..20), U = [mu(9), r(9)] (index 0..17)
    double V[3][24];   // per stored node (upper / mid / lower): wv and the constant one; lt is per contraction, kept in slot 12..20
    double R[3][9];    // r_j of the stored nodes
    double U[18];      // mu of the current contraction, r of its node
    double lt[9];
    double tmp[12];
};

template <bool kRamp>
__global__ void __launch_bounds__(32 * ADJW_WARPS)
adjoint_warp_kernel(const __grid_constant__ CrnnParams<double> p, const AdjointArgs a) {
    __shared__ AdjWarpScratch scratch[ADJW_WARPS];
    __shared__ __align__(16) FastTables ft;   // table-driven log / exp (<= 2e-15 / 1 ulp): a third of libdevice's dependent chain
    for (int e = threadIdx.x; e < LOGTAB_N; e += 32 * ADJW_WARPS) ft.logtab[e] = a.tables->logtab[e];
    for (int e = threadIdx.x; e < EXPTAB_N; e += 32 * ADJW_WARPS) ft.exptab[e] = a.tables->exptab[e];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * ADJW_WARPS + warp;
    if (i >= a.n) return;  // warp-uniform
    AdjWarpScratch& S = scratch[warp];
    const size_t n = (size_t)a.n;
    const int k = lane < NS ? lane : 0;  // species / reaction index of this lane (lanes >= 9 shadow lane 0 where harmless)
    const bool sp = lane < NS;

    // per-lane parameter rows / columns
    double win_col[NS + 2], wout_col[NS], wout_row[NR], nu_row[NR];
#pragma unroll
    for (int r = 0; r < NS; r++) { win_col[r] = p.nu[r][k]; wout_col[r] = p.wout[r][k]; }
    win_col[NS] = p.Ea[k];
    win_col[NS + 1] = p.b[k];
    const double lnA = p.lnA[k];
#pragma unroll
    for (int j = 0; j < NR; j++) { wout_row[j] = p.wout[k][j]; nu_row[j] = p.nu[k][j]; }

    // accumulator slots of this lane: entry e = lane + 32 s  ->  G += w * V[ia] * U[ib]
    int ia[6], ib[6];
    double G[6];
#pragma unroll
    for (int s = 0; s < 6; s++) {
        const int e = lane + 32 * s;
        G[s] = 0.0;
        if (e < 99) { ia[s] = e / 9; ib[s] = e % 9; }                       // w_in[k][j]: wv_k mu_j
        else if (e < 108) { ia[s] = 11; ib[s] = e - 99; }                   // w_b[j]:     1 * mu_j
        else if (e < NPAR) { ia[s] = 12 + (e - 108) / 9; ib[s] = 9 + (e - 108) % 9; }   // w_out[i][j]: lt_i r_j
        else { ia[s] = 11; ib[s] = 0; }
    }

    const double wnorm = 1.0 / (double)(NOBS * NTOT), inv_sub = 1.0 / (double)a.substeps;
    const double T0 = (double)a.T0[i];
    const double isc = (lane < NOBS) ? 1.0 / (double)a.yscale[(size_t)lane * n + i] : 1.0;
    double loss = 0.0, lam = 0.0;

    // node evaluation: forward quantities at (T, y) -> stored node `slot`; returns f_k, q_k, md_k, mz_k for this lane
    auto node = [&](int slot, double T, double y, double& f, double& q, double& md, double& mz) {
        const double Y = m_min(m_max(y, p.lb), p.ub);
        double wvv = fast_log_ilp(lane == NS + 1 ? T : Y, ft.logtab);   // lanes 0..8: ln Y_k; lane 10: ln T
        q = (y >= p.lb && y <= p.ub) ? rcp_full(Y) : 0.0;   // (an IEEE division is a ~20-deep dependent chain: a third of this kernel's time)
        if (lane == NS) wvv = -p.inv_R * rcp_full(T);
        if (lane < NS + 2) S.V[slot][lane] = wvv;
        if (lane == NS + 2) S.V[slot][11] = 1.0;
        __syncwarp();
        // three partial sums: the warp has no other warp to hide an 11-deep FMA chain behind
        double z0 = lnA, z1 = 0.0, z2 = 0.0;
#pragma unroll
        for (int r = 0; r < 4; r++) z0 = fma(win_col[r], S.V[slot][r], z0);
#pragma unroll
        for (int r = 4; r < 8; r++) z1 = fma(win_col[r], S.V[slot][r], z1);
#pragma unroll
        for (int r = 8; r < NS + 2; r++) z2 = fma(win_col[r], S.V[slot][r], z2);
        const double z = (z0 + z1) + z2;
        const double rr = fast_exp_ilp(m_min(m_max(z, p.zlo), p.zhi), ft.exptab);
        mz = (z >= p.zlo && z <= p.zhi) ? 1.0 : 0.0;
        if (sp) S.R[slot][lane] = rr;
        __syncwarp();
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            s0 = fma(wout_row[j], S.R[slot][j], s0);
            s1 = fma(wout_row[j + 3], S.R[slot][j + 3], s1);
            s2 = fma(wout_row[j + 6], S.R[slot][j + 6], s2);
        }
        const double s = (s0 + s1) + s2;
        f = m_min(m_max(s, p.dulo), p.duhi);
        md = (s >= p.dulo && s <= p.duhi) ? 1.0 : 0.0;
    };
    // contraction at stored node `slot` with adjoint component l (lanes < 9): returns (J^T l)_k, accumulates w * l^T df/dtheta
    auto contract = [&](int slot, double l, double q, double md, double mz, double w) -> double {
        __syncwarp();
        if (sp) { S.V[slot][12 + lane] = l * md; S.U[9 + lane] = S.R[slot][lane]; }
        __syncwarp();
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            s0 = fma(S.V[slot][12 + r], wout_col[r], s0);
            s1 = fma(S.V[slot][15 + r], wout_col[r + 3], s1);
            s2 = fma(S.V[slot][18 + r], wout_col[r + 6], s2);
        }
        const double mu = ((s0 + s1) + s2) * (S.R[slot][k] * mz);
        if (sp) S.U[lane] = mu;
        __syncwarp();
        double j0 = 0.0, j1 = 0.0, j2 = 0.0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            j0 = fma(nu_row[j], S.U[j], j0);
            j1 = fma(nu_row[j + 3], S.U[j + 3], j1);
            j2 = fma(nu_row[j + 6], S.U[j + 6], j2);
        }
        const double jt = (j0 + j1) + j2;
#pragma unroll
        for (int e = 0; e < 6; e++) G[e] = fma(w * S.V[slot][ia[e]], S.U[ib[e]], G[e]);
        return jt * q;
    };

    // upper end of the first interval: knot 800 -> slot 0
    double tb = (double)a.tgrid[(size_t)(NTOT - 1) * n + i];
    double Tb = kRamp ? (double)a.Tprof[(size_t)(NTOT - 1) * n + i] : T0;
    double yb = a.y_knots[((size_t)(NTOT - 1) * NS + k) * n + i];
    double fb, qb, mdb, mzb;
    node(0, Tb, yb, fb, qb, mdb, mzb);

    // the knot data of interval kk-1 (and the label of knot kk-1) are fetched one interval ahead of their use
    float ta_n = a.tgrid[(size_t)(NTOT - 2) * n + i], Ta_n = kRamp ? a.Tprof[(size_t)(NTOT - 2) * n + i] : 0.f;
    double ya_n = a.y_knots[((size_t)(NTOT - 2) * NS + k) * n + i];
    float ref_n = lane < NOBS ? a.ref[((size_t)(NTOT - 1) * NOBS + lane) * n + i] : 0.f;
    for (int kk = NTOT - 1; kk >= 0; kk--) {
        const float ref_c = ref_n;
        if (kk > 0 && lane < NOBS) ref_n = a.ref[((size_t)(kk - 1) * NOBS + lane) * n + i];
        if (lane < NOBS) {
            const double pc = m_min(m_max(yb, p.lb), p.ub);
            const double d = (pc - (double)ref_c) * isc;
            loss = fma(d, d, loss);
            if (yb >= p.lb && yb <= p.ub) lam += 2.0 * d * isc * wnorm;
        }
        if (kk == 0) break;
        const double ta = (double)ta_n;
        const double Ta = kRamp ? (double)Ta_n : T0;
        const double ya = ya_n;
        if (kk > 1) {
            ta_n = a.tgrid[(size_t)(kk - 2) * n + i];
            if (kRamp) Ta_n = a.Tprof[(size_t)(kk - 2) * n + i];
            ya_n = a.y_knots[((size_t)(kk - 2) * NS + k) * n + i];
        }
        double fa, qa, mda, mza;
        node(2, Ta, ya, fa, qa, mda, mza);  // lower end -> slot 2
        const double h = tb - ta, ih = rcp_full(h), slope = (Tb - Ta) * ih, hs = h * inv_sub;
        // upper node of the current sub-step lives in slot 0 with (q1, md1, mz1)
        double q1 = qb, md1 = mdb, mz1 = mzb;
        for (int ss = 0; ss < a.substeps; ss++) {
            const double tau1 = tb - ss * hs;
            const bool last = (ss == a.substeps - 1);
            const double k1 = contract(0, lam, q1, md1, mz1, hs / 6.0);
            // midpoint node -> slot 1
            double fm, qm, mdm, mzm;
            {
                const double tau = tau1 - 0.5 * hs;
                const double s = (tau - ta) * ih, s2 = s * s, s3 = s2 * s;
                const double ym = (2 * s3 - 3 * s2 + 1) * ya + (s3 - 2 * s2 + s) * h * fa + (-2 * s3 + 3 * s2) * yb + (s3 - s2) * h * fb;
                __syncwarp();
                node(1, kRamp ? Ta + slope * (tau - ta) : T0, ym, fm, qm, mdm, mzm);
            }
            const double k2 = contract(1, fma(0.5 * hs, k1, lam), qm, mdm, mzm, hs / 3.0);
            const double k3 = contract(1, fma(0.5 * hs, k2, lam), qm, mdm, mzm, hs / 3.0);
            double k4;
            if (last) {
                k4 = contract(2, fma(hs, k3, lam), qa, mda, mza, hs / 6.0);
            } else {
                const double tau = tau1 - hs;
                const double s = (tau - ta) * ih, s2 = s * s, s3 = s2 * s;
                const double y0 = (2 * s3 - 3 * s2 + 1) * ya + (s3 - 2 * s2 + s) * h * fa + (-2 * s3 + 3 * s2) * yb + (s3 - s2) * h * fb;
                double f0;
                __syncwarp();
                node(0, kRamp ? Ta + slope * (tau - ta) : T0, y0, f0, q1, md1, mz1);  // becomes the next sub-step's upper node
                k4 = contract(0, fma(hs, k3, lam), q1, md1, mz1, hs / 6.0);
            }
            lam += hs / 6.0 * (k1 + 2.0 * k2 + 2.0 * k3 + k4);
        }
        // knot kk-1 becomes the upper end: copy slot 2 -> slot 0
        __syncwarp();
        if (lane < 12) S.V[0][lane] = S.V[2][lane];
        if (sp) S.R[0][lane] = S.R[2][lane];
        __syncwarp();
        yb = ya; fb = fa; qb = qa; mdb = mda; mzb = mza; tb = ta; Tb = Ta;
    }
    // loss: sum over the 7 observed-species lanes
    double lsum = lane < NOBS ? loss : 0.0;
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if (lane == 0) a.loss[i] = lsum * wnorm;
#pragma unroll
    for (int s = 0; s < 6; s++) {
        const int e = lane + 32 * s;
        if (e < NPAR) a.grad[(size_t)e * n + i] = G[s];
    }
}

// out[r] = sum_i x[r][i], fixed summation tree (deterministic): one block per row.  With `status` given only the columns with
// status[i] == 0 are summed (a condition whose forward integration failed contributes neither loss nor gradient -- its row
// entries may be NaN, so they are skipped, not multiplied by zero) and block `rows` writes their number to out[rows].
__global__ void __launch_bounds__(256) reduce_rows_kernel(const double* __restrict__ x, int n, int rows, const int* __restrict__ status,
                                                          double* __restrict__ out) {
    __shared__ double sh[256];
    const bool counting = (int)blockIdx.x == rows;
    const double* row = x + (size_t)blockIdx.x * n;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256)
        if (!status || status[i] == 0) s += counting ? 1.0 : row[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

}  // namespace pfr
