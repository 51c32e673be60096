// The training step's gradient in three kernels (pfr_loss_grad_staged): the same continuous adjoint, the same RK4 discretisation
// and the same Hermite interpolant as adjoint_warp_kernel (adjoint.cuh), split by WHAT DEPENDS ON WHAT.
//
// adjoint_warp_kernel walks one condition per warp backwards through 800 knot intervals and does everything inside that walk.
// With a few hundred conditions on 148 SMs nothing hides its latencies: 6.7 us per interval, 5.2 ms for 640 conditions, of
// which only a ninth is inherently sequential.  The adjoint equation lam' = -J^T lam is LINEAR in lam, so:
//
//   phase 1  adjoint_nodes_kernel   everything that depends on the forward trajectory only -- the forward quantities at every
//            quadrature node (interpolated state, rates, clamp masks) and the 9 x 9 matrix M = J^T of the node -- for all
//            conditions x intervals x nodes at once: one thread per (interval, condition), a throughput kernel (2.0 M nodes
//            for 640 conditions).  Node records go to a workspace laid out [node][field][condition] (coalesced stores).
//   phase 2  adjoint_sweep_kernel   the only sequential part: one warp per condition walks the intervals backwards, each RK4
//            stage is ONE 9 x 9 mat-vec (row k of M in lane k, lam by shuffle, three partial sums) instead of two mat-vecs, an
//            exponential / logarithm evaluation and 189 accumulator updates; the next interval's matrices are fetched while
//            the current one is processed.  It also adds the loss jumps at the knots and records the stage vectors.
//   phase 3  adjoint_grad_kernel    the parameter gradient  sum_nodes w (lam^T df/dtheta): needs lam at every stage, but no
//            stage needs another one -- one block per condition, eight warps sharing its intervals, fixed-order reduction.
//
// Numerically this is adjoint_warp_kernel with the sums of a stage re-associated (M is formed first); the two agree to ~1e-12 of
// the gradient's scale (tests/test_training.py).
#pragma once
#include "adjoint.cuh"

namespace pfr {

// Node record, 128 doubles per condition:
//   M part   [row k = 0..8][condition][10]   row k of M = J^T padded to 10 doubles: a lane of the adjoint walk fetches its row as five
//                                            16-byte copies, and a warp of the node kernel stores 32 adjacent rows contiguously
//   vectors  [field = 0..37][condition]      wv[11] | g[9] = r [z unclamped] | r[9] | md[9]
constexpr int ADJ_MROW = 10, ADJ_NV = 38, ADJ_NF = NS * ADJ_MROW + ADJ_NV;
constexpr int ADJ_F_WV = 0, ADJ_F_G = 11, ADJ_F_R = 20, ADJ_F_MD = 29;
constexpr int ADJP_BLOCK = 128;

// nodes per condition: the 801 knots, then for every interval kk = 1..800 its 2 S - 1 interior nodes at tb - (j + 1) hs / 2
__host__ __device__ inline size_t adj_nodes_per_condition(int S) { return (size_t)NTOT + (size_t)(NTOT - 1) * (2 * S - 1); }
__host__ __device__ inline size_t adj_interior_node(int kk, int j, int S) { return (size_t)NTOT + (size_t)(kk - 1) * (2 * S - 1) + j; }
__host__ __device__ inline size_t adj_stages_per_condition(int S) { return (size_t)(NTOT - 1) * S * 4; }

struct AdjPhaseArgs {
    AdjointArgs a;
    double* nodes;    // [nodes_per_condition] records of ADJ_NF * n doubles: M part, then vectors (layout above)
    double* stages;   // [n][stages_per_condition][9]: lam at every RK4 stage, stage index ((kk - 1) S + ss) 4 + st
};

__host__ __device__ inline size_t adj_m_offset(size_t node, int k, size_t i, size_t n) { return (node * ADJ_NF + (size_t)k * ADJ_MROW) * n + i * ADJ_MROW; }
__host__ __device__ inline size_t adj_v_offset(size_t node, int field, size_t i, size_t n) { return (node * ADJ_NF + (size_t)NS * ADJ_MROW + field) * n + i; }

// ---------------------------------------------------------------------------------------------------------------- phase 1
// forward quantities + M = J^T at (T, y); the record is stored unless rec == nullptr; returns f
__device__ __forceinline__ void adj_node_full(const CrnnParams<double>& p, const FastTables& ft, double T, const double (&y)[NS],
                                              double* __restrict__ nodes, long long node, size_t i, size_t n, double (&f)[NS]) {
    double wv[NS + 2], q[NS], g[NR], r[NR], md[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) {
        const double Y = m_min(m_max(y[k], p.lb), p.ub);
        wv[k] = fast_log(Y, ft.logtab);
        q[k] = (y[k] >= p.lb && y[k] <= p.ub) ? rcp_full(Y) : 0.0;
    }
    wv[NS] = -p.inv_R * rcp_full(T);
    wv[NS + 1] = fast_log(T, ft.logtab);
#pragma unroll
    for (int j = 0; j < NR; j++) {
        double z = fma(p.Ea[j], wv[NS], fma(p.b[j], wv[NS + 1], p.lnA[j]));
#pragma unroll
        for (int k = 0; k < NS; k++) z = fma(p.nu[k][j], wv[k], z);
        r[j] = fast_exp(m_min(m_max(z, p.zlo), p.zhi), ft.exptab);
        g[j] = (z >= p.zlo && z <= p.zhi) ? r[j] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < NS; i++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NR; j++) s = fma(p.wout[i][j], r[j], s);
        f[i] = m_min(m_max(s, p.dulo), p.duhi);
        md[i] = (s >= p.dulo && s <= p.duhi) ? 1.0 : 0.0;
    }
    if (node < 0) return;
    // (the vectors that only phase 3 needs leave first, so that their registers are free while M is assembled)
    double* __restrict__ rv = nodes + adj_v_offset((size_t)node, 0, i, n);
#pragma unroll
    for (int e = 0; e < NS + 2; e++) rv[(size_t)(ADJ_F_WV + e) * n] = wv[e];
#pragma unroll
    for (int j = 0; j < NR; j++) {
        rv[(size_t)(ADJ_F_G + j) * n] = g[j];
        rv[(size_t)(ADJ_F_R + j) * n] = r[j];
        rv[(size_t)(ADJ_F_MD + j) * n] = md[j];
    }
    // M[k][i] = q_k md_i sum_j nu[k][j] g_j wout[i][j]   ((J^T lam)_k = sum_i M[k][i] lam_i)
#pragma unroll
    for (int k = 0; k < NS; k++) {
        double aj[NR];
#pragma unroll
        for (int j = 0; j < NR; j++) aj[j] = p.nu[k][j] * g[j];
        double row[ADJ_MROW];
#pragma unroll
        for (int c = 0; c < NS; c++) {
            double m = 0.0;
#pragma unroll
            for (int j = 0; j < NR; j++) m = fma(aj[j], p.wout[c][j], m);
            row[c] = q[k] * md[c] * m;
        }
        row[NS] = 0.0;
        double2* __restrict__ dst = reinterpret_cast<double2*>(nodes + adj_m_offset((size_t)node, k, i, n));   // 80-byte rows: 16-byte aligned
#pragma unroll
        for (int c = 0; c < ADJ_MROW / 2; c++) dst[c] = make_double2(row[2 * c], row[2 * c + 1]);
    }
}

// one thread per (interval kk = blockIdx.y + 1, condition): the record of knot kk (and of knot 0 from the kk = 1 thread) and of the
// interval's 2 S - 1 interior nodes
template <bool kRamp>
__global__ void __launch_bounds__(ADJP_BLOCK)
adjoint_nodes_kernel(const __grid_constant__ CrnnParams<double> p, const AdjPhaseArgs g) {
    __shared__ __align__(16) FastTables ft;
    for (int e = threadIdx.x; e < LOGTAB_N; e += ADJP_BLOCK) ft.logtab[e] = g.a.tables->logtab[e];
    for (int e = threadIdx.x; e < EXPTAB_N; e += ADJP_BLOCK) ft.exptab[e] = g.a.tables->exptab[e];
    __syncthreads();
    const AdjointArgs& a = g.a;
    const int i = blockIdx.x * ADJP_BLOCK + threadIdx.x, kk = blockIdx.y + 1, S = a.substeps;
    if (i >= a.n) return;
    const size_t n = (size_t)a.n;
    const double T0 = (double)a.T0[i];
    const double ta = (double)a.tgrid[(size_t)(kk - 1) * n + i], tb = (double)a.tgrid[(size_t)kk * n + i];
    const double Ta = kRamp ? (double)a.Tprof[(size_t)(kk - 1) * n + i] : T0, Tb = kRamp ? (double)a.Tprof[(size_t)kk * n + i] : T0;
    // knot states and slopes of the interval live in shared memory ([vector][species][thread], conflict-free): 72 registers
    __shared__ double hv[4][NS][ADJP_BLOCK];
    double* const ya = &hv[0][0][threadIdx.x];
    double* const yb = &hv[1][0][threadIdx.x];
    double* const fa = &hv[2][0][threadIdx.x];
    double* const fb = &hv[3][0][threadIdx.x];
#define HV(v, k) (v)[(k) * ADJP_BLOCK]
    double ym[NS], fm[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) {
        HV(ya, k) = a.y_knots[((size_t)(kk - 1) * NS + k) * n + i];
        HV(yb, k) = a.y_knots[((size_t)kk * NS + k) * n + i];
    }
    const double h = tb - ta, ih = rcp_full(h), slope = (Tb - Ta) * ih, half = 0.5 * h / (double)S;
    // ONE instance of the node evaluation in the instruction stream (three inlined copies spilled 700 bytes per thread):
    // j = -2 the upper knot, j = -1 the lower knot (its record belongs to the interval below, except for knot 0), j >= 0 interior
#pragma unroll 1
    for (int j = -2; j < 2 * S - 1; j++) {
        double Tn;
        long long node;   // < 0: no record
        if (j == -2) {
            Tn = Tb;
            node = kk;
#pragma unroll
            for (int k = 0; k < NS; k++) ym[k] = HV(yb, k);
        } else if (j == -1) {
            Tn = Ta;
            node = kk == 1 ? 0 : -1;
#pragma unroll
            for (int k = 0; k < NS; k++) ym[k] = HV(ya, k);
        } else {
            const double tau = tb - (double)(j + 1) * half;
            const double s = (tau - ta) * ih, s2 = s * s, s3 = s2 * s;
            const double ca = 2 * s3 - 3 * s2 + 1, cfa = (s3 - 2 * s2 + s) * h, cb = -2 * s3 + 3 * s2, cfb = (s3 - s2) * h;
#pragma unroll
            for (int k = 0; k < NS; k++) ym[k] = ca * HV(ya, k) + cfa * HV(fa, k) + cb * HV(yb, k) + cfb * HV(fb, k);
            Tn = kRamp ? Ta + slope * (tau - ta) : T0;
            node = (long long)adj_interior_node(kk, j, S);
        }
        adj_node_full(p, ft, Tn, ym, g.nodes, node, (size_t)i, n, fm);
        if (j == -2) {
#pragma unroll
            for (int k = 0; k < NS; k++) HV(fb, k) = fm[k];
        } else if (j == -1) {
#pragma unroll
            for (int k = 0; k < NS; k++) HV(fa, k) = fm[k];
        }
    }
#undef HV
}

// ---------------------------------------------------------------------------------------------------------------- phase 2
constexpr int ADJS_WARPS = 2;    // conditions per block: few, so that the warps spread over all SMs
constexpr int ADJS_DEPTH = 8;    // sub-steps whose matrices are in flight: the walk is a chain of dependent mat-vecs (~0.4 us per
                                 // sub-step) with nothing else to hide a DRAM round trip (~1 us) behind, so the rows it needs
                                 // are copied into a shared-memory ring eight sub-steps ahead (cp.async, 8 bytes per copy; each
                                 // lane copies and later reads its own row, so no synchronisation beyond wait_group is needed)

__device__ __forceinline__ void adj_cp16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void adj_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending> __device__ __forceinline__ void adj_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

__global__ void __launch_bounds__(32 * ADJS_WARPS)
adjoint_sweep_kernel(const __grid_constant__ CrnnParams<double> p, const AdjPhaseArgs g) {
    __shared__ __align__(16) double ring[ADJS_WARPS][ADJS_DEPTH][2][NS][ADJ_MROW];   // [slot][mid | low][row k][column, padded]
    const AdjointArgs& a = g.a;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * ADJS_WARPS + warp;
    if (i >= a.n) return;   // warp-uniform
    const size_t n = (size_t)a.n;
    const int S = a.substeps;
    const bool sp = lane < NS;
    const int k = sp ? lane : 0;
    const double wnorm = 1.0 / (double)(NOBS * NTOT), inv_sub = 1.0 / (double)S;
    const double isc = (lane < NOBS) ? 1.0 / (double)a.yscale[(size_t)lane * n + i] : 1.0;
    double* __restrict__ st_out = g.stages + (size_t)i * adj_stages_per_condition(S) * NS;
    const int total = (NTOT - 1) * S;
    // row k of M at node `node` -> registers (first node only) / shared-memory ring (everything else)
    auto row_src = [&](size_t node) { return g.nodes + adj_m_offset(node, k, (size_t)i, n); };
    auto issue = [&](int q) {   // sub-step q = (800 - kk) S + ss: its mid-point node and its lower node
        if (q < total && sp) {
            const int kk = NTOT - 1 - q / S, ss = q % S;
            const double* m = row_src(adj_interior_node(kk, 2 * ss, S));
            const double* l = row_src(ss == S - 1 ? (size_t)(kk - 1) : adj_interior_node(kk, 2 * ss + 1, S));
            double* dm = &ring[warp][q % ADJS_DEPTH][0][k][0];
            double* dl = &ring[warp][q % ADJS_DEPTH][1][k][0];
#pragma unroll
            for (int c = 0; c < ADJ_MROW; c += 2) {
                adj_cp16(dm + c, m + c);
                adj_cp16(dl + c, l + c);
            }
        }
        adj_cp_commit();   // (an empty group past the end keeps the group count in step with q)
    };
    // (M l)_k with l spread over the lanes
    auto matvec = [&](const double (&row)[NS], double l) -> double {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            s0 = fma(row[c], __shfl_sync(0xffffffffu, l, c), s0);
            s1 = fma(row[c + 3], __shfl_sync(0xffffffffu, l, c + 3), s1);
            s2 = fma(row[c + 6], __shfl_sync(0xffffffffu, l, c + 6), s2);
        }
        return (s0 + s1) + s2;
    };
    for (int q = 0; q < ADJS_DEPTH; q++) issue(q);
    double loss = 0.0, lam = 0.0;
    double up[NS], mid[NS], low[NS];
    {
        const double* r0 = row_src((size_t)(NTOT - 1));   // knot 800
#pragma unroll
        for (int c = 0; c < NS; c++) up[c] = sp ? r0[c] : 0.0;
    }
    double tb = (double)a.tgrid[(size_t)(NTOT - 1) * n + i];
    double yb = a.y_knots[((size_t)(NTOT - 1) * NS + k) * n + i];
    float ref_n = lane < NOBS ? a.ref[((size_t)(NTOT - 1) * NOBS + lane) * n + i] : 0.f;
    float ta_n = a.tgrid[(size_t)(NTOT - 2) * n + i];
    double ya_n = a.y_knots[((size_t)(NTOT - 2) * NS + k) * n + i];
    double hs = 0.0, ta = 0.0, ya = 0.0;
    auto knot_jump = [&](int kk) {             // loss term and adjoint jump at knot kk (its state is yb)
        const float ref_c = ref_n;
        if (kk > 0 && lane < NOBS) ref_n = a.ref[((size_t)(kk - 1) * NOBS + lane) * n + i];
        if (lane < NOBS) {
            const double pc = m_min(m_max(yb, p.lb), p.ub);
            const double d = (pc - (double)ref_c) * isc;
            loss = fma(d, d, loss);
            if (yb >= p.lb && yb <= p.ub) lam += 2.0 * d * isc * wnorm;
        }
    };
    for (int q = 0; q < total; q++) {
        const int kk = NTOT - 1 - q / S, ss = q % S;
        if (ss == 0) {
            knot_jump(kk);
            ta = (double)ta_n;
            ya = ya_n;
            if (kk > 1) {
                ta_n = a.tgrid[(size_t)(kk - 2) * n + i];
                ya_n = a.y_knots[((size_t)(kk - 2) * NS + k) * n + i];
            }
            hs = (tb - ta) * inv_sub;
        }
        adj_cp_wait<ADJS_DEPTH - 1>();   // this lane's copies for sub-step q have landed
#pragma unroll
        for (int c = 0; c < NS; c++) {
            mid[c] = sp ? ring[warp][q % ADJS_DEPTH][0][k][c] : 0.0;
            low[c] = sp ? ring[warp][q % ADJS_DEPTH][1][k][c] : 0.0;
        }
        issue(q + ADJS_DEPTH);             // refills the slot that was just read (same lane, program order)
        double* so = st_out + ((size_t)((kk - 1) * S + ss) * 4) * NS;
        const double l1 = lam;
        const double k1 = matvec(up, l1);
        const double l2 = fma(0.5 * hs, k1, lam);
        const double k2 = matvec(mid, l2);
        const double l3 = fma(0.5 * hs, k2, lam);
        const double k3 = matvec(mid, l3);
        const double l4 = fma(hs, k3, lam);
        const double k4 = matvec(low, l4);
        if (sp) { so[lane] = l1; so[NS + lane] = l2; so[2 * NS + lane] = l3; so[3 * NS + lane] = l4; }
        lam += hs / 6.0 * (k1 + 2.0 * k2 + 2.0 * k3 + k4);
#pragma unroll
        for (int c = 0; c < NS; c++) up[c] = low[c];
        if (ss == S - 1) { tb = ta; yb = ya; }
    }
    knot_jump(0);
    double lsum = lane < NOBS ? loss : 0.0;
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if (lane == 0) a.loss[i] = lsum * wnorm;
}

// ---------------------------------------------------------------------------------------------------------------- phase 3
// G[e] = sum over the RK4 stages s of  w_s V_s[ia(e)] U_s[ib(e)],  V = [wv(11), 1, lt(9)], U = [mu(9), r(9)]  (adjoint.cuh, header):
// a contraction over 6400 stages with 189 outputs per condition.  One block per condition; the stages go through shared memory
// in chunks: all threads load a chunk (independent loads, nothing sequential), 9 x chunk threads form mu, then thread e < 189
// accumulates its own entry over the chunk in stage order (fixed order: bit-reproducible).
constexpr int ADJG_THREADS = 192, ADJG_CHUNK = 64;

__global__ void __launch_bounds__(ADJG_THREADS, 5)   // five blocks per SM = 740 resident: all 640 conditions of a training batch in one wave
adjoint_grad_kernel(const __grid_constant__ CrnnParams<double> p, const AdjPhaseArgs g) {
    __shared__ double Vs[ADJG_CHUNK][21], Us[ADJG_CHUNK][18], Gm[ADJG_CHUNK][NS], Ws[ADJG_CHUNK], wout_s[NS][NR];
    __shared__ size_t node_s[ADJG_CHUNK];
    const AdjointArgs& a = g.a;
    const int tid = threadIdx.x, i = blockIdx.x, S = a.substeps;
    const size_t n = (size_t)a.n;
    for (int e = tid; e < NS * NR; e += ADJG_THREADS) wout_s[e / NR][e % NR] = p.wout[e / NR][e % NR];
    int ia = 11, ib = 0;
    if (tid < 99) { ia = tid / 9; ib = tid % 9; }                                     // w_in[k][j]: wv_k mu_j
    else if (tid < 108) { ia = 11; ib = tid - 99; }                                   // w_b[j]:     1 * mu_j
    else if (tid < NPAR) { ia = 12 + (tid - 108) / 9; ib = 9 + (tid - 108) % 9; }     // w_out[i][j]: lt_i r_j
    double G0 = 0.0, G1 = 0.0, G2 = 0.0, G3 = 0.0;   // four partial sums: the accumulation is a chain of dependent FMAs otherwise
    const double inv_sub = 1.0 / (double)S;
    const int total = (int)adj_stages_per_condition(S);
    const double* __restrict__ st_in = g.stages + (size_t)i * adj_stages_per_condition(S) * NS;
    for (int c0 = 0; c0 < total; c0 += ADJG_CHUNK) {
        const int cn = min(ADJG_CHUNK, total - c0);
        if (tid < cn) {   // node and quadrature weight of stage c0 + tid
            const int stage = c0 + tid, q = stage >> 2, st = stage & 3, kk = q / S + 1, ss = q % S;
            node_s[tid] = st == 0 ? (ss == 0 ? (size_t)kk : adj_interior_node(kk, 2 * ss - 1, S))
                                  : (st == 3 ? (ss == S - 1 ? (size_t)(kk - 1) : adj_interior_node(kk, 2 * ss + 1, S))
                                             : adj_interior_node(kk, 2 * ss, S));
            const double hs = ((double)a.tgrid[(size_t)kk * n + i] - (double)a.tgrid[(size_t)(kk - 1) * n + i]) * inv_sub;
            Ws[tid] = (st == 0 || st == 3) ? hs / 6.0 : hs / 3.0;
            Vs[tid][11] = 1.0;
        }
        __syncthreads();
        for (int t = tid; t < cn * NS; t += ADJG_THREADS) {
            const int s = t / NS, j = t % NS;
            const double* rec = g.nodes + adj_v_offset(node_s[s], 0, (size_t)i, n);
            Vs[s][12 + j] = st_in[(size_t)(c0 + s) * NS + j] * rec[(size_t)(ADJ_F_MD + j) * n];   // lt_j = lam_j [du_j unclamped]
            Us[s][9 + j] = rec[(size_t)(ADJ_F_R + j) * n];
            Gm[s][j] = rec[(size_t)(ADJ_F_G + j) * n];
        }
        for (int t = tid; t < cn * (NS + 2); t += ADJG_THREADS) {
            const int s = t / (NS + 2), r = t % (NS + 2);
            Vs[s][r] = g.nodes[adj_v_offset(node_s[s], ADJ_F_WV + r, (size_t)i, n)];
        }
        __syncthreads();
        for (int t = tid; t < cn * NR; t += ADJG_THREADS) {   // mu_j = g_j sum_i lt_i wout[i][j]
            const int s = t / NR, j = t % NR;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int r = 0; r < 3; r++) {
                s0 = fma(Vs[s][12 + r], wout_s[r][j], s0);
                s1 = fma(Vs[s][15 + r], wout_s[r + 3][j], s1);
                s2 = fma(Vs[s][18 + r], wout_s[r + 6][j], s2);
            }
            Us[s][j] = ((s0 + s1) + s2) * Gm[s][j];
        }
        __syncthreads();
        if (tid < NPAR) {
            int s = 0;
            for (; s + 3 < cn; s += 4) {
                G0 = fma(Ws[s] * Vs[s][ia], Us[s][ib], G0);
                G1 = fma(Ws[s + 1] * Vs[s + 1][ia], Us[s + 1][ib], G1);
                G2 = fma(Ws[s + 2] * Vs[s + 2][ia], Us[s + 2][ib], G2);
                G3 = fma(Ws[s + 3] * Vs[s + 3][ia], Us[s + 3][ib], G3);
            }
            for (; s < cn; s++) G0 = fma(Ws[s] * Vs[s][ia], Us[s][ib], G0);
        }
        __syncthreads();
    }
    if (tid < NPAR) a.grad[(size_t)tid * n + i] = (G0 + G1) + (G2 + G3);
}

}  // namespace pfr
