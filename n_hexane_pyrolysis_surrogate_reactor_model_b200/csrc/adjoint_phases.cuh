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

// Node record, 130 doubles per condition, both parts contiguous per condition so that every consumer moves 16-byte pieces:
//   M part   [row k = 0..8][condition][10]   row k of M = J^T padded to 10 doubles: a lane of the adjoint walk fetches its row as five
//                                            16-byte copies, and a warp of the node kernel stores 32 adjacent rows contiguously
//   vectors  [condition][40]                 wv[11] | g[9] = r [z unclamped] | r[9] | md[9] | 2 pad   (measured against
//                                            [field][condition]: node kernel 0.59 -> 0.80 ms, gradient kernel 1.72 -> 1.08 ms)
// Stage record [condition][stage][9]: lam at the RK4 stage; the four stages of a sub-step are 288 contiguous bytes.
constexpr int ADJ_MROW = 10, ADJ_NV = 40, ADJ_NF = NS * ADJ_MROW + ADJ_NV, ADJ_SREC = NS;
constexpr int ADJ_F_WV = 0, ADJ_F_G = 11, ADJ_F_R = 20, ADJ_F_MD = 29;
constexpr int ADJP_BLOCK = 128;

// nodes per condition: the 801 knots, then for every interval kk = 1..800 its 2 S - 1 interior nodes at tb - (j + 1) hs / 2
__host__ __device__ inline size_t adj_nodes_per_condition(int S) { return (size_t)NTOT + (size_t)(NTOT - 1) * (2 * S - 1); }
__host__ __device__ inline size_t adj_interior_node(int kk, int j, int S) { return (size_t)NTOT + (size_t)(kk - 1) * (2 * S - 1) + j; }
__host__ __device__ inline size_t adj_stages_per_condition(int S) { return (size_t)(NTOT - 1) * S * 4; }

struct AdjPhaseArgs {
    AdjointArgs a;
    double* nodes;    // [nodes_per_condition] records of ADJ_NF * n doubles: M part, then vectors (layout above)
    double* stages;   // [n][stages_per_condition][ADJ_SREC]: lam at every RK4 stage + its weight, stage index ((kk - 1) S + ss) 4 + st
};

__host__ __device__ inline size_t adj_m_offset(size_t node, int k, size_t i, size_t n) { return (node * ADJ_NF + (size_t)k * ADJ_MROW) * n + i * ADJ_MROW; }
__host__ __device__ inline size_t adj_v_offset(size_t node, size_t i, size_t n) { return (node * ADJ_NF + (size_t)NS * ADJ_MROW) * n + i * ADJ_NV; }

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
    {
        double v[ADJ_NV];
#pragma unroll
        for (int e = 0; e < NS + 2; e++) v[ADJ_F_WV + e] = wv[e];
#pragma unroll
        for (int j = 0; j < NR; j++) { v[ADJ_F_G + j] = g[j]; v[ADJ_F_R + j] = r[j]; v[ADJ_F_MD + j] = md[j]; }
        v[ADJ_NV - 2] = v[ADJ_NV - 1] = 0.0;
        double2* __restrict__ rv = reinterpret_cast<double2*>(nodes + adj_v_offset((size_t)node, i, n));   // 320-byte records
#pragma unroll
        for (int e = 0; e < ADJ_NV / 2; e++) rv[e] = make_double2(v[2 * e], v[2 * e + 1]);
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

__device__ __forceinline__ void adj_cp8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void adj_cp4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}

__global__ void __launch_bounds__(32 * ADJS_WARPS)
adjoint_sweep_kernel(const __grid_constant__ CrnnParams<double> p, const AdjPhaseArgs g) {
    __shared__ __align__(16) double ring[ADJS_WARPS][ADJS_DEPTH][2][NS][ADJ_MROW];   // [slot][mid | low][row k][column, padded]
    // Per-interval scalars of the LOWER knot of an interval (its time, this lane's state component, this lane's label) travel through
    // a second ring by cp.async as well.  (As plain loads "one interval ahead" they did not prefetch anything: the compiler moved the
    // loaded register into the variable's register right behind the load, and 37 % of this kernel's stall samples sat on those two
    // moves -- ncu source page, profiles/r02e_ncu_full_training_kernels.txt.)
    struct KnotSlot { double y; float t, ref; };
    __shared__ __align__(16) KnotSlot knots[ADJS_WARPS][ADJS_DEPTH + 1][32];
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
    double* __restrict__ st_out = g.stages + (size_t)i * adj_stages_per_condition(S) * ADJ_SREC;
    const int total = (NTOT - 1) * S;
    auto row_src = [&](size_t node) { return g.nodes + adj_m_offset(node, k, (size_t)i, n); };
    auto issue = [&](int q) {   // sub-step q = (800 - kk) S + ss: its mid-point node and its lower node; with ss = 0 the lower knot's scalars
        if (q < total) {
            const int kk = NTOT - 1 - q / S, ss = q - (q / S) * S;
            if (sp) {
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
            if (ss == 0) {
                KnotSlot* ks = &knots[warp][(q / S) % (ADJS_DEPTH + 1)][lane];
                adj_cp4(&ks->t, a.tgrid + (size_t)(kk - 1) * n + i);
                if (sp) adj_cp8(&ks->y, a.y_knots + ((size_t)(kk - 1) * NS + k) * n + i);
                if (lane < NOBS) adj_cp4(&ks->ref, a.ref + ((size_t)(kk - 1) * NOBS + lane) * n + i);
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
    double yb = sp ? a.y_knots[((size_t)(NTOT - 1) * NS + k) * n + i] : 0.0;
    float ref_b = lane < NOBS ? a.ref[((size_t)(NTOT - 1) * NOBS + lane) * n + i] : 0.f;   // label of the knot whose state is yb
    double hs = 0.0, ta = 0.0, ya = 0.0;
    float ref_a = 0.f;
    auto knot_jump = [&]() {             // loss term and adjoint jump at the knot whose state / label are yb / ref_b
        if (lane < NOBS) {
            const double pc = m_min(m_max(yb, p.lb), p.ub);
            const double d = (pc - (double)ref_b) * isc;
            loss = fma(d, d, loss);
            if (yb >= p.lb && yb <= p.ub) lam += 2.0 * d * isc * wnorm;
        }
    };
    for (int q = 0; q < total; q++) {
        const int iv = q / S, kk = NTOT - 1 - iv, ss = q - iv * S;
        adj_cp_wait<ADJS_DEPTH - 1>();   // this lane's copies for sub-step q have landed
        if (ss == 0) {
            knot_jump();
            const KnotSlot& ks = knots[warp][iv % (ADJS_DEPTH + 1)][lane];
            ta = (double)ks.t;
            ya = sp ? ks.y : 0.0;
            ref_a = lane < NOBS ? ks.ref : 0.f;
            hs = (tb - ta) * inv_sub;
        }
#pragma unroll
        for (int c = 0; c < NS; c++) {
            mid[c] = sp ? ring[warp][q % ADJS_DEPTH][0][k][c] : 0.0;
            low[c] = sp ? ring[warp][q % ADJS_DEPTH][1][k][c] : 0.0;
        }
        issue(q + ADJS_DEPTH);             // refills the slots that were just read (same lane, program order)
        double* so = st_out + ((size_t)((kk - 1) * S + ss) * 4) * ADJ_SREC;
        const double l1 = lam;
        const double k1 = matvec(up, l1);
        const double l2 = fma(0.5 * hs, k1, lam);
        const double k2 = matvec(mid, l2);
        const double l3 = fma(0.5 * hs, k2, lam);
        const double k3 = matvec(mid, l3);
        const double l4 = fma(hs, k3, lam);
        const double k4 = matvec(low, l4);
        if (sp) { so[lane] = l1; so[ADJ_SREC + lane] = l2; so[2 * ADJ_SREC + lane] = l3; so[3 * ADJ_SREC + lane] = l4; }
        lam += hs / 6.0 * (k1 + 2.0 * k2 + 2.0 * k3 + k4);
#pragma unroll
        for (int c = 0; c < NS; c++) up[c] = low[c];
        if (ss == S - 1) { tb = ta; yb = ya; ref_b = ref_a; }
    }
    knot_jump();   // knot 0
    double lsum = lane < NOBS ? loss : 0.0;
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if (lane == 0) a.loss[i] = lsum * wnorm;
}

// ---------------------------------------------------------------------------------------------------------------- phase 3
// G[e] = sum over the RK4 stages s of  w_s V_s[ia(e)] U_s[ib(e)],  V = [wv(11), 1, lt(9)], U = [mu(9), r(9)]  (adjoint.cuh, header):
// a contraction over 6400 stages with 189 outputs per condition.  One block per condition; the stages go through shared memory
// in chunks, double-buffered: while chunk c is contracted, the stage and node-vector records of chunk c + 1 are in flight
// (cp.async, 16 bytes per copy: both record types are contiguous per condition).  Per chunk: 9 x chunk threads form lt and mu, then
// thread e < 189 accumulates its own entry in stage order in four interleaved partial sums (fixed order: bit-reproducible).
constexpr int ADJG_THREADS = 192, ADJG_CHUNK = 32;

template <int kS>   // sub-steps per interval as a compile-time constant (0: read a.substeps): the stage -> node arithmetic divides by it
__global__ void __launch_bounds__(ADJG_THREADS, 5)   // five blocks per SM = 740 resident: all 640 conditions of a training batch in one wave
adjoint_grad_kernel(const __grid_constant__ CrnnParams<double> p, const AdjPhaseArgs g) {
    __shared__ __align__(16) double Srec[2][ADJG_CHUNK][ADJ_SREC];   // lam(9) of every stage of the chunk: one contiguous range of the stage records
    __shared__ __align__(16) double Vrec[2][ADJG_CHUNK][ADJ_NV];     // wv(11), g(9), r(9), md(9), pad
    __shared__ float Tk[2][ADJG_CHUNK][2];                           // knot times of the stage's interval (quadrature weight)
    __shared__ double Lt[ADJG_CHUNK][NS], Mu[ADJG_CHUNK][NR], Wt[ADJG_CHUNK], wout_s[NS][NR];
    const AdjointArgs& a = g.a;
    const int tid = threadIdx.x, i = blockIdx.x, S = kS ? kS : a.substeps;
    const size_t n = (size_t)a.n;
    for (int e = tid; e < NS * NR; e += ADJG_THREADS) wout_s[e / NR][e % NR] = p.wout[e / NR][e % NR];
    // entry e = tid:  w_in[k][j]: wv_k mu_j | w_b[j]: mu_j | w_out[i][j]: lt_i r_j   ->  (source of the first factor, index, source of the second, index)
    int ka = 0, ia = 0, ib = 0;   // ka: 0 = wv (Vrec), 1 = one, 2 = lt
    if (tid < 99) { ka = 0; ia = ADJ_F_WV + tid / 9; ib = tid % 9; }
    else if (tid < 108) { ka = 1; ib = tid - 99; }
    else if (tid < NPAR) { ka = 2; ia = (tid - 108) / 9; ib = (tid - 108) % 9; }
    double G0 = 0.0, G1 = 0.0, G2 = 0.0, G3 = 0.0;
    const double inv_sub = 1.0 / (double)S;
    const int total = (int)adj_stages_per_condition(S);
    const double* __restrict__ st_in = g.stages + (size_t)i * adj_stages_per_condition(S) * ADJ_SREC;
    auto issue = [&](int c0, int buf) {   // copies of the chunk that starts at stage c0 (nothing past the end)
        const int cn = min(ADJG_CHUNK, total - c0);
        if (cn > 0) {
            // chunks start at a multiple of 32 stages = 8 sub-steps = 2304 bytes, and a condition's stage records at a multiple of
            // 4 x 9 x 8 bytes: the chunk is one 16-byte-aligned contiguous range
            for (int t = tid; t < cn * ADJ_SREC / 2; t += ADJG_THREADS) adj_cp16(&Srec[buf][0][0] + 2 * t, st_in + (size_t)c0 * ADJ_SREC + 2 * t);
            if (tid < 2 * cn) {
                const int s = tid >> 1, q = (c0 + s) >> 2, kk = q / S + 1;
                adj_cp4(&Tk[buf][s][tid & 1], a.tgrid + (size_t)(kk - (tid & 1)) * n + i);   // [0] = t(kk), [1] = t(kk - 1)
            }
            for (int t = tid; t < cn * (ADJ_NV / 2); t += ADJG_THREADS) {
                const int s = t / (ADJ_NV / 2), c = t % (ADJ_NV / 2);
                const int stage = c0 + s, q = stage >> 2, st = stage & 3, kk = q / S + 1, ss = q % S;
                const size_t node = st == 0 ? (ss == 0 ? (size_t)kk : adj_interior_node(kk, 2 * ss - 1, S))
                                            : (st == 3 ? (ss == S - 1 ? (size_t)(kk - 1) : adj_interior_node(kk, 2 * ss + 1, S))
                                                       : adj_interior_node(kk, 2 * ss, S));
                adj_cp16(&Vrec[buf][s][2 * c], g.nodes + adj_v_offset(node, (size_t)i, n) + 2 * c);
            }
        }
        adj_cp_commit();
    };
    issue(0, 0);
    int buf = 0;
    for (int c0 = 0; c0 < total; c0 += ADJG_CHUNK, buf ^= 1) {
        const int cn = min(ADJG_CHUNK, total - c0);
        issue(c0 + ADJG_CHUNK, buf ^ 1);
        adj_cp_wait<1>();      // this thread's copies of chunk c0 have landed ...
        __syncthreads();       // ... and so have everybody else's
        for (int t = tid; t < cn * NS; t += ADJG_THREADS) {
            const int s = t / NS, j = t % NS;
            Lt[s][j] = Srec[buf][s][j] * Vrec[buf][s][ADJ_F_MD + j];   // lt_j = lam_j [du_j unclamped]
        }
        if (tid < cn) {   // RK4 quadrature weight of the stage: hs / 6 for the first and last stage of a sub-step, hs / 3 for the middle ones
            const double hs = ((double)Tk[buf][tid][0] - (double)Tk[buf][tid][1]) * inv_sub;
            const int st = (c0 + tid) & 3;
            Wt[tid] = (st == 0 || st == 3) ? hs / 6.0 : hs / 3.0;
        }
        __syncthreads();
        for (int t = tid; t < cn * NR; t += ADJG_THREADS) {            // mu_j = g_j sum_i lt_i wout[i][j]
            const int s = t / NR, j = t % NR;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int r = 0; r < 3; r++) {
                s0 = fma(Lt[s][r], wout_s[r][j], s0);
                s1 = fma(Lt[s][r + 3], wout_s[r + 3][j], s1);
                s2 = fma(Lt[s][r + 6], wout_s[r + 6][j], s2);
            }
            Mu[s][j] = ((s0 + s1) + s2) * Vrec[buf][s][ADJ_F_G + j];
        }
        __syncthreads();
        if (tid < NPAR) {
            auto term = [&](int s) {
                const double va = ka == 0 ? Vrec[buf][s][ia] : (ka == 1 ? 1.0 : Lt[s][ia]);
                const double ub = ka == 2 ? Vrec[buf][s][ADJ_F_R + ib] : Mu[s][ib];
                return Wt[s] * va * ub;
            };
            int s = 0;
            for (; s + 3 < cn; s += 4) { G0 += term(s); G1 += term(s + 1); G2 += term(s + 2); G3 += term(s + 3); }
            for (; s < cn; s++) G0 += term(s);
        }
        __syncthreads();       // the buffers of this chunk are free for the copies issued in the next iteration
    }
    if (tid < NPAR) a.grad[(size_t)tid * n + i] = (G0 + G1) + (G2 + G3);
}

}  // namespace pfr
