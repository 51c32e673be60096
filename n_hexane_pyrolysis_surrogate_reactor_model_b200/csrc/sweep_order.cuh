// Visiting order of a sweep (pfr_sweep_run): a counting sort of the conditions by a cost proxy that is known BEFORE any MLP has
// run, and the small kernels around it.
//
// Why.  The explicit integrators hand conditions to lanes from a work queue, so no exact cost sort is needed for load
// balance (measured, 2^20 LHS conditions: exact sort by outlet knot 78.5 ms, no sort 80.3 ms, proxy order 77.9 ms); what an
// order buys is a short tail (expensive conditions first) and -- the point of doing it up front -- that the MLPs WRITE the
// [801][n] grids in the order in which the integrator visits them.  With round 1's argsort of the outlet knot, computed after
// the MLPs, work item j read grid column perm[j]: every 4-byte knot value was its own 32-byte sector (74.8 KB of DRAM reads
// per condition against 3.4 KB used).  Now work item j reads column j; the eight conditions that share a sector are drawn
// from the queue within microseconds of each other and the sector is served from L2.
//
// Proxy.  Coupled (Eon) path: the outlet knot idx_cut grows with the residence time, i.e. with L / u0 at fixed (T, P).
// Isothermal (Eoff) path: the step count grows with the inlet temperature T0.  Descending key = expensive first.
// The sort is a histogram over ORDER_BINS quantised keys, an exclusive scan and a scatter that carries the four input columns
// along; positions inside one bin are handed out with atomics, i.e. in no particular order -- each condition's result does not
// depend on where it is visited, so this does not affect reproducibility of the outputs.
#pragma once
#include <cuda_runtime.h>

namespace pfr {

constexpr int ORDER_BINS = 2048;

struct OrderKey {
    const float* a;   // key = a[i] / b[i], or a[i] when b == nullptr
    const float* b;
    float lo, scale;  // bin = (key - lo) * scale, clamped to [0, ORDER_BINS), then reversed (descending order)
};
// The histogram and the scatter kernel must agree on the bin of every condition to the bit, or one position is handed out twice
// and another never (seen: one condition in 5000 when the compiler contracted `a * (1 / b) - lo` into an FMA in one of the two
// kernels only): every operation is an explicitly rounded intrinsic, which the compiler may neither fuse nor reorder.
__device__ __forceinline__ int order_bin(const OrderKey& k, int i) {
    const float v = k.b ? __fdiv_rn(k.a[i], k.b[i]) : k.a[i];
    int bin = (int)__fmul_rn(__fsub_rn(v, k.lo), k.scale);
    bin = bin < 0 ? 0 : (bin >= ORDER_BINS ? ORDER_BINS - 1 : bin);
    return ORDER_BINS - 1 - bin;
}

__global__ void __launch_bounds__(256) order_hist_kernel(const OrderKey k, int n, int* __restrict__ hist) {
    __shared__ int h[ORDER_BINS];
    for (int e = threadIdx.x; e < ORDER_BINS; e += blockDim.x) h[e] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&h[order_bin(k, i)], 1);
    __syncthreads();
    for (int e = threadIdx.x; e < ORDER_BINS; e += blockDim.x)
        if (h[e]) atomicAdd(&hist[e], h[e]);
}

// cursor[b] = number of conditions in bins < b (one block of ORDER_BINS / 2 threads, two bins each)
__global__ void __launch_bounds__(ORDER_BINS / 2) order_scan_kernel(const int* __restrict__ hist, int* __restrict__ cursor) {
    __shared__ int s[ORDER_BINS / 2];
    const int t = threadIdx.x;
    const int a = hist[2 * t], b = hist[2 * t + 1];
    s[t] = a + b;
    __syncthreads();
    for (int off = 1; off < ORDER_BINS / 2; off <<= 1) {   // Hillis-Steele inclusive scan of the pair sums
        const int v = t >= off ? s[t - off] : 0;
        __syncthreads();
        s[t] += v;
        __syncthreads();
    }
    const int before = s[t] - (a + b);
    cursor[2 * t] = before;
    cursor[2 * t + 1] = before + a;
}

// order[pos] = i and the input columns gathered into visiting order (L / U may be nullptr)
__global__ void __launch_bounds__(256)
order_scatter_kernel(const OrderKey k, int n, int* __restrict__ cursor, int* __restrict__ order, const float* __restrict__ T,
                     const float* __restrict__ P, const float* __restrict__ L, const float* __restrict__ U, float* __restrict__ Ts,
                     float* __restrict__ Ps, float* __restrict__ Ls, float* __restrict__ Us) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pos = atomicAdd(&cursor[order_bin(k, i)], 1);
    order[pos] = i;
    Ts[pos] = T[i];
    Ps[pos] = P[i];
    if (L) Ls[pos] = L[i];
    if (U) Us[pos] = U[i];
}

// slots (visiting order) whose result carries `flag` in status (caller's order): the list the Rosenbrock fallback works through
__global__ void __launch_bounds__(256)
collect_status_kernel(const int* __restrict__ status, const int* __restrict__ order, int n, int flag, int* __restrict__ list,
                      int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (status[order ? order[i] : i] == flag) list[atomicAdd(count, 1)] = i;
}

// dst[order[i]] = src[i]: per-condition by-products (outlet knot, outlet time) back to the caller's order
template <typename T>
__global__ void __launch_bounds__(256) unorder_kernel(const T* __restrict__ src, const int* __restrict__ order, int n, T* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[order ? order[i] : i] = src[i];
}

__global__ void __launch_bounds__(256) fill_int_kernel(int* __restrict__ dst, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}

}  // namespace pfr
