// Training step of the predictor MLPs: forward, MSE loss, backward and the Adam update of
//   TEMP_PRED_MODEL_TRAINING/temp_profile_model_training_2D.py:121-177   (2 -> 512 -> 512 -> 512 -> 800, ReLU)
//   TIME_PRED_MODEL_TRAINING/time_profile_model_training_4D.py:137-190   (4 -> 512 -> 512 -> 512 -> 800, ReLU)
// i.e. model(images) -> nn.MSELoss() -> optimizer.zero_grad(); loss.backward(); optimizer.step() with
// torch.optim.Adam(lr, betas (0.9, 0.999), eps 1e-8) for one mini-batch of at most 32 samples (the scripts' batch size).
//
// The problem is tiny (0.94 M parameters, 32 rows): nothing here is GEMM-shaped enough for tensor cores (M = 32) and the
// step is bound by launch latency and by streaming the 3.7 MB of parameters plus their two Adam moments once.  So:
//   * every kernel reads each weight exactly once (nn.Linear layout [out][in]); one warp per neuron / per input feature with
//     lane = batch row, so no kernel reduces across lanes and 512-800 warps are in flight per launch;
//   * the weight-gradient kernel never stores the gradient: each thread owns one weight, forms its gradient from the 32
//     batch rows and applies Adam in place (weight, m, v: 12 B read + 12 B written per parameter per step);
//   * the data-gradient of a layer is taken BEFORE that layer's weights are updated;
//   * all reductions have a fixed order (deterministic steps: same data, same weights, bit for bit).
// Twelve launches per step.
#pragma once
#include <cuda_runtime.h>

namespace pfr {
namespace mt {

constexpr int MAXB = 32;     // rows of a mini-batch
constexpr int HID = 512, OUT = 800;

// Activations are kept in two layouts: batch-major A[b][feature] (what the weight-gradient kernel reads, coalesced over the
// feature) and feature-major A_T[feature][32] (what the forward and data-gradient kernels read: the 32 batch rows of one
// feature are one 128-byte line, lane = batch row, no reduction across lanes anywhere).
//
// Both mat-vec kernels below walk the contraction index in chunks of 32: the block's 256 threads stage the chunk's
// [32][32 batch rows] operand in shared memory (double-buffered, the next chunk's loads are in flight while this one is
// consumed), each warp fetches its 32 weights of the chunk with one load and hands them round by shuffle.

// Y[b][n] = act(bias[n] + sum_k X(k, b) W[n][k]),  X(k, b) = Xin[k * sk + b * sb].  One warp per output neuron, lane = batch row.
template <bool kRelu>
__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* __restrict__ Xin, int sk, int sb, const float* __restrict__ W,
                                                         const float* __restrict__ bias, int B, int K, int N, float* __restrict__ Y,
                                                         float* __restrict__ Y_T) {
    __shared__ float xs[2][32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.x * 8 + warp, nc = n < N ? n : N - 1;
    auto fetch = [&](int k0, float (&r)[4]) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int e = threadIdx.x + 256 * q, k = k0 + (e >> 5), bb = e & 31;
            r[q] = (k < K && bb < B) ? Xin[(size_t)k * sk + (size_t)bb * sb] : 0.f;
        }
    };
    float r[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
    fetch(0, r);
    float wv = lane < K ? W[(size_t)nc * K + lane] : 0.f;
    for (int k0 = 0, c = 0; k0 < K; k0 += 32, c ^= 1) {
#pragma unroll
        for (int q = 0; q < 4; q++) { const int e = threadIdx.x + 256 * q; xs[c][e >> 5][e & 31] = r[q]; }
        __syncthreads();
        const float w_cur = wv;
        if (k0 + 32 < K) {
            fetch(k0 + 32, r);
            wv = k0 + 32 + lane < K ? W[(size_t)nc * K + k0 + 32 + lane] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; j++) acc[j & 3] = fmaf(__shfl_sync(0xffffffffu, w_cur, j), xs[c][j][lane], acc[j & 3]);
    }
    if (n >= N) return;
    float y = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + bias[n];
    if (kRelu) y = fmaxf(y, 0.f);
    const bool row = lane < B;
    Y_T[(size_t)n * MAXB + lane] = row ? y : 0.f;
    if (row) Y[(size_t)lane * N + n] = y;
}

// loss = mean((Y - T)^2) over B x N (nn.MSELoss default reduction), dY_T[n][b] = 2 (Y - T) / (B N).  One block, fixed-order tree.
__global__ void __launch_bounds__(1024) mse_kernel(const float* __restrict__ Y, const float* __restrict__ T, int B, int N,
                                                   float* __restrict__ dY_T, float* __restrict__ loss) {
    __shared__ double part[1024];
    const int total = B * N;
    const float scale = 2.0f / (float)total;
    double s = 0.0;
    for (int e = threadIdx.x; e < total; e += 1024) {
        const float d = Y[e] - T[e];
        if (dY_T) dY_T[(size_t)(e % N) * MAXB + e / N] = scale * d;
        s += (double)d * (double)d;
    }
    if (dY_T) {   // rows B..31 of the feature-major gradient are read by the data-gradient kernel: keep them zero
        for (int e = threadIdx.x; e < N * (MAXB - B); e += 1024) dY_T[(size_t)(e / (MAXB - B)) * MAXB + B + e % (MAXB - B)] = 0.f;
    }
    part[threadIdx.x] = s;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (threadIdx.x < w) part[threadIdx.x] += part[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = (float)(part[0] / (double)total);
}

// dA_T[k][b] = [A_T[k][b] > 0] * sum_n W[n][k] dY_T[n][b]  (A = the ReLU output that fed this layer).  One warp per input
// feature k, lane = batch row; the chunk of dY_T ([32 neurons][32 rows], contiguous) is staged as above; the lanes fetch
// W[n0 .. n0+31][k] at once (the eight warps of a block own eight consecutive k, i.e. the same 32-byte sectors).
__global__ void __launch_bounds__(256) linear_bwd_data_kernel(const float* __restrict__ dY_T, const float* __restrict__ W,
                                                              const float* __restrict__ A_T, int K, int N, float* __restrict__ dA_T) {
    __shared__ float gs[2][32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = blockIdx.x * 8 + warp, kc = k < K ? k : K - 1;
    auto fetch = [&](int n0, float (&r)[4]) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int e = threadIdx.x + 256 * q;
            r[q] = n0 + (e >> 5) < N ? dY_T[(size_t)n0 * MAXB + e] : 0.f;
        }
    };
    float r[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
    fetch(0, r);
    float wv = lane < N ? W[(size_t)lane * K + kc] : 0.f;
    for (int n0 = 0, c = 0; n0 < N; n0 += 32, c ^= 1) {
#pragma unroll
        for (int q = 0; q < 4; q++) { const int e = threadIdx.x + 256 * q; gs[c][e >> 5][e & 31] = r[q]; }
        __syncthreads();
        const float w_cur = wv;
        if (n0 + 32 < N) {
            fetch(n0 + 32, r);
            wv = n0 + 32 + lane < N ? W[(size_t)(n0 + 32 + lane) * K + kc] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; j++) acc[j & 3] = fmaf(__shfl_sync(0xffffffffu, w_cur, j), gs[c][j][lane], acc[j & 3]);
    }
    if (k >= K) return;
    const float g = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    dA_T[(size_t)k * MAXB + lane] = A_T[(size_t)k * MAXB + lane] > 0.f ? g : 0.f;
}

struct AdamArgs {
    float beta1, beta2, eps;
    float step_size;      // lr / (1 - beta1^t)
    float bc2_sqrt;       // sqrt(1 - beta2^t)
};

// torch.optim.Adam, single tensor path: m.lerp_(g, 1 - beta1); v.mul_(beta2).addcmul_(g, g, 1 - beta2);
// denom = v.sqrt() / bc2_sqrt + eps; p.addcdiv_(m, denom, value = -step_size)
__device__ __forceinline__ void adam_update(float g, float& p, float& m, float& v, const AdamArgs& a) {
    m = m + (g - m) * (1.0f - a.beta1);
    v = fmaf(g * g, 1.0f - a.beta2, v * a.beta2);
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p = p - a.step_size * (m / denom);
}

// One thread per input feature k and WN consecutive output neurons: the thread loads its 32 batch values X[b][k] once
// (coalesced over k), then for each of its neurons forms g = sum_b dY_T[n][b] X[b][k] (b ascending) and applies Adam to
// W[n][k] in place.  The thread with k == 0 also owns bias[n] (g = sum_b dY_T[n][b]).
constexpr int WN = 4;
__global__ void __launch_bounds__(256) linear_bwd_weight_adam_kernel(const float* __restrict__ dY_T, const float* __restrict__ X, int B,
                                                                     int K, int N, float* __restrict__ W, float* __restrict__ mW,
                                                                     float* __restrict__ vW, float* __restrict__ bias,
                                                                     float* __restrict__ mb, float* __restrict__ vb, AdamArgs a) {
    __shared__ float dy[WN][MAXB];
    const int nb = blockIdx.y * WN, k = blockIdx.x * 256 + threadIdx.x;
    if (threadIdx.x < WN * MAXB) {
        const int j = threadIdx.x / MAXB, b = threadIdx.x % MAXB;
        dy[j][b] = (nb + j < N && b < B) ? dY_T[(size_t)(nb + j) * MAXB + b] : 0.f;
    }
    __syncthreads();
    float x[MAXB];
#pragma unroll
    for (int b = 0; b < MAXB; b++) x[b] = (k < K && b < B) ? X[(size_t)b * K + k] : 0.f;
#pragma unroll
    for (int j = 0; j < WN; j++) {
        const int n = nb + j;
        if (n >= N) break;
        if (k < K) {
            float g = 0.f;
#pragma unroll
            for (int b = 0; b < MAXB; b++) g = fmaf(dy[j][b], x[b], g);
            const size_t e = (size_t)n * K + k;
            float p = W[e], m = mW[e], v = vW[e];
            adam_update(g, p, m, v, a);
            W[e] = p; mW[e] = m; vW[e] = v;
        }
        if (k == 0) {
            float g = 0.f;
            for (int b = 0; b < B; b++) g += dy[j][b];
            float p = bias[n], m = mb[n], v = vb[n];
            adam_update(g, p, m, v, a);
            bias[n] = p; mb[n] = m; vb[n] = v;
        }
    }
}

}  // namespace mt
}  // namespace pfr
