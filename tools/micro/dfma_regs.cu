// Microbenchmark: what does the FP64 pipe sustain when every DFMA reads three DISTINCT register operands (no operand reuse),
// with no loads at all?  9 accumulators x 9 rows per "mat-vec", coefficients held in 18 registers.  3 CTAs of 128 threads per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int kMode>
__global__ void __launch_bounds__(128, 3) k(const double* __restrict__ in, double* __restrict__ out, int steps) {
    double y[9], z[9], c[18];
    for (int i = 0; i < 9; i++) y[i] = in[i * 128 + threadIdx.x];
    for (int i = 0; i < 18; i++) c[i] = in[(i % 9) * 128 + threadIdx.x] * 1e-3 + 0.01 * i;
    for (int s = 0; s < steps; s++) {
#pragma unroll
        for (int j = 0; j < 9; j++) z[j] = 0.0;
#pragma unroll
        for (int kk = 0; kk < 9; kk++) {
            const double l = kMode == 0 ? y[kk] * 0.999 + 1e-3 : y[kk];
#pragma unroll
            for (int j = 0; j < 9; j++) z[j] = fma(c[(j + kk) % 18], l, z[j]);
        }
#pragma unroll
        for (int j = 0; j < 9; j++) y[j] = 0.0;
#pragma unroll
        for (int kk = 0; kk < 9; kk++) {
            const double r = kMode == 0 ? z[kk] * 0.5 + 0.25 : z[kk];
#pragma unroll
            for (int j = 0; j < 9; j++) y[j] = fma(c[(j + 2 * kk + 1) % 18], r, y[j]);
        }
    }
    for (int i = 0; i < 9; i++) out[i * 128 + threadIdx.x + blockIdx.x * 9 * 128] = y[i];
}
// the classic peak loop: 8 independent chains x = fma(x, a, b) with two loop-invariant operands
__global__ void __launch_bounds__(128, 3) peak(const double* __restrict__ in, double* __restrict__ out, int steps) {
    double x[8];
    for (int i = 0; i < 8; i++) x[i] = in[i * 128 + threadIdx.x];
    const double a = in[threadIdx.x] + 0.999, b = in[128 + threadIdx.x] + 1e-3;
    for (int s = 0; s < steps; s++) {
#pragma unroll
        for (int u = 0; u < 20; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
    }
    for (int i = 0; i < 8; i++) out[i * 128 + threadIdx.x + blockIdx.x * 9 * 128] = x[i];
}
int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = 3 * sms, steps = 20000;
    double *in, *out;
    cudaMalloc(&in, 18 * 128 * sizeof(double));
    cudaMalloc(&out, (size_t)grid * 9 * 128 * sizeof(double));
    cudaMemset(in, 0, 18 * 128 * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int variant = 0; variant < 3; variant++)
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (variant == 0) k<0><<<grid, 128>>>(in, out, steps);
            else if (variant == 1) k<1><<<grid, 128>>>(in, out, steps);
            else peak<<<grid, 128>>>(in, out, steps);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double n = (variant == 2 ? 160.0 : (variant == 0 ? 180.0 : 162.0)) * 128 * grid * (double)steps;
            printf("%s: %.3f ms, %.2f TFLOP/s of FP64 instructions, %s\n", variant == 0 ? "register mat-vecs (+18 scalar FMAs)" : (variant == 1 ? "register mat-vecs" : "peak loop (two invariant operands)"),
                   ms, 2 * n / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
