// Microbenchmark: a 9x9 mat-vec pair per step (162 DFMA per thread) with the 162 coefficients coming from
//   (a) a block-shared copy, broadcast ld.volatile.shared.v2.f64 (what bs23_kernel does), or
//   (b) the kernel-parameter constant bank through uniform registers (LDCU.64 + DFMA R, R, UR, R), the coefficient block
//       duplicated and the copy toggled every step so that neither nvcc nor ptxas can hoist the loads out of the loop.
// Prints ns per step per SM-resident warp set and the implied DFMA rate.   nvcc -arch=sm_100a -O3 -o coef_paths coef_paths.cu
#include <cstdio>
#include <cuda_runtime.h>
struct Coef { double nu[2][9][10]; double w[2][9][10]; };
struct CoefS { double nu[9][10]; double w[9][10]; };

__device__ __forceinline__ void lds2(const double* q, double& a, double& b, double after) {
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"((unsigned)__cvta_generic_to_shared(q)), "d"(after));
}

__global__ void __launch_bounds__(128, 12) k_uniform(const __grid_constant__ Coef c, const double* __restrict__ in, double* __restrict__ out, int steps) {
    double y[9], z[9];
    for (int i = 0; i < 9; i++) y[i] = in[i * 128 + threadIdx.x];
    int buf = 0;
    for (int s = 0; s < steps; s++) {
#pragma unroll
        for (int j = 0; j < 9; j++) z[j] = 0.0;
#pragma unroll
        for (int kk = 0; kk < 9; kk++) {
            const double l = y[kk] * 0.999 + 1e-3;
#pragma unroll
            for (int j = 0; j < 9; j++) z[j] = fma(c.nu[buf][kk][j], l, z[j]);
        }
#pragma unroll
        for (int j = 0; j < 9; j++) y[j] = 0.0;
#pragma unroll
        for (int kk = 0; kk < 9; kk++) {
            const double r = z[kk] * 0.5 + 0.25;
#pragma unroll
            for (int j = 0; j < 9; j++) y[j] = fma(c.w[buf][kk][j], r, y[j]);
        }
        buf ^= 1;
    }
    for (int i = 0; i < 9; i++) out[i * 128 + threadIdx.x + blockIdx.x * 9 * 128] = y[i];
}

// (c) as (b) with TWO accumulator sets per mat-vec (rows of even and odd index), added at the end: 18 independent FMA chains per
//     thread instead of 9 -- a dependent DFMA has a long latency on this part and three warps per scheduler x 9 chains do not cover it
__global__ void __launch_bounds__(128, 3) k_uniform2(const __grid_constant__ Coef c, const double* __restrict__ in, double* __restrict__ out, int steps) {
    double y[9], z[9], z2[9];
    for (int i = 0; i < 9; i++) y[i] = in[i * 128 + threadIdx.x];
    int buf = 0;
    for (int s = 0; s < steps; s++) {
#pragma unroll
        for (int j = 0; j < 9; j++) { z[j] = 0.0; z2[j] = 0.0; }
#pragma unroll
        for (int kk = 0; kk < 9; kk++) {
            const double l = y[kk] * 0.999 + 1e-3;
#pragma unroll
            for (int j = 0; j < 9; j++) {
                if (kk & 1) z2[j] = fma(c.nu[buf][kk][j], l, z2[j]);
                else z[j] = fma(c.nu[buf][kk][j], l, z[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 9; j++) { z[j] += z2[j]; y[j] = 0.0; z2[j] = 0.0; }
#pragma unroll
        for (int kk = 0; kk < 9; kk++) {
            const double r = z[kk] * 0.5 + 0.25;
#pragma unroll
            for (int j = 0; j < 9; j++) {
                if (kk & 1) z2[j] = fma(c.w[buf][kk][j], r, z2[j]);
                else y[j] = fma(c.w[buf][kk][j], r, y[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 9; j++) y[j] += z2[j];
        buf ^= 1;
    }
    for (int i = 0; i < 9; i++) out[i * 128 + threadIdx.x + blockIdx.x * 9 * 128] = y[i];
}

__global__ void __launch_bounds__(128, 3) k_shared(const __grid_constant__ CoefS c, const double* __restrict__ in, double* __restrict__ out, int steps) {
    __shared__ __align__(16) CoefS sc;
    for (int e = threadIdx.x; e < 90; e += 128) { sc.nu[e / 10][e % 10] = c.nu[e / 10][e % 10]; sc.w[e / 10][e % 10] = c.w[e / 10][e % 10]; }
    __syncthreads();
    double y[9], z[9];
    for (int i = 0; i < 9; i++) y[i] = in[i * 128 + threadIdx.x];
    for (int s = 0; s < steps; s++) {
#pragma unroll
        for (int j = 0; j < 9; j++) z[j] = 0.0;
#pragma unroll
        for (int kk = 0; kk < 9; kk++) {
            const double l = y[kk] * 0.999 + 1e-3;
            double cc[10];
#pragma unroll
            for (int e = 0; e < 5; e++) lds2(&sc.nu[kk][2 * e], cc[2 * e], cc[2 * e + 1], l);
#pragma unroll
            for (int j = 0; j < 9; j++) z[j] = fma(cc[j], l, z[j]);
        }
#pragma unroll
        for (int j = 0; j < 9; j++) y[j] = 0.0;
#pragma unroll
        for (int kk = 0; kk < 9; kk++) {
            const double r = z[kk] * 0.5 + 0.25;
            double cc[10];
#pragma unroll
            for (int e = 0; e < 5; e++) lds2(&sc.w[kk][2 * e], cc[2 * e], cc[2 * e + 1], r);
#pragma unroll
            for (int j = 0; j < 9; j++) y[j] = fma(cc[j], r, y[j]);
        }
    }
    for (int i = 0; i < 9; i++) out[i * 128 + threadIdx.x + blockIdx.x * 9 * 128] = y[i];
}

// (d) fixed-address 16-byte uniform loads (LDCU.128, two coefficients per load) from a __constant__ block, issued as volatile
//     inline PTX so that nvcc cannot hoist them out of the step loop; ptxas keeps as many as fit in uniform registers and
//     re-loads the rest every step.  (With an index in the address -- the copy toggle of (b) -- ptxas only ever emits LDCU.64.)
extern "C" { __constant__ CoefS pfr_cc; }
template <int OFF>
__device__ __forceinline__ void ldc2(double& a, double& b) {
    asm volatile("ld.const.v2.f64 {%0, %1}, [pfr_cc+%2];" : "=d"(a), "=d"(b) : "n"(OFF));
}
template <int BASE, int KK, int J>
struct RowC {
    static __device__ __forceinline__ void run(double l, double (&z)[10]) {
        double a, b;
        ldc2<BASE + (KK * 10 + J) * 8>(a, b);
        z[J] = fma(a, l, z[J]);
        if (J + 1 < 9) z[J + 1] = fma(b, l, z[J + 1]);
        if constexpr (J + 2 < 10) RowC<BASE, KK, J + 2>::run(l, z);
    }
};
__global__ void __launch_bounds__(128, 3) k_const128(const double* __restrict__ in, double* __restrict__ out, int steps) {
    double y[10], z[10];
    for (int i = 0; i < 9; i++) y[i] = in[i * 128 + threadIdx.x];
    for (int s = 0; s < steps; s++) {
#pragma unroll
        for (int j = 0; j < 10; j++) z[j] = 0.0;
#define ROWN(KK) RowC<0, KK, 0>::run(y[KK] * 0.999 + 1e-3, z);
        ROWN(0) ROWN(1) ROWN(2) ROWN(3) ROWN(4) ROWN(5) ROWN(6) ROWN(7) ROWN(8)
#pragma unroll
        for (int j = 0; j < 10; j++) y[j] = 0.0;
#define ROWW(KK) RowC<720, KK, 0>::run(z[KK] * 0.5 + 0.25, y);
        ROWW(0) ROWW(1) ROWW(2) ROWW(3) ROWW(4) ROWW(5) ROWW(6) ROWW(7) ROWW(8)
    }
    for (int i = 0; i < 9; i++) out[i * 128 + threadIdx.x + blockIdx.x * 9 * 128] = y[i];
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = 3 * sms, steps = 20000;
    double *in, *out;
    cudaMalloc(&in, 9 * 128 * sizeof(double));
    cudaMalloc(&out, (size_t)grid * 9 * 128 * sizeof(double));
    cudaMemset(in, 0, 9 * 128 * sizeof(double));
    Coef c; CoefS cs;
    for (int b = 0; b < 2; b++) for (int k = 0; k < 9; k++) for (int j = 0; j < 10; j++) { c.nu[b][k][j] = 0.01 * (k + j); c.w[b][k][j] = 0.02 * (k - j); }
    for (int k = 0; k < 9; k++) for (int j = 0; j < 10; j++) { cs.nu[k][j] = c.nu[0][k][j]; cs.w[k][j] = c.w[0][k][j]; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFree(out);
    cudaMalloc(&out, (size_t)grid * 4 * 9 * 128 * sizeof(double));
    cudaMemcpyToSymbol(pfr_cc, &cs, sizeof(cs));
    for (int variant = 0; variant < 7; variant++) {
        const int g = variant == 3 ? 2 * grid : (variant == 4 ? 4 * grid : (variant == 5 ? grid / 3 : grid));   // 6, 12 and 1 CTA(s) per SM
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (variant == 0) k_shared<<<g, 128>>>(cs, in, out, steps);
            else if (variant == 2) k_uniform2<<<g, 128>>>(c, in, out, steps);
            else if (variant == 6) k_const128<<<g, 128>>>(in, out, steps);
            else k_uniform<<<g, 128>>>(c, in, out, steps);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double dfma = 162.0 * 128 * g * (double)steps;
            printf("%s: %.3f ms, %.2f TFLOP/s of DFMA, %s\n", variant == 0 ? "shared broadcast (LDS.128)" : (variant == 1 ? "uniform registers (LDCU.64)" : (variant == 2 ? "uniform registers, two accumulator sets" : (variant == 3 ? "uniform, 6 CTAs/SM" : (variant == 4 ? "uniform, 12 CTAs/SM" : (variant == 5 ? "uniform, 1 CTA/SM" : "fixed-address LDCU.128 from __constant__"))))), ms, 2 * dfma / (ms * 1e-3) / 1e12,
                   cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
