// Tensor-core path of the predictor MLPs: tcgen05 (5th-gen tensor cores, TMEM accumulator) with an
// error-compensated 3xTF32 split, operands staged by TMA into 128-byte-swizzled shared memory.
//
// Why a split: the grids the MLPs produce go through enforce_strict, whose keep/repair decisions hang on the
// last bits of a float32 result, so a plain TF32 product (10-bit mantissa) is not acceptable.
// Each float32 operand x is written once as  hi = rn_tf32(x),  lo = rn_tf32(x - hi)  (both exactly
// representable in TF32, round-to-nearest so the dropped lo*lo term is unbiased) and the product is
//   a*b ~= hi_a*hi_b + hi_a*lo_b + lo_a*hi_b     (relative error <= 2^-21 per term, float32 accumulation in TMEM)
// i.e. three tcgen05.mma passes over the same shared-memory stage.
//
// Accumulation: tcgen05 adds each K=8 product group into the float32 TMEM accumulator with truncation, so the error
// of one long chain grows linearly with the number of instructions (measured 1e-5 relative after 3 x 64 of them).  The
// kernel therefore keeps FOUR accumulators per tile in TMEM -- the hi*hi products of three thirds of K and one for all
// the (2^-11 times smaller) cross terms -- and the epilogue adds the four in float32 round-to-nearest: ~21 truncating
// adds per chain, the same error level as a 512-term FP32 FFMA chain.
//
// Layout: activations are K-major here ([condition][512], hi and lo arrays), weights keep nn.Linear's [out][512]
// (K-major as well), so both operands use the canonical K-major SWIZZLE_128B tile (8 rows x 128 B atoms, SBO 1024 B)
// that TMA writes and the UMMA shared-memory descriptor reads.  D[128 conditions x BN outputs] lives in TMEM; the four
// epilogue warps read it back with tcgen05.ld (lane = condition), add the bias, apply ReLU + the hi/lo split (hidden
// layers) or the float32 un-scaling (output layer, written straight into the [801][n] grid rows, coalesced over
// conditions).
//
// Warp roles (192 threads, 1 CTA/SM): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer (one elected lane),
// warps 2-5 epilogue.  Two 96 KB stages; every mbarrier wait is bounded and traps instead of hanging the device.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pfr {
namespace tc {

constexpr int BM = 128;      // conditions per CTA tile = UMMA M
constexpr int BK = 32;       // floats per k-block: 128 bytes, one swizzle row
constexpr int UMMA_K = 8;    // tf32: 32 bytes per instruction
constexpr int STAGES = 3;
constexpr int THREADS = 192;
constexpr int KDIM = 512;
constexpr int TMEM_COLS = 512;
constexpr int BN = 128;     // outputs per CTA tile = UMMA N; four accumulators of BN columns fill the 512 TMEM columns

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s: a protocol error must not hang the device
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = Blackwell):
// start address >> 4 | LBO (unused for swizzled K-major) | SBO = 1024 B between 8-row groups | layout type 2
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor: c F32 (1<<4), a/b TF32 (2<<7, 2<<10), both K-major, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float rn_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

struct GemmArgs {
    const float* bias;     // [n_total]
    float* out_hi;         // hidden: [rows][512] hi part; final: grid row 1, [800][out_ld]
    float* out_lo;         // hidden: lo part
    size_t out_ld;         // final: leading dimension of the grid (n)
    int n_valid;           // final: outputs to store (800)
    int m_valid;           // final: conditions to store
    float span, omin;      // final: v = out * span + omin
};

// one CTA: D[128 x BN] = sum over 16 k-blocks of (Ahi Bhi^T + Ahi Blo^T + Alo Bhi^T)
template <bool kFinal>
__global__ void __launch_bounds__(THREADS, 1)
mlp_tc_gemm_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                   const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, const GemmArgs g) {
    constexpr uint32_t A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_smem;
    constexpr int NKB = KDIM / BK;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < NKB; kb++) {
                const int s = kb % STAGES;
                mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
                unsigned char* st = smem + (size_t)s * STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                tma_load_2d(st, &mapAhi, &full_bar[s], kb * BK, m0);
                tma_load_2d(st + A_BYTES, &mapAlo, &full_bar[s], kb * BK, m0);
                tma_load_2d(st + 2 * A_BYTES, &mapBhi, &full_bar[s], kb * BK, n0);
                tma_load_2d(st + 2 * A_BYTES + B_BYTES, &mapBlo, &full_bar[s], kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(BN);
            for (int kb = 0; kb < NKB; kb++) {
                const int s = kb % STAGES;
                mbar_wait(&full_bar[s], (kb / STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = smem_u32(smem + (size_t)s * STAGE_BYTES), a_lo = a_hi + A_BYTES;
                const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                for (int ks = 0; ks < BK / UMMA_K; ks++) {
                    const uint32_t off = ks * UMMA_K * 4;
                    const int kg = kb * (BK / UMMA_K) + ks;          // 0..63
                    const int part = kg < 22 ? 0 : (kg < 43 ? 1 : 2);  // three thirds of K for the leading term
                    const bool first = kg == 0 || kg == 22 || kg == 43;
                    // cross terms -> accumulator 0 ; leading term -> accumulator 1 + part
                    umma_tf32(tmem_d, umma_desc_sw128(a_lo + off), umma_desc_sw128(b_hi + off), idesc, kg != 0);
                    umma_tf32(tmem_d, umma_desc_sw128(a_hi + off), umma_desc_sw128(b_lo + off), idesc, 1);
                    umma_tf32(tmem_d + (uint32_t)((1 + part) * BN), umma_desc_sw128(a_hi + off), umma_desc_sw128(b_hi + off), idesc, !first);
                }
                umma_commit(&empty_bar[s]);  // the stage is free once these MMAs have read it
            }
            umma_commit(&tmem_full_bar);
        }
    } else {
        // epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 ; thread = one condition row
        const int q = warp & 3;
        const int m = m0 + 32 * q + lane;
        mbar_wait(&tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c = 0; c < BN / 16; c++) {
            uint32_t v[4][16];
#pragma unroll
            for (int acc = 0; acc < 4; acc++) {
                const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * BN + c * 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[acc][0]), "=r"(v[acc][1]), "=r"(v[acc][2]), "=r"(v[acc][3]), "=r"(v[acc][4]), "=r"(v[acc][5]),
                      "=r"(v[acc][6]), "=r"(v[acc][7]), "=r"(v[acc][8]), "=r"(v[acc][9]), "=r"(v[acc][10]), "=r"(v[acc][11]),
                      "=r"(v[acc][12]), "=r"(v[acc][13]), "=r"(v[acc][14]), "=r"(v[acc][15])
                    : "r"(taddr));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int o0 = n0 + c * 16;
            float x[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                // float32 round-to-nearest sum of the three K-thirds, then the cross terms, then the bias
                const float main = (__uint_as_float(v[1][j]) + __uint_as_float(v[2][j])) + __uint_as_float(v[3][j]);
                const int o = o0 + j;
                x[j] = (main + __uint_as_float(v[0][j])) + ((kFinal && o >= g.n_valid) ? 0.f : __ldg(&g.bias[o]));
            }
            if (!kFinal) {
                float hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const float r = fmaxf(x[j], 0.f);
                    hi[j] = rn_tf32(r);
                    lo[j] = rn_tf32(r - hi[j]);
                }
                float4* ph = reinterpret_cast<float4*>(g.out_hi + (size_t)m * KDIM + o0);
                float4* pl = reinterpret_cast<float4*>(g.out_lo + (size_t)m * KDIM + o0);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    ph[j] = make_float4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                    pl[j] = make_float4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                }
            } else if (m < g.m_valid) {  // m is chunk-local, and so is the grid pointer
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const int o = o0 + j;
                    if (o < g.n_valid) g.out_hi[(size_t)o * g.out_ld + m] = __fadd_rn(__fmul_rn(x[j], g.span), g.omin);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
    }
}

// layer 1 (K <= 4) fused with the input scaling, written K-major as a TF32 hi/lo pair
__global__ void __launch_bounds__(256)
mlp_tc_layer1_kernel(const float* __restrict__ W1, const float* __restrict__ b1, int in_dim, float lo0, float lo1, float lo2,
                     float lo3, float sp0, float sp1, float sp2, float sp3, float fullL, float fullU,
                     const float* __restrict__ T, const float* __restrict__ P, const float* __restrict__ L,
                     const float* __restrict__ U, int m_valid, int rows, float* __restrict__ Hhi, float* __restrict__ Hlo) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows * KDIM) return;
    const int m = (int)(idx / KDIM), k = (int)(idx % KDIM);
    const int ms = m < m_valid ? m : m_valid - 1;
    float x[4];
    x[0] = __fdiv_rn(__fsub_rn(T[ms], lo0), sp0);
    x[1] = __fdiv_rn(__fsub_rn(P[ms], lo1), sp1);
    x[2] = __fdiv_rn(__fsub_rn(L ? L[ms] : fullL, lo2), sp2);
    x[3] = __fdiv_rn(__fsub_rn(U ? U[ms] : fullU, lo3), sp3);
    float acc = 0.f;
    for (int i = 0; i < in_dim; i++) acc = fmaf(x[i], W1[k * in_dim + i], acc);
    acc = fmaxf(acc + b1[k], 0.f);
    const float hi = rn_tf32(acc);
    Hhi[idx] = hi;
    Hlo[idx] = rn_tf32(acc - hi);
}

}  // namespace tc
}  // namespace pfr
