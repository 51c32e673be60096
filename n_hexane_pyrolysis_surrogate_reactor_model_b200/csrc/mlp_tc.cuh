// Tensor-core path of the predictor MLPs: tcgen05 (5th-gen tensor cores, TMEM accumulator) with an
// error-compensated three-product operand split, operands staged by TMA into 128-byte-swizzled shared memory.
// Two operand formats, chosen per MLP handle (pfr_mlp_set_mode): TF32 pairs (round 1) and FLOAT16 pairs (round 2, default).
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
// FLOAT16 pairs (kHalf): an 11-bit significand is an 11-bit significand -- float16 carries exactly as many bits as TF32, in half
// the bytes, and kind::f16 issues K = 16 per instruction where kind::tf32 issues K = 8.  The split is
//   hi = rn_f16(x),  lo' = rn_f16((x - hi) * 2^11)      (x - hi is exact in float32, the scaling keeps lo' a NORMAL float16
// down to |x| ~ 1e-7; below that its absolute error is < 2^-35), a*b ~= hi_a*hi_b + 2^-11 (hi_a*lo'_b + lo'_a*hi_b):
// the cross terms have their own accumulator anyway, so the 2^-11 is one exact multiply in the epilogue.  Same three
// products, same accumulators, HALF the shared-memory bytes per product and half the instructions: the GEMM is bound by the
// shared-memory bandwidth of SS-operand MMA and, in the sustained pass, by board power, and both scale
// with the bytes.  The K-thirds are 11 instructions long instead of 21, so the truncating accumulation drifts less.  Range:
// float16 tops out at 65 504 -- the shipped predictors' activations stay below 2 and their weights below 0.35; a weight
// beyond the range makes pfr_mlp_set_mode refuse this mode, an activation beyond it turns into inf -> NaN in the grid (loud).
//
// Layout: activations are K-major here ([condition][512], hi and lo arrays), weights keep nn.Linear's [out][512]
// (K-major as well), so both operands use the canonical K-major SWIZZLE_128B tile (8 rows x 128 B atoms, SBO 1024 B)
// that TMA writes and the UMMA shared-memory descriptor reads.  D[128 conditions x BN outputs] lives in TMEM; the four
// epilogue warps read it back with tcgen05.ld (lane = condition), add the bias, apply ReLU + the hi/lo split (hidden
// layers) or the float32 un-scaling (output layer, written straight into the [801][n] grid rows, coalesced over
// conditions).
//
// Warp roles (320 threads, one persistent CTA per SM): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer (one
// elected lane), warps 2-9 epilogue.  Three 64 KB stages + 32 KB of epilogue slabs; every mbarrier wait is bounded and
// traps instead of hanging the device.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pfr {
namespace tc {

constexpr int BM = 128;      // conditions per CTA tile = UMMA M
constexpr int STAGES = 3;
#ifndef PFR_TC_EPI_WARPS
#define PFR_TC_EPI_WARPS 8
#endif
constexpr int EPI_WARPS = PFR_TC_EPI_WARPS;   // 4: one warp per TMEM lane quarter; 8: two, each draining half of the columns
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int KDIM = 512;
constexpr int TMEM_COLS = 512;
constexpr int BN = 128;     // outputs per CTA tile = UMMA N; four accumulators of BN columns fill the 512 TMEM columns

// operand format: what one 128-byte swizzle row / one 32-byte instruction slice holds, and where the K-thirds are cut
template <bool kHalf>
struct Operand {
    static constexpr int ESIZE = kHalf ? 2 : 4;
    static constexpr int BK = 128 / ESIZE;          // elements per k-block: 128 bytes, one swizzle row
    static constexpr int UMMA_K = 32 / ESIZE;       // elements per instruction: 32 bytes
    static constexpr int NKB = KDIM / BK;           // k-blocks (= pipeline stages filled) per tile
    static constexpr int KG = NKB * (BK / UMMA_K);  // instruction groups per tile: 64 | 32
    static constexpr int CUT1 = KG / 3 + 1, CUT2 = 2 * KG / 3 + 1;   // 22, 43 | 11, 22
};

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
// cute::UMMA::InstrDescriptor: c F32 (1<<4), a/b format at bits 7 / 10 (kind::tf32: TF32 = 2; kind::f16: F16 = 0), both K-major,
// N>>3 at bit 17, M>>4 at bit 24
template <bool kHalf>
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
    return (1u << 4) | ((kHalf ? 0u : 2u) << 7) | ((kHalf ? 0u : 2u) << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
template <bool kHalf>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (kHalf) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
            : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// packed float32 pairs (sm_100: add.rn.f32x2 adds two IEEE float32 lanes with one instruction)
__device__ __forceinline__ uint64_t pack2(uint32_t a, uint32_t b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, uint32_t& a, uint32_t& b) { asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t sub_f32x2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// the float16 pair of two float32 values: hi = rn_f16(r), lo' = rn_f16((r - hi) * 2^11), two values per packed word
__device__ __forceinline__ void split_f16x2(float r0, float r1, uint32_t& hi2, uint32_t& lo2) {
    const __half2 h = __floats2half2_rn(r0, r1);
    const float2 hf = __half22float2(h);
    uint32_t d0, d1;
    unpack2(mul_f32x2(sub_f32x2(pack2(__float_as_uint(r0), __float_as_uint(r1)), pack2(__float_as_uint(hf.x), __float_as_uint(hf.y))),
                      pack2(0x45000000u, 0x45000000u)), d0, d1);   // 2048.0f
    const __half2 l = __floats2half2_rn(__uint_as_float(d0), __uint_as_float(d1));
    hi2 = *reinterpret_cast<const uint32_t*>(&h);
    lo2 = *reinterpret_cast<const uint32_t*>(&l);
}

// Round to TF32 (10 explicit mantissa bits), nearest with ties away from zero -- what cvt.rna.tf32.f32 returns for every finite
// input, as two integer instructions: the PTX conversion expands to ~3.2 SASS instructions with its NaN handling (FSETP / SEL),
// and the hidden-layer epilogue, which is bound by instruction issue, performs 256 of them per thread and tile.  (Infinity
// stays infinity, the largest finite values round up to it, a NaN stays a NaN with some payload bits cleared.)
__device__ __forceinline__ float rn_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

struct GemmArgs {
    const float* bias;     // [n_total]
    float* out_hi;         // hidden: [rows][512] hi part; final: grid row 1, [800][out_ld]
    float* out_lo;         // hidden: lo part
    size_t out_ld;         // final: leading dimension of the grid (n)
    int n_valid;           // final: outputs to store (800)
    int m_valid;           // final: conditions to store
    float span, omin;      // final: v = out * span + omin
    int n_tiles, m_tiles;  // tiles along the outputs / the conditions; tile t -> (t / n_tiles, t % n_tiles)
    unsigned long long* trace;  // development (-DPFR_TC_TRACE, tools/trace_tc.py): [tiles][8] %globaltimer stamps, else nullptr
};

constexpr uint32_t A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // rows of 128 bytes in both formats
constexpr int EPI_COLS = 16;                               // accumulator columns per epilogue slab
constexpr uint32_t SLAB_BYTES = 32 * EPI_COLS * 4;         // one warp's [32 conditions][16 outputs] slab of one array
constexpr uint32_t STAGING_BYTES = EPI_WARPS * 2 * SLAB_BYTES;   // every epilogue warp x (hi, lo)
constexpr size_t SMEM_DYN = (size_t)STAGES * STAGE_BYTES + STAGING_BYTES + 1024;

#ifdef PFR_TC_TRACE
__device__ __forceinline__ unsigned long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory"); return t; }
#define TC_STAMP(ptr, slot) do { if (ptr) (ptr)[slot] = gtime_ns(); } while (0)
#else
#define TC_STAMP(ptr, slot) do { } while (0)
#endif
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory"); }   // the epilogue warps
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the four accumulators' columns [16 c, 16 c + 16) of this warp's 32 TMEM lanes -> v[acc][j]  (asynchronous until wait::ld)
__device__ __forceinline__ void tmem_load_slab(uint32_t (&v)[4][16], uint32_t tmem_d, int q, int c) {
#pragma unroll
    for (int acc = 0; acc < 4; acc++) {
        const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * BN + c * EPI_COLS);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[acc][0]), "=r"(v[acc][1]), "=r"(v[acc][2]), "=r"(v[acc][3]), "=r"(v[acc][4]), "=r"(v[acc][5]), "=r"(v[acc][6]),
              "=r"(v[acc][7]), "=r"(v[acc][8]), "=r"(v[acc][9]), "=r"(v[acc][10]), "=r"(v[acc][11]), "=r"(v[acc][12]),
              "=r"(v[acc][13]), "=r"(v[acc][14]), "=r"(v[acc][15])
            : "r"(taddr));
    }
}

// Persistent CTA (one per SM): D[128 x BN] = sum over 16 k-blocks of (Ahi Bhi^T + Ahi Blo^T + Alo Bhi^T) for tiles
// blockIdx.x, blockIdx.x + gridDim.x, ...  The TMA producer runs ahead across tile boundaries (the next tile's first
// stages load while this tile's epilogue drains TMEM); the MMA warp starts the next tile as soon as the epilogue has
// read the last accumulator slab.  Hidden layers leave through a warp-private shared-memory transpose so that one store
// instruction writes 8 rows x 64 contiguous bytes instead of 32 scattered 16-byte pieces; the output layer stores
// straight into the knot-major grid, where a warp's 32 conditions are contiguous.
template <bool kFinal, bool kHalf>
__global__ void __launch_bounds__(THREADS, 1)
mlp_tc_gemm_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                   const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, const GemmArgs g) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* staging = smem + (size_t)STAGES * STAGE_BYTES;
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar, tmem_empty_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float bias_s[BN];   // read as float2

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = g.n_tiles * g.m_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        mbar_init(&tmem_empty_bar, EPI_WARPS);   // one arrival per epilogue warp
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
    using Op = Operand<kHalf>;
    constexpr int NKB = Op::NKB, BK = Op::BK;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int n0 = (tile % g.n_tiles) * BN, m0 = (tile / g.n_tiles) * BM;
                for (int kb = 0; kb < NKB; kb++, it++) {
                    const uint32_t s = it % STAGES;
                    mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    unsigned char* st = smem + (size_t)s * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                    tma_load_2d(st, &mapAhi, &full_bar[s], kb * BK, m0);
                    tma_load_2d(st + A_BYTES, &mapAlo, &full_bar[s], kb * BK, m0);
                    tma_load_2d(st + 2 * A_BYTES, &mapBhi, &full_bar[s], kb * BK, n0);
                    tma_load_2d(st + 2 * A_BYTES + B_BYTES, &mapBlo, &full_bar[s], kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc<kHalf>(BN);
            uint32_t it = 0, lt = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x, lt++) {
                unsigned long long* tr = g.trace ? g.trace + 8 * (size_t)tile : nullptr;
                (void)tr;
                TC_STAMP(tr, 0);
                if (lt > 0) {   // the epilogue has read every accumulator of the previous tile
                    mbar_wait(&tmem_empty_bar, (lt - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                TC_STAMP(tr, 1);
                for (int kb = 0; kb < NKB; kb++, it++) {
                    const uint32_t s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(smem + (size_t)s * STAGE_BYTES), a_lo = a_hi + A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) {                       // four 32-byte instruction slices per 128-byte row
                        const uint32_t off = ks * 32;
                        const int kg = kb * 4 + ks;                          // 0 .. Op::KG - 1
                        const int part = kg < Op::CUT1 ? 0 : (kg < Op::CUT2 ? 1 : 2);  // three thirds of K for the leading term
                        const bool first = kg == 0 || kg == Op::CUT1 || kg == Op::CUT2;
                        // cross terms -> accumulator 0 ; leading term -> accumulator 1 + part
                        umma_ss<kHalf>(tmem_d, umma_desc_sw128(a_lo + off), umma_desc_sw128(b_hi + off), idesc, kg != 0);
                        umma_ss<kHalf>(tmem_d, umma_desc_sw128(a_hi + off), umma_desc_sw128(b_lo + off), idesc, 1);
                        umma_ss<kHalf>(tmem_d + (uint32_t)((1 + part) * BN), umma_desc_sw128(a_hi + off), umma_desc_sw128(b_hi + off), idesc, !first);
                    }
                    umma_commit(&empty_bar[s]);  // the stage is free once these MMAs have read it
                }
                umma_commit(&tmem_full_bar);
                TC_STAMP(tr, 2);
            }
        }
    } else {
        // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31 ; thread = one condition row.  With eight warps, warps w and w + 4
        // share a lane quarter and drain one half of the accumulator columns each: an epilogue warp is a chain of dependent
        // TMEM loads, packed adds, conversions and a shared-memory transpose, so one warp per scheduler ran at a fraction of
        // an instruction per clock while the MMA warp waited for TMEM (4.6 us per tile); two per scheduler overlap.
        const int q = warp & 3;
        const int e = threadIdx.x - 64;            // 0 .. 32 * EPI_WARPS - 1
        constexpr int SLABS = BN / EPI_COLS / (EPI_WARPS / 4);   // slabs of 16 columns per warp
        const int c_first = ((warp - 2) >> 2) * SLABS;
        unsigned char* mine = staging + (size_t)(warp - 2) * 2 * SLAB_BYTES;
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, lt++) {
            const int n0 = (tile % g.n_tiles) * BN, m0 = (tile / g.n_tiles) * BM;
            const int m = m0 + 32 * q + lane;
            const float bias_mine = (e >= BN || (kFinal && n0 + e >= g.n_valid)) ? 0.f : __ldg(&g.bias[n0 + e]);
            epi_bar_sync();                        // everybody is done with the previous tile's bias
            if (e < BN) bias_s[e] = bias_mine;
            epi_bar_sync();
            unsigned long long* tr = (g.trace && e == 0) ? g.trace + 8 * (size_t)tile : nullptr;
            (void)tr;
            TC_STAMP(tr, 3);
            mbar_wait(&tmem_full_bar, lt & 1);
            TC_STAMP(tr, 4);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // slabs of 16 accumulator columns, software-pipelined: slab c + 1 is in flight while slab c is processed
            uint32_t v[2][4][16];
            tmem_load_slab(v[0], tmem_d, q, c_first);
#pragma unroll
            for (int cc = 0; cc < SLABS; cc++) {
                const int c = c_first + cc;        // (SLABS is even: c and cc have the same parity)
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cc + 1 < SLABS) {
                    tmem_load_slab(v[(cc + 1) & 1], tmem_d, q, c + 1);
                } else {                           // last slab is in registers: hand TMEM back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty_bar);
                    TC_STAMP(tr, 5);
                }
                const uint32_t (&w)[4][16] = v[cc & 1];
                const int o0 = n0 + c * EPI_COLS;
                float x[16];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    // float32 round-to-nearest sum of the three K-thirds, then the cross terms, then the bias -- two columns per
                    // instruction (add.rn.f32x2: the epilogue is bound by instruction issue)
                    const uint64_t main = add_f32x2(add_f32x2(pack2(w[1][j], w[1][j + 1]), pack2(w[2][j], w[2][j + 1])), pack2(w[3][j], w[3][j + 1]));
                    const float2 bb = *reinterpret_cast<const float2*>(&bias_s[c * EPI_COLS + j]);
                    // (float16 pairs: the cross-term accumulator carries the 2^11 of lo'; 0x3a000000 = 2^-11, exact)
                    const uint64_t with_cross = kHalf ? fma_f32x2(pack2(w[0][j], w[0][j + 1]), pack2(0x3a000000u, 0x3a000000u), main)
                                                      : add_f32x2(main, pack2(w[0][j], w[0][j + 1]));
                    const uint64_t xx = add_f32x2(with_cross, pack2(__float_as_uint(bb.x), __float_as_uint(bb.y)));
                    uint32_t x0, x1;
                    unpack2(xx, x0, x1);
                    x[j] = __uint_as_float(x0);
                    x[j + 1] = __uint_as_float(x1);
                }
                if constexpr (!kFinal && kHalf) {
                    // float16 pairs: this slab is 32 bytes per row and array; two slabs make the 64-byte rows of the transpose
                    uint32_t hp[8], lp[8];
#pragma unroll
                    for (int j = 0; j < 16; j += 2) split_f16x2(fmaxf(x[j], 0.f), fmaxf(x[j + 1], 0.f), hp[j >> 1], lp[j >> 1]);
                    if ((cc & 1) == 0) __syncwarp();   // the previous pair of slabs has been read out
                    const int fw = (lane >> 1) & 3;
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const int chunk = 2 * (cc & 1) + j;
                        *reinterpret_cast<uint4*>(mine + lane * 64 + ((chunk ^ fw) << 4)) = make_uint4(hp[4 * j], hp[4 * j + 1], hp[4 * j + 2], hp[4 * j + 3]);
                        *reinterpret_cast<uint4*>(mine + SLAB_BYTES + lane * 64 + ((chunk ^ fw) << 4)) =
                            make_uint4(lp[4 * j], lp[4 * j + 1], lp[4 * j + 2], lp[4 * j + 3]);
                    }
                    if (cc & 1) {
                        __syncwarp();
                        __half* oh = reinterpret_cast<__half*>(g.out_hi);
                        __half* ol = reinterpret_cast<__half*>(g.out_lo);
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int row = (lane >> 2) + 8 * i, chunk = lane & 3, phys = chunk ^ ((row >> 1) & 3);
                            const uint4 h4 = *reinterpret_cast<const uint4*>(mine + row * 64 + (phys << 4));
                            const uint4 l4 = *reinterpret_cast<const uint4*>(mine + SLAB_BYTES + row * 64 + (phys << 4));
                            const size_t off = (size_t)(m0 + 32 * q + row) * KDIM + (o0 - EPI_COLS) + 8 * chunk;
                            *reinterpret_cast<uint4*>(oh + off) = h4;
                            *reinterpret_cast<uint4*>(ol + off) = l4;
                        }
                    }
                } else if constexpr (!kFinal) {
                    float hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const float r0 = fmaxf(x[j], 0.f), r1 = fmaxf(x[j + 1], 0.f);
                        hi[j] = rn_tf32(r0);
                        hi[j + 1] = rn_tf32(r1);
                        uint32_t d0, d1;
                        unpack2(sub_f32x2(pack2(__float_as_uint(r0), __float_as_uint(r1)), pack2(__float_as_uint(hi[j]), __float_as_uint(hi[j + 1]))), d0, d1);
                        lo[j] = rn_tf32(__uint_as_float(d0));
                        lo[j + 1] = rn_tf32(__uint_as_float(d1));
                    }
                    // transpose through this warp's slab: lane = row on the way in (16-byte chunks XOR-swizzled by the
                    // row so that neither side has bank conflicts), 4 lanes per row on the way out
                    __syncwarp();                  // the previous slab has been read out
                    const int fw = (lane >> 1) & 3;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        *reinterpret_cast<float4*>(mine + lane * 64 + ((j ^ fw) << 4)) = make_float4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                        *reinterpret_cast<float4*>(mine + SLAB_BYTES + lane * 64 + ((j ^ fw) << 4)) =
                            make_float4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int row = (lane >> 2) + 8 * i, chunk = lane & 3, phys = chunk ^ ((row >> 1) & 3);
                        const float4 h4 = *reinterpret_cast<const float4*>(mine + row * 64 + (phys << 4));
                        const float4 l4 = *reinterpret_cast<const float4*>(mine + SLAB_BYTES + row * 64 + (phys << 4));
                        const size_t off = (size_t)(m0 + 32 * q + row) * KDIM + o0 + 4 * chunk;
                        *reinterpret_cast<float4*>(g.out_hi + off) = h4;
                        *reinterpret_cast<float4*>(g.out_lo + off) = l4;
                    }
                } else if (m < g.m_valid) {  // m is chunk-local, and so is the grid pointer
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int o = o0 + j;
                        if (o < g.n_valid) g.out_hi[(size_t)o * g.out_ld + m] = __fadd_rn(__fmul_rn(x[j], g.span), g.omin);
                    }
                }
            }
            TC_STAMP(tr, 6);
#ifdef PFR_TC_TRACE
            if (tr) { uint32_t sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); tr[7] = sm; }
#endif
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
    }
}

// layer 1 (K <= 4) fused with the input scaling, written K-major as a hi/lo pair (TF32 pairs in float32 containers, or
// float16 pairs).  One warp per condition row: the four scaled inputs (IEEE divisions, as the reference computes them) once
// per lane, then lane l produces 16 outputs in runs of 16 bytes per array so that every store instruction of the warp writes
// 512 contiguous bytes.  Memory-bound on the 4 KB (2 KB) it writes per condition; a warp walks rows with a grid stride.
template <bool kHalf>
__global__ void __launch_bounds__(256)
mlp_tc_layer1_kernel(const float* __restrict__ W1, const float* __restrict__ b1, int in_dim, float lo0, float lo1, float lo2,
                     float lo3, float sp0, float sp1, float sp2, float sp3, float fullL, float fullU,
                     const float* __restrict__ T, const float* __restrict__ P, const float* __restrict__ L,
                     const float* __restrict__ U, int m_valid, int rows, float* __restrict__ Hhi, float* __restrict__ Hlo) {
    // Lane l always produces the same 16 outputs (outputs k0 .. k0 + RUN - 1 of each 32 RUN wide slab), so their fc1 rows and biases live in REGISTERS
    // for the whole kernel (a shared-memory copy indexed by k0 = RUN * lane put the 32 lanes on 4 banks: an 8-way conflict on every
    // weight read, 134 us per 75 776-condition lane for a kernel that only has 155 MB to write).  Missing inputs (in_dim = 2) carry
    // zero weights: fma(x, 0, acc) == acc, so the sum is the one of the FP32 path, i ascending.
    constexpr int RUN = kHalf ? 8 : 4;           // outputs per 16-byte store
    constexpr int NJ = KDIM / (32 * RUN);
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    float w_r[NJ][RUN][4], b_r[NJ][RUN];
#pragma unroll
    for (int j = 0; j < NJ; j++)
#pragma unroll
        for (int t = 0; t < RUN; t++) {
            const int k = 32 * RUN * j + RUN * lane + t;
            b_r[j][t] = b1[k];
#pragma unroll
            for (int i = 0; i < 4; i++) w_r[j][t][i] = i < in_dim ? W1[k * in_dim + i] : 0.f;
        }
    for (int m = blockIdx.x * wpb + (threadIdx.x >> 5); m < rows; m += gridDim.x * wpb) {
        const int ms = m < m_valid ? m : m_valid - 1;
        float x[4];
        x[0] = __fdiv_rn(__fsub_rn(T[ms], lo0), sp0);
        x[1] = __fdiv_rn(__fsub_rn(P[ms], lo1), sp1);
        x[2] = in_dim > 2 ? __fdiv_rn(__fsub_rn(L ? L[ms] : fullL, lo2), sp2) : 0.f;
        x[3] = in_dim > 2 ? __fdiv_rn(__fsub_rn(U ? U[ms] : fullU, lo3), sp3) : 0.f;
#pragma unroll
        for (int j = 0; j < NJ; j++) {
            const int k0 = 32 * RUN * j + RUN * lane;
            float a[RUN];
#pragma unroll
            for (int t = 0; t < RUN; t++) {
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 4; i++) acc = fmaf(x[i], w_r[j][t][i], acc);   // i ascending, like the FP32 path
                a[t] = fmaxf(acc + b_r[j][t], 0.f);
            }
            if constexpr (kHalf) {
                uint32_t hp[4], lp[4];
#pragma unroll
                for (int t = 0; t < 4; t++) split_f16x2(a[2 * t], a[2 * t + 1], hp[t], lp[t]);
                *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(Hhi) + (size_t)m * KDIM + k0) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(Hlo) + (size_t)m * KDIM + k0) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
            } else {
                float hi[4], lo[4];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    hi[t] = rn_tf32(a[t]);
                    lo[t] = rn_tf32(a[t] - hi[t]);
                }
                *reinterpret_cast<float4*>(Hhi + (size_t)m * KDIM + k0) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(Hlo + (size_t)m * KDIM + k0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
    }
}

}  // namespace tc
}  // namespace pfr
