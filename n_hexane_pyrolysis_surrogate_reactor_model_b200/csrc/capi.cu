// C-ABI entry points (include/crnn_pfr.h): handle management, argument checks, kernel launches.
#include "../../include/crnn_pfr.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <new>
#include <vector>

#include "crnn_device.cuh"
#include "integrate_dopri5.cuh"
#include "integrate_rodas.cuh"
#include "integrate_explicit.cuh"
#include "integrate_rodas_coop.cuh"
#include "integrate_explicit.cuh"
#include "integrate_lanes.cuh"
#include "integrate_taylor.cuh"
#include "adjoint.cuh"
#include "adjoint_phases.cuh"
#include "mlp.cuh"
#include "mlp_tc.cuh"
#include "mlp_train.cuh"
#include "sweep_order.cuh"

using namespace pfr;

static thread_local char g_cuda_err[256] = "";
static std::atomic<unsigned long long> g_launches{0};  // kernels launched by this library (bench.py reports it)

static int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", where, cudaGetErrorString(e));
    return PFR_ECUDA;
}
#define CK(call)                                             \
    do {                                                     \
        cudaError_t e_ = (call);                             \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);  \
    } while (0)
#define CK_LAUNCH(name)                                        \
    do {                                                       \
        cudaError_t e_ = cudaGetLastError();                   \
        if (e_ != cudaSuccess) return cuda_fail(e_, name);     \
        g_launches++;                                          \
    } while (0)

struct crnn_model {
    CrnnParams<double> pd;
    CrnnParams<float> pf;
};

struct pfr_mlp {
    int device;  // the device that holds the weights
    int in_dim;
    int npad4;  // output layer padded to a multiple of GEMM_BM
    float *W1, *b1, *Wt2, *b2, *Wt3, *b3, *Wt4, *b4;
    float span, omin;
    MlpInputScale sc;
    // tensor-core path (mlp_tc.cuh): hi/lo split of fc2..fc4 in nn.Linear layout [out][512] and their TMA maps -- TF32 pairs
    // (float32 containers) and float16 pairs (lo scaled by 2^11); the maps describe whichever pair the mode uses
    int mode;  // PFR_MLP_FP32 | PFR_MLP_TF32X3 | PFR_MLP_F16X3
    float *Whi[3], *Wlo[3];
    __half *Whh[3], *Wlh[3];
    bool f16_ok;  // every fc2..fc4 weight is inside float16's range
    CUtensorMap mapWhi[3], mapWlo[3];
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link dependency)
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn g_encode = nullptr;
static int make_map_2d(CUtensorMap* map, const void* base, uint64_t rows, uint32_t box_rows, bool half) {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) return cuda_fail(cudaErrorNotSupported, "cuTensorMapEncodeTiled entry point");
        g_encode = (encode_tiled_fn)fn;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)tc::KDIM, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)tc::KDIM * (half ? sizeof(__half) : sizeof(float))};
    const cuuint32_t box[2] = {(cuuint32_t)(half ? tc::Operand<true>::BK : tc::Operand<false>::BK), box_rows};   // one 128-byte row
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode(map, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return PFR_ECUDA;
    }
    return PFR_OK;
}

static inline float host_rn_tf32(float x) {  // round to nearest (ties away) on the 13 dropped mantissa bits, like cvt.rna.tf32.f32
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}

// ------------------------------------------------------------------------------------------------
// Per-DEVICE library state.  Everything the library itself keeps on a GPU -- the log / exp tables of fastmath.cuh, the work-queue
// counters of the explicit integrators, the auxiliary streams and fork / join events of the MLP chunk pipeline, the opt-in to
// more than 48 KB of dynamic shared memory (a per-device function attribute) -- lives in the context of the device that is
// current when a call arrives, created on first use under a mutex.  Handles that own device memory (pfr_mlp, pfr_mlp_trainer,
// pfr_sweep) record their device, and a call that arrives with another device current returns PFR_EINVAL.
constexpr int MAX_DEVICES = 64, COUNTER_RING = 64, AUX_STREAMS = 4;
struct DeviceCtx {
    bool ready = false;
    int num_sms = 0;
    FastTables* tables = nullptr;
    int* counters = nullptr;                  // ring of work-queue counters: launches in flight on different streams do not share one
    std::atomic<unsigned> next_counter{0};
    cudaStream_t aux[AUX_STREAMS] = {};       // second lane of the MLP chunk pipeline, round robin over calls
    cudaEvent_t fork[AUX_STREAMS] = {}, join[AUX_STREAMS] = {};
    std::atomic<unsigned> next_aux{0};
    // the device's one block of CRNN coefficients in __constant__ memory (pfr_const_coef, read by the float64 explicit integrators):
    // what it holds, and an event behind the last launch that reads it
    std::mutex coef_mutex;
    ConstCoef coef_host;
    bool coef_valid = false, coef_launched = false;
    cudaEvent_t coef_event = nullptr;
    static constexpr int COEF_SLOTS = 8;      // page-locked staging ring for the uploads (a pageable source would make the copy wait for the stream)
    ConstCoef* coef_pinned = nullptr;
    cudaEvent_t coef_slot_done[COEF_SLOTS] = {};
    unsigned coef_slot = 0;
};
static DeviceCtx g_ctx[MAX_DEVICES];
static std::mutex g_ctx_mutex;

template <typename K>
static int opt_in_smem(K kernel, size_t bytes) {
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return PFR_OK;
}

static int current_device(int* dev) {
    CK(cudaGetDevice(dev));
    if (*dev < 0 || *dev >= MAX_DEVICES) return PFR_EINVAL;
    return PFR_OK;
}

static int device_ctx(DeviceCtx** out) {
    int dev = 0, rc;
    if ((rc = current_device(&dev))) return rc;
    DeviceCtx& c = g_ctx[dev];
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    if (!c.ready) {
        CK(cudaDeviceGetAttribute(&c.num_sms, cudaDevAttrMultiProcessorCount, dev));
        // log / exp tables, computed in long double
        FastTables h;
        for (int i = 0; i < LOGTAB_N; i++) {
            const long double cc = 1.0L + ((long double)i + 0.5L) / LOGTAB_N;
            const double inv = (double)(1.0L / cc);
            h.logtab[i].x = inv;
            h.logtab[i].y = (double)(-logl((long double)inv));
        }
        for (int j = 0; j < EXPTAB_N; j++) h.exptab[j] = (double)powl(2.0L, (long double)j / EXPTAB_N);
        CK(cudaMalloc((void**)&c.tables, sizeof(FastTables)));
        CK(cudaMemcpy(c.tables, &h, sizeof(FastTables), cudaMemcpyHostToDevice));
        CK(cudaMalloc((void**)&c.counters, COUNTER_RING * sizeof(int)));
        for (int i = 0; i < AUX_STREAMS; i++) {
            CK(cudaStreamCreateWithFlags(&c.aux[i], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&c.fork[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c.join[i], cudaEventDisableTiming));
        }
        // kernels that need more than 48 KB of dynamic shared memory
        if ((rc = opt_in_smem(tc::mlp_tc_gemm_kernel<false, false>, tc::SMEM_DYN)) || (rc = opt_in_smem(tc::mlp_tc_gemm_kernel<true, false>, tc::SMEM_DYN)) ||
            (rc = opt_in_smem(tc::mlp_tc_gemm_kernel<false, true>, tc::SMEM_DYN)) || (rc = opt_in_smem(tc::mlp_tc_gemm_kernel<true, true>, tc::SMEM_DYN)) ||
            (rc = opt_in_smem(taylor4_kernel<double, true>, taylor_smem_bytes<double>())) || (rc = opt_in_smem(taylor4_kernel<double, false>, taylor_smem_bytes<double>())) ||
            (rc = opt_in_smem(dp54_kernel<double>, dp54_smem_bytes<double>())) || (rc = opt_in_smem(dp54_kernel<float>, dp54_smem_bytes<float>())) ||
            (rc = opt_in_smem(rodas4_kernel<double, true, true>, (size_t)sm_entries<true>() * RODAS_BLOCK * sizeof(double))) ||
            (rc = opt_in_smem(rodas4_kernel<double, false, true>, (size_t)sm_entries<false>() * RODAS_BLOCK * sizeof(double))) ||
            (rc = opt_in_smem(rodas4_kernel<double, false, false>, (size_t)sm_entries<false>() * RODAS_BLOCK * sizeof(double))) ||
            (rc = opt_in_smem(rodas4_kernel<float, true, true>, (size_t)sm_entries<true>() * RODAS_BLOCK * sizeof(float))) ||
            (rc = opt_in_smem(rodas4_kernel<float, false, true>, (size_t)sm_entries<false>() * RODAS_BLOCK * sizeof(float))) ||
            (rc = opt_in_smem(rodas4_kernel<float, false, false>, (size_t)sm_entries<false>() * RODAS_BLOCK * sizeof(float))))
            return rc;
        c.ready = true;
    }
    *out = &c;
    return PFR_OK;
}

static int check_device(int owner) {
    int dev = 0, rc;
    if ((rc = current_device(&dev))) return rc;
    if (dev != owner) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "handle belongs to device %d, but device %d is current", owner, dev);
        return PFR_EINVAL;
    }
    return PFR_OK;
}

// Conditions per MLP chunk.  The tensor-core path splits a chunk into two lanes of 75 776 rows = 592 row tiles of 128: the
// hidden layers then have 592 x 4 = 2368 = 16 x 148 tiles and the output layer 592 x 7 = 4144 = 28 x 148, i.e. whole waves
// of the 148 persistent CTAs for both GEMM shapes (with 65 536 the output layer ran 12.1 -> 13 waves; measured 15.2 ->
// 14.4 ms per pass at 2^20 conditions).
constexpr int DEFAULT_CHUNK = 2 * 4 * 148 * 128;
static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int eff_chunk(int n, int chunk) {
    if (chunk <= 0) chunk = DEFAULT_CHUNK;
    chunk = round_up(chunk, GEMM_BN);
    const int nn = round_up(n < 1 ? 1 : n, GEMM_BN);
    return chunk < nn ? chunk : nn;
}

extern "C" int pfr_version(void) { return 100; }

extern "C" const char* pfr_status_string(int code) {
    switch (code) {
        case PFR_OK: return "ok";
        case PFR_EINVAL: return "invalid argument";
        case PFR_ECUDA: return "CUDA error";
        case PFR_EWORKSPACE: return "workspace too small";
        default: return "unknown";
    }
}
extern "C" const char* pfr_last_cuda_error(void) { return g_cuda_err; }
extern "C" unsigned long long pfr_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------------
static void crnn_model_fill(crnn_model* m, const float* w_in, const float* w_b, const float* w_out) {
    for (int k = 0; k < NS; k++)
        for (int j = 0; j < NR; j++) {
            m->pd.nu[k][j] = (double)w_in[k * NR + j];
            m->pf.nu[k][j] = w_in[k * NR + j];
            m->pd.wout[k][j] = (double)w_out[k * NR + j];
            m->pf.wout[k][j] = w_out[k * NR + j];
        }
    for (int j = 0; j < NR; j++) {
        m->pd.Ea[j] = (double)w_in[9 * NR + j];
        m->pf.Ea[j] = w_in[9 * NR + j];
        m->pd.b[j] = (double)w_in[10 * NR + j];
        m->pf.b[j] = w_in[10 * NR + j];
        m->pd.lnA[j] = (double)w_b[j];
        m->pf.lnA[j] = w_b[j];
    }
}

extern "C" int crnn_model_create(const float* w_in, const float* w_b, const float* w_out, const double* clamps, crnn_model_t* out) {
    if (!w_in || !w_b || !w_out || !out) return PFR_EINVAL;
    crnn_model* m = new (std::nothrow) crnn_model;
    if (!m) return PFR_EINVAL;
    const double def[6] = {1.0e-6, 6.0e1, -3.0e1, 3.0e1, -1.0e5, 1.0e5};
    const double* c = clamps ? clamps : def;
    // the reference holds R_kcal as a float32 tensor / rounds it to float32 in the multiply
    const float Rk = 1.9872036e-3f;
    crnn_model_fill(m, w_in, w_b, w_out);
    m->pd.lb = c[0]; m->pd.ub = c[1]; m->pd.zlo = c[2]; m->pd.zhi = c[3]; m->pd.dulo = c[4]; m->pd.duhi = c[5];
    m->pf.lb = (float)c[0]; m->pf.ub = (float)c[1]; m->pf.zlo = (float)c[2]; m->pf.zhi = (float)c[3];
    m->pf.dulo = (float)c[4]; m->pf.duhi = (float)c[5];
    m->pd.inv_R = 1.0 / (double)Rk;
    m->pf.inv_R = 1.0f / Rk;
    *out = m;
    return PFR_OK;
}

// The parameters travel to the kernels BY VALUE (a __grid_constant__ argument copied at launch time), so a model handle holds no
// device memory: new parameters take effect with the next launch and launches already enqueued keep the values they were
// given.  The training loop calls this once per optimisation step instead of creating and destroying a handle.
extern "C" int crnn_model_update(crnn_model_t m, const float* w_in, const float* w_b, const float* w_out) {
    if (!m || !w_in || !w_b || !w_out) return PFR_EINVAL;
    crnn_model_fill(m, w_in, w_b, w_out);
    return PFR_OK;
}

extern "C" int crnn_model_destroy(crnn_model_t m) {
    delete m;
    return PFR_OK;
}

// ------------------------------------------------------------------------------------------------
static int upload(const std::vector<float>& h, float** d) {
    CK(cudaMalloc((void**)d, h.size() * sizeof(float)));
    CK(cudaMemcpy(*d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    return PFR_OK;
}

extern "C" int pfr_mlp_create(int in_dim, const float* const weights[4], const float* const biases[4], double out_min,
                   double out_max, const double* in_lo, const double* in_hi, pfr_mlp_t* out) {
    if ((in_dim != 2 && in_dim != 4) || !weights || !biases || !in_lo || !in_hi || !out) return PFR_EINVAL;
    for (int i = 0; i < 4; i++)
        if (!weights[i] || !biases[i]) return PFR_EINVAL;
    pfr_mlp* m = new (std::nothrow) pfr_mlp;
    if (!m) return PFR_EINVAL;
    memset(m, 0, sizeof(*m));
    {
        const int rc = current_device(&m->device);
        if (rc) { delete m; return rc; }
    }
    m->in_dim = in_dim;
    m->f16_ok = true;
    m->npad4 = round_up(MLP_OUT, GEMM_BM);
    // `out * (max - min) + min`: (max - min) in double, both scalars rounded to float32 by the tensor op
    m->span = (float)(out_max - out_min);
    m->omin = (float)out_min;
    for (int k = 0; k < 4; k++) { m->sc.lo[k] = 0.f; m->sc.span[k] = 1.f; }
    for (int k = 0; k < in_dim; k++) { m->sc.lo[k] = (float)in_lo[k]; m->sc.span[k] = (float)(in_hi[k] - in_lo[k]); }
    m->sc.fullL = 1.0f;
    m->sc.fullU = 2.5f;
    int rc;
    std::vector<float> h;
    h.assign(weights[0], weights[0] + MLP_HID * in_dim);
    if ((rc = upload(h, &m->W1))) return rc;
    h.assign(biases[0], biases[0] + MLP_HID);
    if ((rc = upload(h, &m->b1))) return rc;
    // hidden layers: Wt[k][o] = W[o][k]
    float** wt[2] = {&m->Wt2, &m->Wt3};
    float** bb[2] = {&m->b2, &m->b3};
    for (int l = 0; l < 2; l++) {
        h.assign((size_t)MLP_HID * MLP_HID, 0.f);
        for (int o = 0; o < MLP_HID; o++)
            for (int k = 0; k < MLP_HID; k++) h[(size_t)k * MLP_HID + o] = weights[l + 1][(size_t)o * MLP_HID + k];
        if ((rc = upload(h, wt[l]))) return rc;
        h.assign(biases[l + 1], biases[l + 1] + MLP_HID);
        if ((rc = upload(h, bb[l]))) return rc;
    }
    h.assign((size_t)MLP_HID * m->npad4, 0.f);
    for (int o = 0; o < MLP_OUT; o++)
        for (int k = 0; k < MLP_HID; k++) h[(size_t)k * m->npad4 + o] = weights[3][(size_t)o * MLP_HID + k];
    if ((rc = upload(h, &m->Wt4))) return rc;
    h.assign(m->npad4, 0.f);
    for (int o = 0; o < MLP_OUT; o++) h[o] = biases[3][o];
    if ((rc = upload(h, &m->b4))) return rc;
    // tensor-core path: fc2..fc4 as TF32 hi/lo pairs, nn.Linear layout
    for (int l = 0; l < 3; l++) {
        const int rows = l < 2 ? MLP_HID : MLP_OUT;
        std::vector<float> hi((size_t)rows * MLP_HID), lo((size_t)rows * MLP_HID);
        for (size_t e = 0; e < hi.size(); e++) {
            hi[e] = host_rn_tf32(weights[l + 1][e]);
            lo[e] = host_rn_tf32(weights[l + 1][e] - hi[e]);
        }
        if ((rc = upload(hi, &m->Whi[l]))) return rc;
        if ((rc = upload(lo, &m->Wlo[l]))) return rc;
        // float16 pairs: hi = rn_f16(w), lo' = rn_f16((w - hi) * 2^11)  (the difference and the scaling are exact in float32)
        std::vector<__half> hh(hi.size()), lh(hi.size());
        for (size_t e = 0; e < hh.size(); e++) {
            const float w = weights[l + 1][e];
            if (!(fabsf(w) <= 65504.f)) m->f16_ok = false;
            hh[e] = __float2half_rn(w);
            lh[e] = __float2half_rn((w - __half2float(hh[e])) * 2048.f);
        }
        CK(cudaMalloc((void**)&m->Whh[l], hh.size() * sizeof(__half)));
        CK(cudaMemcpy(m->Whh[l], hh.data(), hh.size() * sizeof(__half), cudaMemcpyHostToDevice));
        CK(cudaMalloc((void**)&m->Wlh[l], lh.size() * sizeof(__half)));
        CK(cudaMemcpy(m->Wlh[l], lh.data(), lh.size() * sizeof(__half), cudaMemcpyHostToDevice));
    }
    m->mode = PFR_MLP_FP32;
    *out = m;
    return PFR_OK;
}

extern "C" int pfr_mlp_set_mode(pfr_mlp_t m, int mode) {
    if (!m || (mode != PFR_MLP_FP32 && mode != PFR_MLP_TF32X3 && mode != PFR_MLP_F16X3)) return PFR_EINVAL;
    {
        const int rc = check_device(m->device);
        if (rc) return rc;
    }
    if (mode == PFR_MLP_F16X3 && !m->f16_ok) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "PFR_MLP_F16X3: a weight of fc2..fc4 is outside float16's range; use PFR_MLP_TF32X3");
        return PFR_EINVAL;
    }
    if (mode != PFR_MLP_FP32) {
        const bool half = mode == PFR_MLP_F16X3;
        for (int l = 0; l < 3; l++) {
            const int rows = l < 2 ? MLP_HID : MLP_OUT;
            const uint32_t box = tc::BN;   // the 800-row output layer's last tile reads past the end: TMA zero-fills
            int rc;
            if ((rc = make_map_2d(&m->mapWhi[l], half ? (const void*)m->Whh[l] : (const void*)m->Whi[l], rows, box, half))) return rc;
            if ((rc = make_map_2d(&m->mapWlo[l], half ? (const void*)m->Wlh[l] : (const void*)m->Wlo[l], rows, box, half))) return rc;
        }
    }
    m->mode = mode;
    return PFR_OK;
}

extern "C" int pfr_mlp_destroy(pfr_mlp_t m) {
    if (!m) return PFR_OK;
    float* ptrs[14] = {m->W1, m->b1, m->Wt2, m->b2, m->Wt3, m->b3, m->Wt4, m->b4, m->Whi[0], m->Whi[1], m->Whi[2],
                       m->Wlo[0], m->Wlo[1], m->Wlo[2]};
    for (float* p : ptrs)
        if (p) cudaFree(p);
    for (int l = 0; l < 3; l++) {
        if (m->Whh[l]) cudaFree(m->Whh[l]);
        if (m->Wlh[l]) cudaFree(m->Wlh[l]);
    }
    delete m;
    return PFR_OK;
}

extern "C" size_t pfr_mlp_workspace_bytes(int n, int chunk) {
    const size_t ld = (size_t)eff_chunk(n, chunk);
    // FP32 path: two activation buffers [512][ld]; tensor-core path: two hi/lo pairs [ld][512]; + one output scratch [800][ld]
    return (4 * (size_t)MLP_HID + (size_t)MLP_OUT) * ld * sizeof(float);
}

extern "C" int pfr_inlet_concentration(const float* T, const float* P, int n, float* c0, void* stream) {
    if (n == 0) return PFR_OK;
    if (!T || !P || !c0 || n < 0) return PFR_EINVAL;
    const double mw_hex = 6 * 12.011 + 14 * 1.008, mw_h2o = 2 * 1.008 + 15.999;  // Cantera 3.0 atomic weights
    const double factor = 1.0 / (0.7 * (mw_hex / mw_h2o) + 1);
    inlet_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(T, P, n, (float)8.314462618, factor, c0);
    CK_LAUNCH("inlet_kernel");
    return PFR_OK;
}

// Tensor-core variant of mlp_run's chunk loop (mlp_tc.cuh): activations [ld][512] as TF32 hi/lo pairs, three tcgen05 GEMMs.
// development hook (library built with -DPFR_TC_TRACE): device buffers [tiles][8] receiving the per-tile time stamps of the
// three GEMMs of the next MLP chunk; not declared in include/crnn_pfr.h
static unsigned long long* g_tc_trace[3] = {nullptr, nullptr, nullptr};
extern "C" int pfr_dev_set_tc_trace(unsigned long long* l2, unsigned long long* l3, unsigned long long* l4) {
    g_tc_trace[0] = l2;
    g_tc_trace[1] = l3;
    g_tc_trace[2] = l4;
    return PFR_OK;
}
template <bool kFinal>
static int launch_tc_gemm(const DeviceCtx& ctx, bool half, const CUtensorMap& ahi, const CUtensorMap& alo, const CUtensorMap& bhi, const CUtensorMap& blo,
                          const tc::GemmArgs& g, cudaStream_t st) {
    auto kern = half ? tc::mlp_tc_gemm_kernel<kFinal, true> : tc::mlp_tc_gemm_kernel<kFinal, false>;
    const int total = g.n_tiles * g.m_tiles;
    kern<<<total < ctx.num_sms ? total : ctx.num_sms, tc::THREADS, tc::SMEM_DYN, st>>>(ahi, alo, bhi, blo, g);   // persistent: one CTA per SM
    CK_LAUNCH("mlp_tc_gemm_kernel");
    return PFR_OK;
}

// Chunks alternate between two lanes (each with its own activation buffers, half of the workspace): lane 0 runs on the
// caller's stream, lane 1 on an auxiliary stream forked from / joined to it by events.  The GEMM kernels of the two lanes
// take the SMs in turn (one persistent CTA per SM with ~210 KB of shared memory), while the HBM-bound kernels of one lane
// (first layer, enforce_strict) run beside the other lane's GEMM CTAs instead of in front of them.
// The auxiliary stream and its fork / join events belong to the device context (round robin over calls, so that concurrent
// passes on different caller streams do not queue behind each other on one auxiliary stream); whatever happens inside the
// chunk loop, the auxiliary stream is joined back to the caller's stream before this function returns.
static int mlp_run_tc_chunks(DeviceCtx& ctx, pfr_mlp_t m, const float* T, const float* P, const float* L, const float* U, int n, float* grid,
                             float* t_end, bool is_time, int raw, float* const* Ahi, float* const* Alo, float* const* Bhi, float* const* Blo,
                             float* const* Sl, const CUtensorMap (*mA)[2][2], int ld, int lanes, const cudaStream_t* lane_stream);

static int mlp_run_tc(pfr_mlp_t m, const float* T, const float* P, const float* L, const float* U, int n, float* grid,
                      float* t_end, bool is_time, int raw, float* H, float* S, int ld_full, cudaStream_t st) {
    DeviceCtx* ctxp = nullptr;
    int rc0 = device_ctx(&ctxp);
    if (rc0) return rc0;
    DeviceCtx& ctx = *ctxp;
    const bool two = n > ld_full / 2 && ld_full % (2 * tc::BM) == 0;
    const int ld = two ? ld_full / 2 : ld_full;
    const int lanes = two ? 2 : 1;
    CUtensorMap mA[2][2][2];
    float *Ahi[2], *Alo[2], *Bhi[2], *Blo[2], *Sl[2];
    int rc;
    for (int l = 0; l < lanes; l++) {
        float* Hl = H + (size_t)l * 4 * MLP_HID * ld;
        Ahi[l] = Hl;
        Alo[l] = Hl + (size_t)MLP_HID * ld;
        Bhi[l] = Hl + (size_t)2 * MLP_HID * ld;
        Blo[l] = Hl + (size_t)3 * MLP_HID * ld;
        Sl[l] = S + (size_t)l * MLP_OUT * ld;
        // (float16 pairs use the first half of each array: the workspace layout does not depend on the mode)
        const bool half = m->mode == PFR_MLP_F16X3;
        if ((rc = make_map_2d(&mA[l][0][0], Ahi[l], ld, tc::BM, half)) || (rc = make_map_2d(&mA[l][0][1], Alo[l], ld, tc::BM, half)) ||
            (rc = make_map_2d(&mA[l][1][0], Bhi[l], ld, tc::BM, half)) || (rc = make_map_2d(&mA[l][1][1], Blo[l], ld, tc::BM, half)))
            return rc;
    }
    cudaStream_t lane_stream[2] = {st, st};
    unsigned slot = 0;
    if (two) {
        slot = ctx.next_aux.fetch_add(1) % AUX_STREAMS;
        lane_stream[1] = ctx.aux[slot];
        CK(cudaEventRecord(ctx.fork[slot], st));
        CK(cudaStreamWaitEvent(ctx.aux[slot], ctx.fork[slot], 0));
    }
    rc = mlp_run_tc_chunks(ctx, m, T, P, L, U, n, grid, t_end, is_time, raw, Ahi, Alo, Bhi, Blo, Sl, mA, ld, lanes, lane_stream);
    if (two) {   // join on every path, error or not: the caller's stream must not run ahead of the auxiliary lane
        cudaError_t e1 = cudaEventRecord(ctx.join[slot], lane_stream[1]);
        cudaError_t e2 = cudaStreamWaitEvent(st, ctx.join[slot], 0);
        if (rc == PFR_OK && e1 != cudaSuccess) return cuda_fail(e1, "cudaEventRecord(join)");
        if (rc == PFR_OK && e2 != cudaSuccess) return cuda_fail(e2, "cudaStreamWaitEvent(join)");
    }
    return rc;
}

static int mlp_run_tc_chunks(DeviceCtx& ctx, pfr_mlp_t m, const float* T, const float* P, const float* L, const float* U, int n, float* grid,
                             float* t_end, bool is_time, int raw, float* const* Ahi, float* const* Alo, float* const* Bhi, float* const* Blo,
                             float* const* Sl, const CUtensorMap (*mA)[2][2], int ld, int lanes, const cudaStream_t* lane_stream) {
    int rc;
    const float span = raw ? 1.f : m->span, omin = raw ? 0.f : m->omin;
    int ci = 0;
    for (int c0 = 0; c0 < n; c0 += ld, ci++) {
        const int l = ci % lanes;
        cudaStream_t ls = lane_stream[l];
        const int mv = (n - c0) < ld ? (n - c0) : ld;
        const int rows = round_up(mv, tc::BM);
        const bool half = m->mode == PFR_MLP_F16X3;
        auto layer1 = half ? tc::mlp_tc_layer1_kernel<true> : tc::mlp_tc_layer1_kernel<false>;
        // 8 warps per block; at most two blocks per SM, each warp walking rows with a grid stride: a warp reads its lanes' 80
        // weights / biases into registers once and amortises that over >= 32 rows of a full chunk
        const unsigned l1_blocks = (unsigned)((rows + 63) / 64), l1_cap = 2u * (unsigned)ctx.num_sms;
        layer1<<<l1_blocks < l1_cap ? l1_blocks : l1_cap, 256, 0, ls>>>(
            m->W1, m->b1, m->in_dim, m->sc.lo[0], m->sc.lo[1], m->sc.lo[2], m->sc.lo[3], m->sc.span[0], m->sc.span[1],
            m->sc.span[2], m->sc.span[3], m->sc.fullL, m->sc.fullU, T + c0, P + c0, L ? L + c0 : nullptr, U ? U + c0 : nullptr, mv,
            rows, Ahi[l], Alo[l]);
        CK_LAUNCH("mlp_tc_layer1_kernel");
        const int mt = rows / tc::BM;
        tc::GemmArgs g2{m->b2, Bhi[l], Blo[l], 0, 0, 0, 1.f, 0.f, MLP_HID / tc::BN, mt, g_tc_trace[0]};
        if ((rc = launch_tc_gemm<false>(ctx, half, mA[l][0][0], mA[l][0][1], m->mapWhi[0], m->mapWlo[0], g2, ls))) return rc;
        tc::GemmArgs g3{m->b3, Ahi[l], Alo[l], 0, 0, 0, 1.f, 0.f, MLP_HID / tc::BN, mt, g_tc_trace[1]};
        if ((rc = launch_tc_gemm<false>(ctx, half, mA[l][1][0], mA[l][1][1], m->mapWhi[1], m->mapWlo[1], g3, ls))) return rc;
        float* out_rows = grid ? grid + (size_t)n + c0 : Sl[l];
        const size_t out_ld = grid ? (size_t)n : (size_t)ld;
        tc::GemmArgs g4{m->b4, out_rows, nullptr, out_ld, MLP_OUT, mv, span, omin, (MLP_OUT + tc::BN - 1) / tc::BN, mt, g_tc_trace[2]};
        if ((rc = launch_tc_gemm<true>(ctx, half, mA[l][0][0], mA[l][0][1], m->mapWhi[2], m->mapWlo[2], g4, ls))) return rc;
        if (is_time) {
            if (!raw) {
                enforce_strict_kernel<<<(mv + 255) / 256, 256, 0, ls>>>(out_rows, out_ld, mv, grid ? grid + c0 : nullptr,
                                                                        t_end ? t_end + c0 : nullptr, grid ? 1 : 0);
                CK_LAUNCH("enforce_strict_kernel");
            } else if (grid) {
                CK(cudaMemsetAsync(grid + c0, 0, (size_t)mv * sizeof(float), ls));
            }
        } else {
            copy_row_kernel<<<(mv + 255) / 256, 256, 0, ls>>>(T + c0, grid + c0, mv);
            CK_LAUNCH("copy_row_kernel");
        }
    }
    return PFR_OK;
}

// Shared driver of pfr_time_grid / pfr_temp_profile: chunks of `ld` conditions through the four layers.
static int mlp_run(pfr_mlp_t m, const float* T, const float* P, const float* L, const float* U, int n, float* grid,
                   float* t_end, bool is_time, int raw, void* ws, size_t ws_bytes, int chunk, cudaStream_t st) {
    if (n == 0) return PFR_OK;
    if (!m || !T || !P || n < 0 || (!grid && !t_end) || !ws) return PFR_EINVAL;
    {
        const int rc = check_device(m->device);
        if (rc) return rc;
    }
    const int ld = eff_chunk(n, chunk);
    if (ws_bytes < (4 * (size_t)MLP_HID + (size_t)MLP_OUT) * ld * sizeof(float)) return PFR_EWORKSPACE;
    float* H1 = static_cast<float*>(ws);
    float* H2 = H1 + (size_t)MLP_HID * ld;
    float* S = H1 + (size_t)4 * MLP_HID * ld;  // [800][ld] scratch for the t_end-only mode
    if (m->mode != PFR_MLP_FP32) return mlp_run_tc(m, T, P, L, U, n, grid, t_end, is_time, raw, H1, S, ld, st);
    const float span = raw ? 1.f : m->span, omin = raw ? 0.f : m->omin;
    for (int c0 = 0; c0 < n; c0 += ld) {
        const int mv = (n - c0) < ld ? (n - c0) : ld;
        const int ldc = round_up(mv, GEMM_BN);  // columns actually computed for this chunk
        mlp_layer1_kernel<<<(ldc + 255) / 256, 256, 0, st>>>(m->W1, m->b1, m->in_dim, m->sc, T + c0, P + c0,
                                                             L ? L + c0 : nullptr, U ? U + c0 : nullptr, mv, ld, H1);
        CK_LAUNCH("mlp_layer1_kernel");
        dim3 gh(MLP_HID / GEMM_BM, ldc / GEMM_BN);
        mlp_gemm_kernel<false><<<gh, GEMM_THREADS, 0, st>>>(m->Wt2, m->b2, H1, MLP_HID, MLP_HID, ld, H2, (size_t)ld,
                                                            MLP_HID, ldc, 1.f, 0.f);
        CK_LAUNCH("mlp_gemm_kernel<hidden>");
        mlp_gemm_kernel<false><<<gh, GEMM_THREADS, 0, st>>>(m->Wt3, m->b3, H2, MLP_HID, MLP_HID, ld, H1, (size_t)ld,
                                                            MLP_HID, ldc, 1.f, 0.f);
        CK_LAUNCH("mlp_gemm_kernel<hidden>");
        dim3 go(m->npad4 / GEMM_BM, ldc / GEMM_BN);
        float* rows = grid ? grid + (size_t)n + c0 : S;  // row 1 of the [801][n] grid, or the scratch
        const size_t rows_ld = grid ? (size_t)n : (size_t)ld;
        mlp_gemm_kernel<true><<<go, GEMM_THREADS, 0, st>>>(m->Wt4, m->b4, H1, MLP_HID, m->npad4, ld, rows, rows_ld,
                                                           MLP_OUT, mv, span, omin);
        CK_LAUNCH("mlp_gemm_kernel<final>");
        if (is_time) {
            if (!raw) {
                enforce_strict_kernel<<<(mv + 255) / 256, 256, 0, st>>>(rows, rows_ld, mv, grid ? grid + c0 : nullptr,
                                                                        t_end ? t_end + c0 : nullptr, grid ? 1 : 0);
                CK_LAUNCH("enforce_strict_kernel");
            } else if (grid) {
                CK(cudaMemsetAsync(grid + c0, 0, (size_t)mv * sizeof(float), st));
            }
        } else {
            copy_row_kernel<<<(mv + 255) / 256, 256, 0, st>>>(T + c0, grid + c0, mv);
            CK_LAUNCH("copy_row_kernel");
        }
    }
    return PFR_OK;
}

extern "C" int pfr_time_grid(pfr_mlp_t mlp, const float* T, const float* P, const float* L, const float* u0, int n, float* tgrid,
                  float* t_end, int raw, void* workspace, size_t workspace_bytes, int chunk, void* stream) {
    if (n == 0) return PFR_OK;
    if (!mlp || mlp->in_dim != 4) return PFR_EINVAL;
    if ((L == nullptr) != (u0 == nullptr)) return PFR_EINVAL;
    if (raw && !tgrid) return PFR_EINVAL;
    return mlp_run(mlp, T, P, L, u0, n, tgrid, t_end, true, raw, workspace, workspace_bytes, chunk, (cudaStream_t)stream);
}

extern "C" int pfr_temp_profile(pfr_mlp_t mlp, const float* T, const float* P, int n, float* Tprof, int raw, void* workspace,
                     size_t workspace_bytes, int chunk, void* stream) {
    if (n == 0) return PFR_OK;
    if (!mlp || mlp->in_dim != 2 || !Tprof) return PFR_EINVAL;
    return mlp_run(mlp, T, P, nullptr, nullptr, n, Tprof, nullptr, false, raw, workspace, workspace_bytes, chunk,
                   (cudaStream_t)stream);
}

extern "C" int pfr_idx_cut(const float* t_full, const float* t_end, int n, int* idx, void* stream) {
    if (n == 0) return PFR_OK;
    if (!t_full || !t_end || !idx || n < 0) return PFR_EINVAL;
    idx_cut_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(t_full, (size_t)n, t_end, n, idx);
    CK_LAUNCH("idx_cut_kernel");
    return PFR_OK;
}

extern "C" int pfr_accuracy(const void* dense, int precision, const float* label, const int* idx_end, int n, int abs_den,
                            double* out, void* stream) {
    if (n == 0) return PFR_OK;
    if (!dense || !label || !out || n < 0 || (precision != 32 && precision != 64)) return PFR_EINVAL;
    const int total = n * 7;
    if (precision == 64)
        accuracy_kernel<double><<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const double*)dense, label, idx_end, n, 7, abs_den, out);
    else
        accuracy_kernel<float><<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float*)dense, label, idx_end, n, 7, abs_den, out);
    CK_LAUNCH("accuracy_kernel");
    return PFR_OK;
}

// ------------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(128)
rhs_kernel(const __grid_constant__ CrnnParams<real> p, int n, const real* __restrict__ T, const real* __restrict__ u,
           real* __restrict__ du) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    real kT[NR], dkT[NR], y[NS], f[NS], g[NR], q[NS], md[NS];
    arrhenius_T<real, false>(p, T[i], kT, dkT);
#pragma unroll
    for (int k = 0; k < NS; k++) y[k] = u[(size_t)k * n + i];
    crnn_rhs<real, false>(p, kT, y, f, g, q, md);
#pragma unroll
    for (int k = 0; k < NS; k++) du[(size_t)k * n + i] = f[k];
}

extern "C" int pfr_rhs(crnn_model_t m, int n, const void* T, const void* u, void* du, int precision, void* stream) {
    if (n == 0) return PFR_OK;
    if (!m || !T || !u || !du || n < 0 || (precision != 32 && precision != 64)) return PFR_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == 64)
        rhs_kernel<double><<<(n + 127) / 128, 128, 0, st>>>(m->pd, n, (const double*)T, (const double*)u, (double*)du);
    else
        rhs_kernel<float><<<(n + 127) / 128, 128, 0, st>>>(m->pf, n, (const float*)T, (const float*)u, (float*)du);
    CK_LAUNCH("rhs_kernel");
    return PFR_OK;
}

// ------------------------------------------------------------------------------------------------
template <typename real, bool kRamp, bool kKnots>
static int launch_rodas(const CrnnParams<real>& p, const RodasArgs& a, cudaStream_t st) {
    auto kern = rodas4_kernel<real, kRamp, kKnots>;
    const size_t smem = (size_t)sm_entries<kRamp>() * RODAS_BLOCK * sizeof(real);
    kern<<<(a.n + RODAS_BLOCK - 1) / RODAS_BLOCK, RODAS_BLOCK, smem, st>>>(p, a);   // (shared-memory opt-in: device_ctx)
    CK_LAUNCH("rodas4_kernel");
    return PFR_OK;
}

template <typename real>
static int dispatch_rodas(const CrnnParams<real>& p, const RodasArgs& a, cudaStream_t st) {
    if (a.Tprof) return launch_rodas<real, true, true>(p, a, st);
    // isothermal: knot-limited only when the knots matter (dense output, or an outlet knot other than the last)
    if (a.y_dense || a.idx_end) return launch_rodas<real, false, true>(p, a, st);
    return launch_rodas<real, false, false>(p, a, st);
}

template <typename real, bool kRamp, bool kKnots, int kMethod>
static int launch_rodas_coop(const CrnnParams<real>& p, const RodasArgs& a, cudaStream_t st) {
    const int blocks = (a.n + COOP_PER_BLOCK - 1) / COOP_PER_BLOCK;
    const size_t dyn = PFR_AINV_SMEM ? (size_t)3 * NS * COOP_BLOCK * sizeof(real) : 0;
    rodas4_coop_kernel<real, kRamp, kKnots, kMethod><<<blocks, COOP_BLOCK, dyn, st>>>(p, a);
    CK_LAUNCH("rodas4_coop_kernel");
    return PFR_OK;
}

template <typename real, int kMethod>
static int dispatch_rodas_coop(const CrnnParams<real>& p, const RodasArgs& a, cudaStream_t st) {
    if (a.Tprof) return launch_rodas_coop<real, true, true, kMethod>(p, a, st);
    if (a.y_dense || a.idx_end) return launch_rodas_coop<real, false, true, kMethod>(p, a, st);
    return launch_rodas_coop<real, false, false, kMethod>(p, a, st);
}

// bs23_kernel / dp54_kernel draw their work items from a zero-initialised device counter: one of a small per-device ring, so
// that launches in flight on different streams do not share one
static int next_counter(DeviceCtx& ctx, cudaStream_t st, int** out) {
    int* c = ctx.counters + ctx.next_counter.fetch_add(1) % COUNTER_RING;
    CK(cudaMemsetAsync(c, 0, sizeof(int), st));
    *out = c;
    return PFR_OK;
}

// The float64 explicit integrators read their two 9 x 9 coefficient matrices from ONE __constant__ block per device (fixed
// addresses are what lets ptxas use 16-byte uniform loads, integrate_explicit.cuh).  Launches that use the block are chained
// through an event -- each of them fills the GPU with a persistent grid, so ordering them across streams costs nothing -- and the
// block is re-uploaded, in stream order behind the previous launch, only when the coefficients differ from what it holds.
static int stage_const_coef(DeviceCtx& ctx, const CrnnParams<double>& p, cudaStream_t st) {
#if PFR_CONST_COEF
    ConstCoef h;
    fill_const_coef(h, p);
    std::lock_guard<std::mutex> lock(ctx.coef_mutex);
    if (!ctx.coef_event) {
        CK(cudaEventCreateWithFlags(&ctx.coef_event, cudaEventDisableTiming));
        CK(cudaMallocHost((void**)&ctx.coef_pinned, DeviceCtx::COEF_SLOTS * sizeof(ConstCoef)));
        for (int i = 0; i < DeviceCtx::COEF_SLOTS; i++) CK(cudaEventCreateWithFlags(&ctx.coef_slot_done[i], cudaEventDisableTiming));
    }
    if (ctx.coef_launched) CK(cudaStreamWaitEvent(st, ctx.coef_event, 0));
    if (!ctx.coef_valid || memcmp(&h, &ctx.coef_host, sizeof(h)) != 0) {
        const unsigned slot = ctx.coef_slot++ % DeviceCtx::COEF_SLOTS;
        CK(cudaEventSynchronize(ctx.coef_slot_done[slot]));   // (its previous upload, eight uploads ago: long done)
        ctx.coef_pinned[slot] = h;
        CK(cudaMemcpyToSymbolAsync(pfr_const_coef, &ctx.coef_pinned[slot], sizeof(h), 0, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(ctx.coef_slot_done[slot], st));
        ctx.coef_host = h;
        ctx.coef_valid = true;
    }
#endif
    return PFR_OK;
}
static int stage_const_coef(DeviceCtx&, const CrnnParams<float>&, cudaStream_t) { return PFR_OK; }
static int const_coef_launched(DeviceCtx& ctx, const CrnnParams<double>&, cudaStream_t st) {
#if PFR_CONST_COEF
    std::lock_guard<std::mutex> lock(ctx.coef_mutex);
    CK(cudaEventRecord(ctx.coef_event, st));
    ctx.coef_launched = true;
#endif
    return PFR_OK;
}
static int const_coef_launched(DeviceCtx&, const CrnnParams<float>&, cudaStream_t) { return PFR_OK; }

template <typename real>
static int dispatch_bs23(DeviceCtx& ctx, const CrnnParams<real>& p, const RodasArgs& a0, cudaStream_t st) {
    RodasArgs a = a0;
    int rc = next_counter(ctx, st, &a.work_counter);
    if (rc) return rc;
    const int full = (a.n + BS23_BLOCK - 1) / BS23_BLOCK, persistent = BS23_CTAS_PER_SM * ctx.num_sms;
    const int grid = full < persistent ? full : persistent;
    CoefDup<real> cd;
    fill_coef_dup(cd, p);
    if ((rc = stage_const_coef(ctx, p, st))) return rc;
    if (a.Tprof) bs23_kernel<real, true><<<grid, BS23_BLOCK, bs23_smem_bytes<real>(), st>>>(p, cd, a);
    else bs23_kernel<real, false><<<grid, BS23_BLOCK, bs23_smem_bytes<real>(), st>>>(p, cd, a);
    CK_LAUNCH("bs23_kernel");
    return const_coef_launched(ctx, p, st);
}

template <typename real>
static int dispatch_taylor4(DeviceCtx& ctx, const CrnnParams<real>& p, const RodasArgs& a0, cudaStream_t st) {
    RodasArgs a = a0;
    int rc = next_counter(ctx, st, &a.work_counter);
    if (rc) return rc;
    const int full = (a.n + TAYLOR_BLOCK - 1) / TAYLOR_BLOCK, persistent = TAYLOR_CTAS_PER_SM * ctx.num_sms;
    const int grid = full < persistent ? full : persistent;
    if (a.Tprof) taylor4_kernel<real, true><<<grid, TAYLOR_BLOCK, taylor_smem_bytes<real>(), st>>>(p, a);
    else taylor4_kernel<real, false><<<grid, TAYLOR_BLOCK, taylor_smem_bytes<real>(), st>>>(p, a);
    CK_LAUNCH("taylor4_kernel");
    return PFR_OK;
}

template <typename real>
static int dispatch_dp54(DeviceCtx& ctx, const CrnnParams<real>& p, const RodasArgs& a0, cudaStream_t st) {
    RodasArgs a = a0;
    int rc = next_counter(ctx, st, &a.work_counter);
    if (rc) return rc;
    const int full = (a.n + DP54_BLOCK - 1) / DP54_BLOCK, persistent = DP54_CTAS_PER_SM * ctx.num_sms;
    CoefDup<real> cd;
    fill_coef_dup(cd, p);
    if ((rc = stage_const_coef(ctx, p, st))) return rc;
    dp54_kernel<real><<<full < persistent ? full : persistent, DP54_BLOCK, dp54_smem_bytes<real>(), st>>>(p, cd, a);
    CK_LAUNCH("dp54_kernel");
    return const_coef_launched(ctx, p, st);
}

template <typename real>
static int dispatch_dopri5(const CrnnParams<real>& p, const Dopri5Args& a, cudaStream_t st) {
    const int grid = (a.n + DOPRI_BLOCK - 1) / DOPRI_BLOCK;
    if (a.Tprof) dopri5_kernel<real, true><<<grid, DOPRI_BLOCK, 0, st>>>(p, a);
    else dopri5_kernel<real, false><<<grid, DOPRI_BLOCK, 0, st>>>(p, a);
    CK_LAUNCH("dopri5_kernel");
    return PFR_OK;
}

// One integrator launch for fully prepared arguments (pfr_integrate and the sweep pipeline share it)
static int integrate_launch(DeviceCtx& ctx, crnn_model_t m, int method, int precision, const RodasArgs& a, cudaStream_t st) {
    const bool d = precision == 64;
    switch (method) {
        case PFR_METHOD_RODAS4: return d ? dispatch_rodas_coop<double, COOP_RODAS4>(m->pd, a, st) : dispatch_rodas_coop<float, COOP_RODAS4>(m->pf, a, st);
        case PFR_METHOD_ROS3: return d ? dispatch_rodas_coop<double, COOP_ROS3>(m->pd, a, st) : dispatch_rodas_coop<float, COOP_ROS3>(m->pf, a, st);
        case PFR_METHOD_DP54: return d ? dispatch_dp54<double>(ctx, m->pd, a, st) : dispatch_dp54<float>(ctx, m->pf, a, st);
        case PFR_METHOD_BS23: return d ? dispatch_bs23<double>(ctx, m->pd, a, st) : dispatch_bs23<float>(ctx, m->pf, a, st);
        case PFR_METHOD_TAYLOR4: return d ? dispatch_taylor4<double>(ctx, m->pd, a, st) : dispatch_taylor4<float>(ctx, m->pf, a, st);
        case PFR_METHOD_BS23_WARP: {
            if (!d) return PFR_EINVAL;
            const int blocks = (a.n + LANES_WARPS - 1) / LANES_WARPS;
            if (a.Tprof) bs23_lanes_kernel<true><<<blocks, 32 * LANES_WARPS, 0, st>>>(m->pd, a);
            else bs23_lanes_kernel<false><<<blocks, 32 * LANES_WARPS, 0, st>>>(m->pd, a);
            CK_LAUNCH("bs23_lanes_kernel");
            return PFR_OK;
        }
        case PFR_METHOD_DP54_WARP: {
            if (!d || a.Tprof) return PFR_EINVAL;
            dp54_lanes_kernel<<<(a.n + LANES_WARPS - 1) / LANES_WARPS, 32 * LANES_WARPS, 0, st>>>(m->pd, a);
            CK_LAUNCH("dp54_lanes_kernel");
            return PFR_OK;
        }
        case PFR_METHOD_RODAS4_TPC: return d ? dispatch_rodas<double>(m->pd, a, st) : dispatch_rodas<float>(m->pf, a, st);
        case PFR_METHOD_DOPRI5: {
            Dopri5Args da{a.n, a.T0, a.c0, a.tgrid, a.Tprof, a.t_end, a.idx_end, a.perm, a.rtol, a.atol, a.y_out, a.y_dense, a.status, a.stats, a.max_steps};
            return d ? dispatch_dopri5<double>(m->pd, da, st) : dispatch_dopri5<float>(m->pf, da, st);
        }
    }
    return PFR_EINVAL;
}

extern "C" int pfr_integrate(crnn_model_t m, int method, int precision, int n, const float* T0, const float* c0,
                  const float* tgrid, const float* Tprof, const float* t_end, const int* idx_end, const int* perm,
                  double rtol, double atol, int max_steps, int flags, void* y_out, void* y_dense, int* status, int* stats, void* stream) {
    if (n == 0) return PFR_OK;
    if (!m || !T0 || !c0 || !y_out || !status || n < 0) return PFR_EINVAL;
    if (precision != 32 && precision != 64) return PFR_EINVAL;
    if (method != PFR_METHOD_RODAS4 && method != PFR_METHOD_DOPRI5 && method != PFR_METHOD_RODAS4_TPC && method != PFR_METHOD_ROS3 &&
        method != PFR_METHOD_BS23 && method != PFR_METHOD_DP54 && method != PFR_METHOD_BS23_WARP && method != PFR_METHOD_TAYLOR4 &&
        method != PFR_METHOD_DP54_WARP)
        return PFR_EINVAL;
    if (method == PFR_METHOD_BS23_WARP && (!tgrid || precision != 64)) return PFR_EINVAL;   // knot-limited, float64 state
    if (method == PFR_METHOD_DP54_WARP && (!tgrid || Tprof || precision != 64)) return PFR_EINVAL;   // isothermal, outputs at tgrid's knots, float64
    if (method == PFR_METHOD_DP54 && (tgrid || !t_end || Tprof || y_dense || idx_end)) return PFR_EINVAL;   // isothermal outlet at t_end only
    if ((method == PFR_METHOD_BS23 || method == PFR_METHOD_TAYLOR4) && !tgrid) return PFR_EINVAL;   // the explicit fast paths are knot-limited steppers
    if (!tgrid && (!t_end || Tprof || y_dense || idx_end)) return PFR_EINVAL;
    if (!(rtol > 0) || !(atol > 0)) return PFR_EINVAL;
    if (n == 0) return PFR_OK;
    if (max_steps <= 0) max_steps = 1000000;
    DeviceCtx* ctx = nullptr;
    {
        const int rc = device_ctx(&ctx);
        if (rc != PFR_OK) return rc;
    }
    RodasArgs a{n, T0, c0, tgrid, Tprof, t_end, idx_end, perm, rtol, atol, y_out, y_dense, status, stats, max_steps, ctx->tables, flags};
    return integrate_launch(*ctx, m, method, precision, a, (cudaStream_t)stream);
}


// Second half of an explicit-fast-path integration, for callers that drive the stages themselves (Surrogate.integrate): the
// conditions that pfr_integrate(PFR_METHOD_BS23 | PFR_METHOD_DP54) left with status PFR_ST_STIFF are collected into a list ON THE
// DEVICE and integrated again, from the inlet, with the Rosenbrock kernel, which overwrites their y_out / y_dense / status / stats.
// Nothing waits for the host; with no flagged condition the second launch finds an empty list and returns at once.
// scratch: n + 1 ints of device memory; scratch[n] receives the number of conditions handed over.
extern "C" int pfr_stiff_fallback(crnn_model_t m, int method, int precision, int n, const float* T0, const float* c0, const float* tgrid,
                                  const float* Tprof, const float* t_end, const int* idx_end, double rtol, double atol, int max_steps,
                                  int flags, void* y_out, void* y_dense, int* status, int* stats, int* scratch, void* stream) {
    if (n == 0) return PFR_OK;
    if (!m || !T0 || !c0 || !y_out || !status || !scratch || n < 0) return PFR_EINVAL;
    if (method != PFR_METHOD_ROS3 && method != PFR_METHOD_RODAS4) return PFR_EINVAL;
    if (precision != 32 && precision != 64) return PFR_EINVAL;
    if (!tgrid && (!t_end || Tprof || y_dense || idx_end)) return PFR_EINVAL;
    if (max_steps <= 0) max_steps = 1000000;
    DeviceCtx* ctx = nullptr;
    int rc = device_ctx(&ctx);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(scratch + n, 0, sizeof(int), st));
    collect_status_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(status, nullptr, n, PFR_ST_STIFF, scratch, scratch + n);
    CK_LAUNCH("collect_status_kernel");
    RodasArgs a{n, T0, c0, tgrid, Tprof, t_end, idx_end, scratch, rtol, atol, y_out, y_dense, status, stats, max_steps, ctx->tables, flags};
    a.n_work = scratch + n;
    return integrate_launch(*ctx, m, method, precision, a, st);
}

// ------------------------------------------------------------------------------------------------
// The whole hot path of one model variant as ONE call (main() sweep loops of SURROGATE_MODEL/surrogate_model_Eoff_single_model.py:
// 339-369 and ...Eon_single_model.py:296-368): visiting order -> inlet concentration -> MLP passes (the independent passes of the
// coupled path on two side streams) -> idx_cut -> integrator -> Rosenbrock fallback for conditions the explicit fast path
// flagged stiff (list and count stay on the device: nothing here waits for the host) -> results in the caller's order.
struct pfr_sweep {
    int device, n_max;
    crnn_model_t crnn;
    pfr_mlp_t time_mlp, temp_mlp;
    float *Ts, *Ps, *Ls, *Us, *c0, *t_end, *t_full, *Tprof;
    int *order, *idx, *hist, *cursor, *stiff_list, *stiff_count;
    void* ws[3];
    size_t ws_bytes;
    cudaStream_t side[2];
    cudaEvent_t ready, done[2];
    // integrator timing: a ring of event pairs, one per run; pfr_sweep_integrator_ms averages the runs since its last call
    static constexpr int TIMING_RING = 16;
    cudaEvent_t k0[TIMING_RING], k1[TIMING_RING];
    unsigned long long runs, runs_reported;
};

extern "C" int pfr_sweep_destroy(pfr_sweep_t s) {
    if (!s) return PFR_OK;
    void* ptrs[] = {s->Ts, s->Ps, s->Ls, s->Us, s->c0, s->t_end, s->t_full, s->Tprof, s->order, s->idx, s->hist, s->cursor,
                    s->stiff_list, s->stiff_count, s->ws[0], s->ws[1], s->ws[2]};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (int i = 0; i < 2; i++) {
        if (s->side[i]) cudaStreamDestroy(s->side[i]);
        if (s->done[i]) cudaEventDestroy(s->done[i]);
    }
    if (s->ready) cudaEventDestroy(s->ready);
    for (int i = 0; i < pfr_sweep::TIMING_RING; i++) {
        if (s->k0[i]) cudaEventDestroy(s->k0[i]);
        if (s->k1[i]) cudaEventDestroy(s->k1[i]);
    }
    delete s;
    return PFR_OK;
}

extern "C" int pfr_sweep_create(crnn_model_t crnn, pfr_mlp_t time_mlp, pfr_mlp_t temp_mlp, int n_max, pfr_sweep_t* out) {
    if (!crnn || !time_mlp || time_mlp->in_dim != 4 || (temp_mlp && temp_mlp->in_dim != 2) || n_max < 1 || !out) return PFR_EINVAL;
    pfr_sweep* s = new (std::nothrow) pfr_sweep();
    if (!s) return PFR_EINVAL;
    memset(s, 0, sizeof(*s));
    int rc = current_device(&s->device);
    if (rc == PFR_OK) rc = check_device(time_mlp->device);
    if (rc == PFR_OK && temp_mlp) rc = check_device(temp_mlp->device);
    if (rc) { delete s; return rc; }
    s->n_max = n_max;
    s->crnn = crnn;
    s->time_mlp = time_mlp;
    s->temp_mlp = temp_mlp;
    const size_t n = (size_t)n_max;
    s->ws_bytes = pfr_mlp_workspace_bytes(n_max, 0);
    const int passes = temp_mlp ? 3 : 1;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    alloc((void**)&s->Ts, n * 4); alloc((void**)&s->Ps, n * 4); alloc((void**)&s->Ls, n * 4); alloc((void**)&s->Us, n * 4);
    alloc((void**)&s->c0, n * 4); alloc((void**)&s->t_end, n * 4);
    alloc((void**)&s->order, n * 4); alloc((void**)&s->idx, n * 4); alloc((void**)&s->stiff_list, n * 4);
    alloc((void**)&s->hist, ORDER_BINS * 4); alloc((void**)&s->cursor, ORDER_BINS * 4); alloc((void**)&s->stiff_count, 4);
    if (temp_mlp) { alloc((void**)&s->t_full, n * NTOT * 4); alloc((void**)&s->Tprof, n * NTOT * 4); }
    for (int i = 0; i < passes; i++) alloc(&s->ws[i], s->ws_bytes);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaStreamCreateWithFlags(&s->side[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->done[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ready, cudaEventDisableTiming);
    for (int i = 0; i < pfr_sweep::TIMING_RING && e == cudaSuccess; i++) {
        e = cudaEventCreate(&s->k0[i]);
        if (e == cudaSuccess) e = cudaEventCreate(&s->k1[i]);
    }
    if (e == cudaSuccess) e = cudaMemset(s->stiff_count, 0, sizeof(int));
    if (e != cudaSuccess) {
        pfr_sweep_destroy(s);
        return cuda_fail(e, "pfr_sweep_create");
    }
    *out = s;
    return PFR_OK;
}

extern "C" size_t pfr_sweep_device_bytes(int n_max, int energy_on) {
    const size_t n = (size_t)(n_max < 1 ? 1 : n_max);
    return n * 4 * 9 + (energy_on ? 2 * n * NTOT * 4 : 0) + (energy_on ? 3 : 1) * pfr_mlp_workspace_bytes(n_max, 0) + 2 * ORDER_BINS * 4 + 4;
}

extern "C" int pfr_sweep_run(pfr_sweep_t s, const float* T, const float* P, const float* L, const float* u0, int n, int method, int precision,
                             double rtol, double atol, int max_steps, int flags, void* y_out, int* status, int* stats, int* idx_cut_out,
                             float* t_end_out, int* stiff_count_out, void* stream) {
    if (n == 0) return PFR_OK;
    if (!s || !T || !P || !y_out || !status || n < 0 || n > s->n_max) return PFR_EINVAL;
    if ((L == nullptr) != (u0 == nullptr)) return PFR_EINVAL;
    if (precision != 32 && precision != 64) return PFR_EINVAL;
    if (!(rtol > 0) || !(atol > 0)) return PFR_EINVAL;
    const bool eon = s->temp_mlp != nullptr;
    if (eon ? (method != PFR_METHOD_BS23 && method != PFR_METHOD_TAYLOR4 && method != PFR_METHOD_ROS3 && method != PFR_METHOD_RODAS4)
            : (method != PFR_METHOD_DP54 && method != PFR_METHOD_ROS3 && method != PFR_METHOD_RODAS4))
        return PFR_EINVAL;
    if (!eon && !L) return PFR_EINVAL;   // the isothermal sweep integrates to the end of the (T, P, L, u0) grid
    int rc = check_device(s->device);
    if (rc) return rc;
    DeviceCtx* ctx = nullptr;
    if ((rc = device_ctx(&ctx))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (max_steps <= 0) max_steps = 1000000;
    const unsigned blocks = (unsigned)((n + 255) / 256);

    // 1. visiting order (cost proxy known before the MLPs run) and the inputs gathered into it
    const MlpInputScale& sc = s->time_mlp->sc;
    const bool ordered = !(flags & PFR_SWEEP_NO_ORDER) && (eon ? L != nullptr : true);
    const float *Tv = T, *Pv = P, *Lv = L, *Uv = u0;
    const int* order = nullptr;
    if (ordered) {
        OrderKey key;
        if (eon) {   // residence-time ratio L / u0
            const float lo = sc.lo[2] / (sc.lo[3] + sc.span[3]), hi = (sc.lo[2] + sc.span[2]) / sc.lo[3];
            key = {L, u0, lo, (float)ORDER_BINS / (hi - lo)};
        } else {     // inlet temperature
            key = {T, nullptr, sc.lo[0], (float)ORDER_BINS / sc.span[0]};
        }
        CK(cudaMemsetAsync(s->hist, 0, ORDER_BINS * sizeof(int), st));
        order_hist_kernel<<<blocks < 1184u ? blocks : 1184u, 256, 0, st>>>(key, n, s->hist);
        CK_LAUNCH("order_hist_kernel");
        order_scan_kernel<<<1, ORDER_BINS / 2, 0, st>>>(s->hist, s->cursor);
        CK_LAUNCH("order_scan_kernel");
        order_scatter_kernel<<<blocks, 256, 0, st>>>(key, n, s->cursor, s->order, T, P, L, u0, s->Ts, s->Ps, s->Ls, s->Us);
        CK_LAUNCH("order_scatter_kernel");
        Tv = s->Ts; Pv = s->Ps; Lv = L ? s->Ls : nullptr; Uv = u0 ? s->Us : nullptr;
        order = s->order;
    }
    // 2. inlet concentration
    if ((rc = pfr_inlet_concentration(Tv, Pv, n, s->c0, st))) return rc;

    RodasArgs a{n, Tv, s->c0, nullptr, nullptr, nullptr, nullptr, nullptr, rtol, atol, y_out, nullptr, status, stats, max_steps, ctx->tables, 0};
    a.out_index = order;
    if (eon) {
        // 3. the three MLP passes are independent: temperature profile and outlet time on the side streams, the full-length grid
        //    on the caller's stream (GEMM CTAs fill the SMs one kernel at a time; the HBM-bound kernels between them overlap)
        CK(cudaEventRecord(s->ready, st));
        CK(cudaStreamWaitEvent(s->side[0], s->ready, 0));
        rc = pfr_temp_profile(s->temp_mlp, Tv, Pv, n, s->Tprof, 0, s->ws[1], s->ws_bytes, 0, s->side[0]);
        cudaError_t e0 = cudaEventRecord(s->done[0], s->side[0]);
        int rc2 = PFR_OK;
        cudaError_t e1 = cudaSuccess;
        if (L) {
            CK(cudaStreamWaitEvent(s->side[1], s->ready, 0));
            rc2 = pfr_time_grid(s->time_mlp, Tv, Pv, Lv, Uv, n, nullptr, s->t_end, 0, s->ws[2], s->ws_bytes, 0, s->side[1]);
            e1 = cudaEventRecord(s->done[1], s->side[1]);
        }
        int rc3 = pfr_time_grid(s->time_mlp, Tv, Pv, nullptr, nullptr, n, s->t_full, L ? nullptr : s->t_end, 0, s->ws[0], s->ws_bytes, 0, st);
        // join on every path before reporting an error
        cudaError_t e2 = cudaStreamWaitEvent(st, s->done[0], 0);
        cudaError_t e3 = L ? cudaStreamWaitEvent(st, s->done[1], 0) : cudaSuccess;
        if (rc || rc2 || rc3) return rc ? rc : (rc2 ? rc2 : rc3);
        for (cudaError_t e : {e0, e1, e2, e3})
            if (e != cudaSuccess) return cuda_fail(e, "pfr_sweep_run: stream join");
        // 4. outlet knot
        if (L) {
            idx_cut_kernel<<<blocks, 256, 0, st>>>(s->t_full, (size_t)n, s->t_end, n, s->idx);
            CK_LAUNCH("idx_cut_kernel");
        } else {
            fill_int_kernel<<<blocks, 256, 0, st>>>(s->idx, n, NTOT - 1);
            CK_LAUNCH("fill_int_kernel");
        }
        a.tgrid = s->t_full;
        a.Tprof = s->Tprof;
        a.idx_end = s->idx;
    } else {
        if ((rc = pfr_time_grid(s->time_mlp, Tv, Pv, Lv, Uv, n, nullptr, s->t_end, 0, s->ws[0], s->ws_bytes, 0, st))) return rc;
        a.t_end = s->t_end;
    }
    // 5. the integrator (timed by a pair of events that belong to the handle)
    const int slot = (int)(s->runs % pfr_sweep::TIMING_RING);
    CK(cudaEventRecord(s->k0[slot], st));
    if ((rc = integrate_launch(*ctx, s->crnn, method, precision, a, st))) return rc;
    CK(cudaEventRecord(s->k1[slot], st));
    s->runs++;
    // 6. conditions the explicit fast path flagged stiff go through the Rosenbrock kernel; the list is built and counted on the device
    if ((method == PFR_METHOD_BS23 || method == PFR_METHOD_TAYLOR4 || method == PFR_METHOD_DP54) && !(flags & PFR_SWEEP_NO_FALLBACK)) {
        int* count = stiff_count_out ? stiff_count_out : s->stiff_count;
        CK(cudaMemsetAsync(count, 0, sizeof(int), st));
        collect_status_kernel<<<blocks, 256, 0, st>>>(status, order, n, PFR_ST_STIFF, s->stiff_list, count);
        CK_LAUNCH("collect_status_kernel");
        RodasArgs f = a;
        f.perm = s->stiff_list;
        f.n_work = count;
        if ((rc = integrate_launch(*ctx, s->crnn, eon ? PFR_METHOD_ROS3 : PFR_METHOD_RODAS4, precision, f, st))) return rc;
    }
    else if (stiff_count_out) CK(cudaMemsetAsync(stiff_count_out, 0, sizeof(int), st));
    // 7. by-products in the caller's order
    if (idx_cut_out && eon) {
        unorder_kernel<int><<<blocks, 256, 0, st>>>(s->idx, order, n, idx_cut_out);
        CK_LAUNCH("unorder_kernel");
    }
    if (t_end_out) {
        unorder_kernel<float><<<blocks, 256, 0, st>>>(s->t_end, order, n, t_end_out);
        CK_LAUNCH("unorder_kernel");
    }
    return PFR_OK;
}

extern "C" int pfr_sweep_stiff_count(pfr_sweep_t s, int* count) {
    if (!s || !count) return PFR_EINVAL;
    CK(cudaMemcpy(count, s->stiff_count, sizeof(int), cudaMemcpyDeviceToHost));   // synchronises
    return PFR_OK;
}

extern "C" int pfr_sweep_integrator_ms(pfr_sweep_t s, float* ms) {
    if (!s || !ms || s->runs == 0) return PFR_EINVAL;
    unsigned long long first = s->runs_reported;
    if (first == s->runs) first = s->runs - 1;                                        // nothing new: the last run again
    if (s->runs - first > pfr_sweep::TIMING_RING) first = s->runs - pfr_sweep::TIMING_RING;
    double sum = 0;
    for (unsigned long long r = first; r < s->runs; r++) {
        const int slot = (int)(r % pfr_sweep::TIMING_RING);
        float one = 0;
        CK(cudaEventSynchronize(s->k1[slot]));
        CK(cudaEventElapsedTime(&one, s->k0[slot], s->k1[slot]));
        sum += one;
    }
    *ms = (float)(sum / (double)(s->runs - first));
    s->runs_reported = s->runs;
    return PFR_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" int pfr_loss_grad(crnn_model_t m, int n, const float* T0, const float* tgrid, const float* Tprof,
                             const double* y_knots, const float* ref, const float* yscale, int substeps, double* loss,
                             double* grad, void* stream) {
    if (n == 0) return PFR_OK;
    if (!m || !T0 || !tgrid || !y_knots || !ref || !yscale || !loss || !grad || n < 0 || substeps == 0) return PFR_EINVAL;
    DeviceCtx* ctx = nullptr;
    {
        const int rc = device_ctx(&ctx);
        if (rc != PFR_OK) return rc;
    }
    AdjointArgs a{n, T0, tgrid, Tprof, y_knots, ref, yscale, substeps, loss, grad, ctx->tables};
    if (substeps > 0) {
        // one condition per warp
        const int blocks = (n + ADJW_WARPS - 1) / ADJW_WARPS;
        if (Tprof) adjoint_warp_kernel<true><<<blocks, 32 * ADJW_WARPS, 0, (cudaStream_t)stream>>>(m->pd, a);
        else adjoint_warp_kernel<false><<<blocks, 32 * ADJW_WARPS, 0, (cudaStream_t)stream>>>(m->pd, a);
    } else {
        // negative sub-step count: the one-thread-per-condition kernel (cross-check)
        a.substeps = -substeps;
        const int blocks = (n + ADJ_BLOCK - 1) / ADJ_BLOCK;
        if (Tprof) adjoint_kernel<true><<<blocks, ADJ_BLOCK, 0, (cudaStream_t)stream>>>(m->pd, a);
        else adjoint_kernel<false><<<blocks, ADJ_BLOCK, 0, (cudaStream_t)stream>>>(m->pd, a);
    }
    CK_LAUNCH("adjoint_kernel");
    return PFR_OK;
}

// The same loss and gradient in three kernels (adjoint_phases.cuh): node quantities for all conditions x intervals at once, the
// sequential adjoint walk reduced to one 9 x 9 mat-vec per stage, the parameter-gradient quadrature in parallel.
extern "C" size_t pfr_loss_grad_workspace_bytes(int n, int substeps) {
    if (n < 1 || substeps < 1) return 0;
    return ((size_t)adj_nodes_per_condition(substeps) * ADJ_NF + (size_t)adj_stages_per_condition(substeps) * ADJ_SREC) * (size_t)n * sizeof(double);
}

extern "C" int pfr_loss_grad_staged(crnn_model_t m, int n, const float* T0, const float* tgrid, const float* Tprof, const double* y_knots,
                                    const float* ref, const float* yscale, int substeps, double* loss, double* grad, void* workspace,
                                    size_t workspace_bytes, void* stream) {
    if (n == 0) return PFR_OK;
    if (!m || !T0 || !tgrid || !y_knots || !ref || !yscale || !loss || !grad || !workspace || n < 0 || substeps < 1) return PFR_EINVAL;
    if (workspace_bytes < pfr_loss_grad_workspace_bytes(n, substeps)) return PFR_EWORKSPACE;
    DeviceCtx* ctx = nullptr;
    {
        const int rc = device_ctx(&ctx);
        if (rc != PFR_OK) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    AdjPhaseArgs g;
    g.a = AdjointArgs{n, T0, tgrid, Tprof, y_knots, ref, yscale, substeps, loss, grad, ctx->tables};
    g.nodes = static_cast<double*>(workspace);
    g.stages = g.nodes + (size_t)adj_nodes_per_condition(substeps) * ADJ_NF * (size_t)n;
    const dim3 grid1((unsigned)((n + ADJP_BLOCK - 1) / ADJP_BLOCK), (unsigned)(NTOT - 1));
    if (Tprof) adjoint_nodes_kernel<true><<<grid1, ADJP_BLOCK, 0, st>>>(m->pd, g);
    else adjoint_nodes_kernel<false><<<grid1, ADJP_BLOCK, 0, st>>>(m->pd, g);
    CK_LAUNCH("adjoint_nodes_kernel");
    adjoint_sweep_kernel<<<(unsigned)((n + ADJS_WARPS - 1) / ADJS_WARPS), 32 * ADJS_WARPS, 0, st>>>(m->pd, g);
    CK_LAUNCH("adjoint_sweep_kernel");
    switch (substeps) {   // sub-steps per interval as a compile-time constant where it is one of the usual values
        case 1: adjoint_grad_kernel<1><<<(unsigned)n, ADJG_THREADS, 0, st>>>(m->pd, g); break;
        case 2: adjoint_grad_kernel<2><<<(unsigned)n, ADJG_THREADS, 0, st>>>(m->pd, g); break;
        case 3: adjoint_grad_kernel<3><<<(unsigned)n, ADJG_THREADS, 0, st>>>(m->pd, g); break;
        default: adjoint_grad_kernel<0><<<(unsigned)n, ADJG_THREADS, 0, st>>>(m->pd, g);
    }
    CK_LAUNCH("adjoint_grad_kernel");
    return PFR_OK;
}

extern "C" int pfr_reduce_rows(const double* x, int rows, int n, double* out, void* stream) {
    if (rows == 0) return PFR_OK;
    if (!x || !out || rows < 0 || n < 0) return PFR_EINVAL;
    reduce_rows_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(x, n, rows, nullptr, out);
    CK_LAUNCH("reduce_rows_kernel");
    return PFR_OK;
}

extern "C" int pfr_reduce_rows_ok(const double* x, int rows, int n, const int* status, double* out, void* stream) {
    if (!x || !out || !status || rows < 0 || n < 0) return PFR_EINVAL;
    reduce_rows_kernel<<<rows + 1, 256, 0, (cudaStream_t)stream>>>(x, n, rows, status, out);
    CK_LAUNCH("reduce_rows_kernel");
    return PFR_OK;
}

// ------------------------------------------------------------------------------------------------
// Predictor-MLP training step (mlp_train.cuh)
struct pfr_mlp_trainer {
    int device;
    int in_dim;
    long long step;
    float *W[4], *b[4], *mW[4], *vW[4], *mb[4], *vb[4];
    float *act[4], *actT[4];   // ReLU outputs of fc1..fc3 and the network output, batch-major [32][N] and feature-major [N][32]
    float *grad[2];            // ping-pong gradient buffers, feature-major [800][32]
};
static inline int tr_K(const pfr_mlp_trainer* t, int l) { return l == 0 ? t->in_dim : mt::HID; }
static inline int tr_N(int l) { return l == 3 ? mt::OUT : mt::HID; }

extern "C" int pfr_mlp_trainer_create(int in_dim, const float* const weights[4], const float* const biases[4], pfr_mlp_trainer_t* out) {
    if (!out || !weights || !biases || (in_dim != 2 && in_dim != 4)) return PFR_EINVAL;
    pfr_mlp_trainer* t = new (std::nothrow) pfr_mlp_trainer();
    if (!t) return PFR_ECUDA;
    {
        const int rc = current_device(&t->device);
        if (rc) { delete t; return rc; }
    }
    t->in_dim = in_dim;
    t->step = 0;
    for (int l = 0; l < 4; l++) {
        const size_t nw = (size_t)tr_N(l) * tr_K(t, l), nb = (size_t)tr_N(l);
        CK(cudaMalloc((void**)&t->W[l], nw * 4)); CK(cudaMalloc((void**)&t->mW[l], nw * 4)); CK(cudaMalloc((void**)&t->vW[l], nw * 4));
        CK(cudaMalloc((void**)&t->b[l], nb * 4)); CK(cudaMalloc((void**)&t->mb[l], nb * 4)); CK(cudaMalloc((void**)&t->vb[l], nb * 4));
        CK(cudaMemcpy(t->W[l], weights[l], nw * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(t->b[l], biases[l], nb * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(t->mW[l], 0, nw * 4)); CK(cudaMemset(t->vW[l], 0, nw * 4));
        CK(cudaMemset(t->mb[l], 0, nb * 4)); CK(cudaMemset(t->vb[l], 0, nb * 4));
        CK(cudaMalloc((void**)&t->act[l], (size_t)mt::MAXB * tr_N(l) * 4));
        CK(cudaMalloc((void**)&t->actT[l], (size_t)mt::MAXB * tr_N(l) * 4));
    }
    for (int g = 0; g < 2; g++) CK(cudaMalloc((void**)&t->grad[g], (size_t)mt::MAXB * mt::OUT * 4));
    *out = t;
    return PFR_OK;
}

extern "C" int pfr_mlp_trainer_destroy(pfr_mlp_trainer_t t) {
    if (!t) return PFR_OK;
    for (int l = 0; l < 4; l++) {
        cudaFree(t->W[l]); cudaFree(t->mW[l]); cudaFree(t->vW[l]); cudaFree(t->b[l]); cudaFree(t->mb[l]); cudaFree(t->vb[l]);
        cudaFree(t->act[l]); cudaFree(t->actT[l]);
    }
    cudaFree(t->grad[0]); cudaFree(t->grad[1]);
    delete t;
    return PFR_OK;
}

// fc1..fc4 on at most 32 rows; activations stay in the trainer's buffers (both layouts), the network output in act[3]
static int trainer_forward32(pfr_mlp_trainer_t t, const float* x, int B, cudaStream_t st) {
    {
        const int rc = check_device(t->device);
        if (rc) return rc;
    }
    for (int l = 0; l < 4; l++) {
        const int K = tr_K(t, l), N = tr_N(l);
        const float* in = l == 0 ? x : t->actT[l - 1];
        const int sk = l == 0 ? 1 : mt::MAXB, sb = l == 0 ? K : 1;   // x is [B][in_dim]; hidden activations feature-major
        if (l < 3) mt::linear_fwd_kernel<true><<<(N + 7) / 8, 256, 0, st>>>(in, sk, sb, t->W[l], t->b[l], B, K, N, t->act[l], t->actT[l]);
        else mt::linear_fwd_kernel<false><<<(N + 7) / 8, 256, 0, st>>>(in, sk, sb, t->W[l], t->b[l], B, K, N, t->act[l], t->actT[l]);
        CK_LAUNCH("linear_fwd_kernel");
    }
    return PFR_OK;
}

extern "C" int pfr_mlp_trainer_step(pfr_mlp_trainer_t t, const float* x, const float* y, int B, double lr, double beta1, double beta2,
                                    double eps, float* loss, void* stream) {
    if (!t || !x || !y || !loss || B < 1 || B > mt::MAXB || !(lr > 0)) return PFR_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = trainer_forward32(t, x, B, st);
    if (rc) return rc;
    mt::mse_kernel<<<1, 1024, 0, st>>>(t->act[3], y, B, mt::OUT, t->grad[0], loss);
    CK_LAUNCH("mse_kernel");
    t->step++;
    const double bc1 = 1.0 - pow(beta1, (double)t->step), bc2 = 1.0 - pow(beta2, (double)t->step);
    const mt::AdamArgs a{(float)beta1, (float)beta2, (float)eps, (float)(lr / bc1), (float)sqrt(bc2)};
    int cur = 0;   // grad[cur] = gradient w.r.t. the output of layer l
    for (int l = 3; l >= 0; l--) {
        const int K = tr_K(t, l), N = tr_N(l);
        const float* in = l == 0 ? x : t->act[l - 1];
        if (l > 0) {   // gradient w.r.t. this layer's input, through the ReLU that produced it, with the weights of BEFORE the update
            mt::linear_bwd_data_kernel<<<(K + 7) / 8, 256, 0, st>>>(t->grad[cur], t->W[l], t->actT[l - 1], K, N, t->grad[cur ^ 1]);
            CK_LAUNCH("linear_bwd_data_kernel");
        }
        mt::linear_bwd_weight_adam_kernel<<<dim3((K + 255) / 256, (N + mt::WN - 1) / mt::WN), 256, 0, st>>>(
            t->grad[cur], in, B, K, N, t->W[l], t->mW[l], t->vW[l], t->b[l], t->mb[l], t->vb[l], a);
        CK_LAUNCH("linear_bwd_weight_adam_kernel");
        cur ^= 1;
    }
    return PFR_OK;
}

extern "C" int pfr_mlp_trainer_forward(pfr_mlp_trainer_t t, const float* x, int n, float* out, void* stream) {
    if (n == 0) return PFR_OK;
    if (!t || !x || !out || n < 0) return PFR_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    for (int r0 = 0; r0 < n; r0 += mt::MAXB) {
        const int B = n - r0 < mt::MAXB ? n - r0 : mt::MAXB;
        const int rc = trainer_forward32(t, x + (size_t)r0 * t->in_dim, B, st);
        if (rc) return rc;
        CK(cudaMemcpyAsync(out + (size_t)r0 * mt::OUT, t->act[3], (size_t)B * mt::OUT * 4, cudaMemcpyDeviceToDevice, st));
    }
    return PFR_OK;
}

extern "C" int pfr_mlp_trainer_loss(pfr_mlp_trainer_t t, const float* x, const float* y, int B, float* loss, void* stream) {
    if (!t || !x || !y || !loss || B < 1 || B > mt::MAXB) return PFR_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = trainer_forward32(t, x, B, st);
    if (rc) return rc;
    mt::mse_kernel<<<1, 1024, 0, st>>>(t->act[3], y, B, mt::OUT, nullptr, loss);
    CK_LAUNCH("mse_kernel");
    return PFR_OK;
}

extern "C" int pfr_mlp_trainer_read(pfr_mlp_trainer_t t, float* const weights[4], float* const biases[4]) {
    if (!t || !weights || !biases) return PFR_EINVAL;
    CK(cudaDeviceSynchronize());
    for (int l = 0; l < 4; l++) {
        CK(cudaMemcpy(weights[l], t->W[l], (size_t)tr_N(l) * tr_K(t, l) * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(biases[l], t->b[l], (size_t)tr_N(l) * 4, cudaMemcpyDeviceToHost));
    }
    return PFR_OK;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) fastmath_kernel(const FastTables* __restrict__ ft, int kind, int n,
                                                       const double* __restrict__ x, double* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        y[i] = kind == 0   ? fast_log(x[i], ft->logtab)
               : kind == 1 ? fast_exp(x[i], ft->exptab)
               : kind == 2 ? fast_exp_scaled(x[i] * EXP_ARG_SCALE, ft->exptab)   // (the explicit integrators' form: argument in units of ln2 / 256)
               : kind == 3 ? fast_log_ilp(x[i], ft->logtab)
                           : fast_exp_ilp(x[i], ft->exptab);
}

extern "C" int pfr_fastmath(int kind, int n, const double* x, double* y, void* stream) {
    if (n == 0) return PFR_OK;
    if (!x || !y || n < 0 || kind < 0 || kind > 4) return PFR_EINVAL;
    DeviceCtx* ctx = nullptr;
    const int rc = device_ctx(&ctx);
    if (rc != PFR_OK) return rc;
    fastmath_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ctx->tables, kind, n, x, y);
    CK_LAUNCH("fastmath_kernel");
    return PFR_OK;
}

// ------------------------------------------------------------------------------------------------
// pipe micro-benchmarks: dependent-chain-free unrolled loops, enough warps to saturate the pipe
template <int MODE>
__global__ void __launch_bounds__(256) peak_kernel(float* sink, int iters, long long* clk) {
    const long long c0 = clock64();
    if (MODE == 0) {
        float a[8], b = 1.0001f + threadIdx.x * 1e-7f, c = 0.5f;
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = i + threadIdx.x;
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], b, c);
        float s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s += a[i];
        if (s == 12345.678f) sink[0] = s;
    } else if (MODE == 1) {
        double a[8], b = 1.0001 + threadIdx.x * 1e-9, c = 0.5;
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = i + threadIdx.x;
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, c);
        double s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s += a[i];
        if (s == 12345.678) sink[0] = (float)s;
    } else {
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = 0.001f * (i + threadIdx.x);
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        float s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s += a[i];
        if (s == 12345.678f) sink[0] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) clk[MODE] = clock64() - c0;
}

extern "C" int pfr_measure_peaks(double out[4]) {
    if (!out) return PFR_EINVAL;
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float* sink;
    long long* clk;
    CK(cudaMalloc((void**)&sink, 16));
    CK(cudaMalloc((void**)&clk, 3 * sizeof(long long)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256;
    const int iters[3] = {8192, 2048, 2048};
    double clock_hz = 0;
    for (int mode = 0; mode < 3; mode++) {
        double best = 0;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(e0));
            if (mode == 0) peak_kernel<0><<<blocks, threads>>>(sink, iters[0], clk);
            else if (mode == 1) peak_kernel<1><<<blocks, threads>>>(sink, iters[1], clk);
            else peak_kernel<2><<<blocks, threads>>>(sink, iters[2], clk);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK_LAUNCH("peak_kernel");
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double ops = (double)blocks * threads * iters[mode] * 8.0;
            const double rate = ops / (ms * 1e-3) * (mode == 2 ? 1.0 : 2.0);  // FMA = 2 flop
            if (rep > 0 && rate > best) best = rate;
            if (mode == 0 && rep == 3) {
                long long hc[3];
                CK(cudaMemcpy(hc, clk, sizeof(hc), cudaMemcpyDeviceToHost));
                // one block's cycle count over the kernel's wall time (8 blocks/SM run concurrently)
                clock_hz = (double)hc[0] / (ms * 1e-3);
            }
        }
        out[mode] = best;
    }
    out[3] = clock_hz;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(clk);
    return PFR_OK;
}

