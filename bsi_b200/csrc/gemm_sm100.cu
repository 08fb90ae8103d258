// bf16 x bf16 -> fp32 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA into 128B-swizzled shared memory, persistent CTAs, warp-specialised:
//   warp 0  : TMA producer (one elected lane)
//   warp 1  : MMA issuer   (one elected lane, tcgen05.mma cta_group::1, UMMA 128 x 256 x 16)
//   warp 2  : TMEM allocator
//   warps 4-7: epilogue (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
// The accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps
// the main loop of tile i+1.
//
// C[b][m][n] = epi( sum_k A[b][m][k] * W[b][n][k] ): both operands K-major (row-major activations
// and nn.Linear weights), which is what every Linear of the reference DiT needs
// (bsi/models/dit.py:33-34,71-76,79-81,154,163-165).
//
// Roofline: tensor-bound.  Algorithmic work 2*M*N*K flop per launch.
#include <cuda.h>

#include <vector>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace bsi {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int UMMA_K = 16;
constexpr int kStages = 4;
constexpr int kABytes = BM * BK * 2, kBBytes = BN * BK * 2, kStageBytes = kABytes + kBBytes;
constexpr int kGemmThreads = 256;
constexpr int kTmemCols = 512;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;

struct EpiParams {
    void* C;
    const float* bias;
    int M, N, ldc;
    long long stride_c, stride_bias;
    bsi_rowref gate;
    const int* step_ptr;
    int rows_per_sample;
    const float* pos;
    int patch, grid_w, channels;
};

__device__ __forceinline__ float gelu_tanh(float x) {
    // 0.5*x*(1+tanh(sqrt(2/pi)*(x+0.044715x^3)))  (nn.GELU(approximate="tanh"), bsi/models/dit.py:75)
    float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    return 0.5f * x * (1.0f + t);
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

// Epilogue for one thread's row and a chunk of 32 consecutive columns starting at n0.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& ep, int batch, int row, int n0, const uint32_t (&acc)[32], int step) {
    if (row >= ep.M || n0 >= ep.N) return;
    const float* bias = ep.bias ? ep.bias + (long long)batch * ep.stride_bias : nullptr;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
    const int ncols = min(32, ep.N - n0);  // multiple of 8 by contract (4 for UNPATCH)
    if (bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            if (j < ncols) {
                float4 b4 = *reinterpret_cast<const float4*>(bias + n0 + j);
                v[j] += b4.x, v[j + 1] += b4.y, v[j + 2] += b4.z, v[j + 3] += b4.w;
            }
        }
    }
    if constexpr (EPI == BSI_EPI_BIAS_BF16 || EPI == BSI_EPI_BIAS_GELU_BF16 || EPI == BSI_EPI_BIAS_SILU_BF16) {
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(ep.C) + (long long)batch * ep.stride_c + (long long)row * ep.ldc + n0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            if (j < ncols) {
                float w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float t = v[j + i];
                    if constexpr (EPI == BSI_EPI_BIAS_GELU_BF16) t = gelu_tanh(t);
                    if constexpr (EPI == BSI_EPI_BIAS_SILU_BF16) t = silu(t);
                    w[i] = t;
                }
                uint4 pk = make_uint4(pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]), pack_bf16(w[4], w[5]), pack_bf16(w[6], w[7]));
                *reinterpret_cast<uint4*>(out + j) = pk;
            }
        }
    } else if constexpr (EPI == BSI_EPI_BIAS_F32) {
        float* out = reinterpret_cast<float*>(ep.C) + (long long)batch * ep.stride_c + (long long)row * ep.ldc + n0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            if (j < ncols) *reinterpret_cast<float4*>(out + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else if constexpr (EPI == BSI_EPI_GATE_RESID_F32) {
        // x = addcmul(x, gate, branch)  (bsi/models/dit.py:93-102)
        float* out = reinterpret_cast<float*>(ep.C) + (long long)batch * ep.stride_c + (long long)row * ep.ldc + n0;
        const float* g = rowref_ptr(ep.gate, row / ep.rows_per_sample, step) + n0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            if (j < ncols) {
                float4 x4 = *reinterpret_cast<const float4*>(out + j);
                float4 g4 = *reinterpret_cast<const float4*>(g + j);
                x4.x = fmaf(g4.x, v[j], x4.x), x4.y = fmaf(g4.y, v[j + 1], x4.y);
                x4.z = fmaf(g4.z, v[j + 2], x4.z), x4.w = fmaf(g4.w, v[j + 3], x4.w);
                *reinterpret_cast<float4*>(out + j) = x4;
            }
        }
    } else if constexpr (EPI == BSI_EPI_POS_F32) {
        float* out = reinterpret_cast<float*>(ep.C) + (long long)batch * ep.stride_c + (long long)row * ep.ldc + n0;
        const float* p = ep.pos + (long long)(row % ep.rows_per_sample) * ep.N + n0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            if (j < ncols) {
                float4 p4 = *reinterpret_cast<const float4*>(p + j);
                *reinterpret_cast<float4*>(out + j) = make_float4(v[j] + p4.x, v[j + 1] + p4.y, v[j + 2] + p4.z, v[j + 3] + p4.w);
            }
        }
    } else if constexpr (EPI == BSI_EPI_UNPATCH_F32) {
        // b (nh nw) (ph pw c) -> b c (nh ph) (nw pw)   (bsi/models/dit.py:166-172)
        float* out = reinterpret_cast<float*>(ep.C);
        const int T = ep.rows_per_sample, p = ep.patch, gw = ep.grid_w, ch = ep.channels;
        const int b = row / T, tok = row - b * T;
        const int gy = tok / gw, gx = tok - gy * gw;
        const int Wimg = gw * p, Himg = (T / gw) * p;
        for (int j = 0; j < ncols; ++j) {
            int n = n0 + j;
            int c = n % ch, within = n / ch;
            int py = within / p, px = within - py * p;
            out[(((long long)b * ch + c) * Himg + gy * p + py) * Wimg + gx * p + px] = v[j];
        }
    }
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
    k_gemm_bf16(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const EpiParams ep,
                const int m_tiles, const int n_tiles, const int k_blocks, const int batch, const int a_shared) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = batch * m_tiles * n_tiles;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&tmem_full[b], 1);
            ptx::mbar_init(&tmem_empty[b], 4);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc<1>(tmem_slot, kTmemCols);
        ptx::tmem_relinquish<1>();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n_t = tile % n_tiles, m_t = (tile / n_tiles) % m_tiles, b = tile / (n_tiles * m_tiles);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
                    uint8_t* sa = smem + stage * kStageBytes;
                    ptx::tma_load_3d(sa, &map_a, &full_bar[stage], kb * BK, m_t * BM, a_shared ? 0 : b);
                    ptx::tma_load_3d(sa + kABytes, &map_w, &full_bar[stage], kb * BK, n_t * BN, b);
                    if (++stage == kStages) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                ptx::mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * kStageBytes);
                    const uint64_t da = ptx::umma_desc_k_sw128(sa), db = ptx::umma_desc_k_sw128(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 32 B (16 bf16) inside the 128 B swizzle row: +2 in 16-byte address units
                        ptx::umma_bf16_ss<1>(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit<1>(&empty_bar[stage]);  // frees the smem stage when these MMAs retire
                    if (++stage == kStages) stage = 0, phase ^= 1;
                }
                ptx::umma_commit<1>(&tmem_full[buf]);  // accumulator complete
            }
        }
    } else if (warp >= 4) {
        const int q = warp - 4;  // TMEM lane quarter == warp_id % 4
        const int step = ep.step_ptr ? *ep.step_ptr : 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int n_t = tile % n_tiles, m_t = (tile / n_tiles) % m_tiles, b = tile / (n_tiles * m_tiles);
            const int buf = it & 1;
            ptx::mbar_wait(&tmem_full[buf], (it >> 1) & 1);
            ptx::tc_fence_after();
            const int row = m_t * BM + q * 32 + lane;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t acc[32];
                ptx::tmem_ld_32x32b_x32(taddr + c0, acc);
                ptx::tmem_ld_wait();
                if (c0 + 32 >= BN) {
                    // all TMEM reads of this buffer are done: hand it back to the MMA warp before the last stores
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tmem_empty[buf]);
                }
                epilogue_chunk<EPI>(ep, b, row, n_t * BN + c0, acc, step);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// bf16 [batch][rows][ld] tensor, box {64 (K), box_rows, 1}, 128-byte swizzle, zero fill out of bounds.
int make_operand_map(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld, int64_t batch, int64_t batch_stride,
                     int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return BSI_ERR_CUDA;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 8) != 0 || (batch > 1 && (batch_stride % 8) != 0)) {
        set_error("GEMM operand must be 16-byte aligned with pitch multiple of 8 elements (ld=%lld)", (long long)ld);
        return BSI_ERR_INVALID_ARGUMENT;
    }
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(batch > 1 ? batch : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? batch_stride : rows * ld) * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld K=%lld ld=%lld batch=%lld)", (int)r, (long long)rows,
                  (long long)K, (long long)ld, (long long)batch);
        return BSI_ERR_CUDA;
    }
    return BSI_OK;
}

// Optional per-launch timing of the GEMM kernel (bench.py roofline): CUDA events on the launching stream.
struct GemmRecord {
    cudaEvent_t start, stop;
    double flops;
};
static bool g_profile = false;
static std::vector<GemmRecord> g_records;

template <int EPI>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mw, const EpiParams& ep, int m_tiles, int n_tiles, int k_blocks,
                       int batch, int a_shared, cudaStream_t stream, int k_dim) {
    static bool configured = false;
    if (!configured) {
        BSI_CUDA_OK(cudaFuncSetAttribute(k_gemm_bf16<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        configured = true;
    }
    int total = batch * m_tiles * n_tiles;
    int grid = total < sm_count() ? total : sm_count();
    GemmRecord rec{};
    if (g_profile) {
        BSI_CUDA_OK(cudaEventCreate(&rec.start));
        BSI_CUDA_OK(cudaEventCreate(&rec.stop));
        rec.flops = 2.0 * batch * (double)ep.M * (double)ep.N * (double)k_dim;
        BSI_CUDA_OK(cudaEventRecord(rec.start, stream));
    }
    k_gemm_bf16<EPI><<<grid, kGemmThreads, kSmemBytes, stream>>>(ma, mw, ep, m_tiles, n_tiles, k_blocks, batch, a_shared);
    BSI_LAUNCH_OK("k_gemm_bf16");
    if (g_profile) {
        BSI_CUDA_OK(cudaEventRecord(rec.stop, stream));
        g_records.push_back(rec);
    }
    return BSI_OK;
}

}  // namespace bsi

using namespace bsi;

extern "C" int bsi_gemm_bf16(const bsi_gemm_args* a, void* stream) {
    BSI_CHECK_ARG(a && a->A && a->W && a->C, "bsi_gemm_bf16: null operand");
    BSI_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0 && a->batch >= 1, "bsi_gemm_bf16: bad shape M=%d N=%d K=%d batch=%d", a->M,
                  a->N, a->K, a->batch);
    BSI_CHECK_ARG(a->N % (a->epilogue == BSI_EPI_UNPATCH_F32 ? 4 : 8) == 0, "bsi_gemm_bf16: N=%d must be a multiple of 8 (4 for UNPATCH)", a->N);
    BSI_CHECK_ARG(a->lda >= a->K && a->ldw >= a->K, "bsi_gemm_bf16: pitch smaller than K");
    const bool f32_out = a->epilogue >= BSI_EPI_BIAS_F32;
    if (a->epilogue != BSI_EPI_UNPATCH_F32) {
        BSI_CHECK_ARG(a->ldc >= a->N && a->ldc % (f32_out ? 4 : 8) == 0, "bsi_gemm_bf16: ldc=%d invalid for N=%d", a->ldc, a->N);
        BSI_CHECK_ARG((reinterpret_cast<uintptr_t>(a->C) & 15) == 0, "bsi_gemm_bf16: C must be 16-byte aligned");
    }
    BSI_CHECK_ARG(!a->bias || (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0, "bsi_gemm_bf16: bias must be 16-byte aligned");
    if (a->epilogue == BSI_EPI_GATE_RESID_F32)
        BSI_CHECK_ARG(a->gate.base && a->rows_per_sample > 0, "bsi_gemm_bf16: GATE_RESID needs gate and rows_per_sample");
    if (a->epilogue == BSI_EPI_POS_F32) BSI_CHECK_ARG(a->pos && a->rows_per_sample > 0, "bsi_gemm_bf16: POS needs pos table");
    if (a->epilogue == BSI_EPI_UNPATCH_F32)
        BSI_CHECK_ARG(a->patch > 0 && a->grid_w > 0 && a->channels > 0 && a->rows_per_sample % a->grid_w == 0 &&
                          a->N == a->patch * a->patch * a->channels && a->batch == 1,
                      "bsi_gemm_bf16: UNPATCH geometry invalid");

    CUtensorMap ma, mw;
    const int a_shared = (a->batch > 1 && a->stride_a == 0) ? 1 : 0;
    int rc = make_operand_map(&ma, a->A, a->M, a->K, a->lda, a_shared ? 1 : a->batch, a->stride_a, BM);
    if (rc != BSI_OK) return rc;
    rc = make_operand_map(&mw, a->W, a->N, a->K, a->ldw, a->batch, a->stride_w, BN);
    if (rc != BSI_OK) return rc;

    EpiParams ep;
    ep.C = a->C, ep.bias = a->bias, ep.M = a->M, ep.N = a->N, ep.ldc = a->ldc;
    ep.stride_c = a->stride_c, ep.stride_bias = a->stride_bias;
    ep.gate = a->gate, ep.step_ptr = a->step_ptr, ep.rows_per_sample = a->rows_per_sample > 0 ? a->rows_per_sample : 1;
    ep.pos = a->pos, ep.patch = a->patch, ep.grid_w = a->grid_w, ep.channels = a->channels;

    const int m_tiles = (a->M + BM - 1) / BM, n_tiles = (a->N + BN - 1) / BN, k_blocks = (a->K + BK - 1) / BK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (a->epilogue) {
        case BSI_EPI_BIAS_BF16: return launch_gemm<BSI_EPI_BIAS_BF16>(ma, mw, ep, m_tiles, n_tiles, k_blocks, a->batch, a_shared, st, a->K);
        case BSI_EPI_BIAS_GELU_BF16: return launch_gemm<BSI_EPI_BIAS_GELU_BF16>(ma, mw, ep, m_tiles, n_tiles, k_blocks, a->batch, a_shared, st, a->K);
        case BSI_EPI_BIAS_SILU_BF16: return launch_gemm<BSI_EPI_BIAS_SILU_BF16>(ma, mw, ep, m_tiles, n_tiles, k_blocks, a->batch, a_shared, st, a->K);
        case BSI_EPI_BIAS_F32: return launch_gemm<BSI_EPI_BIAS_F32>(ma, mw, ep, m_tiles, n_tiles, k_blocks, a->batch, a_shared, st, a->K);
        case BSI_EPI_GATE_RESID_F32: return launch_gemm<BSI_EPI_GATE_RESID_F32>(ma, mw, ep, m_tiles, n_tiles, k_blocks, a->batch, a_shared, st, a->K);
        case BSI_EPI_POS_F32: return launch_gemm<BSI_EPI_POS_F32>(ma, mw, ep, m_tiles, n_tiles, k_blocks, a->batch, a_shared, st, a->K);
        case BSI_EPI_UNPATCH_F32: return launch_gemm<BSI_EPI_UNPATCH_F32>(ma, mw, ep, m_tiles, n_tiles, k_blocks, a->batch, a_shared, st, a->K);
        default: set_error("bsi_gemm_bf16: unknown epilogue %d", a->epilogue); return BSI_ERR_INVALID_ARGUMENT;
    }
}

// Start / stop timing every bsi_gemm_bf16 launch with CUDA events (not legal during stream capture).
extern "C" int bsi_profile_gemm_begin(void) {
    g_records.clear();
    g_profile = true;
    return BSI_OK;
}
// Synchronises the recorded events and returns the summed device time [ms], algorithmic flops and launch count.
extern "C" int bsi_profile_gemm_end(double* total_ms, double* total_flops, int32_t* launches) {
    g_profile = false;
    double ms = 0.0, fl = 0.0;
    for (auto& r : g_records) {
        float t = 0.0f;
        BSI_CUDA_OK(cudaEventSynchronize(r.stop));
        BSI_CUDA_OK(cudaEventElapsedTime(&t, r.start, r.stop));
        ms += t, fl += r.flops;
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    if (total_ms) *total_ms = ms;
    if (total_flops) *total_flops = fl;
    if (launches) *launches = (int32_t)g_records.size();
    g_records.clear();
    return BSI_OK;
}
