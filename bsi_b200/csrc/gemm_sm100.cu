// bf16 x bf16 -> fp32 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA into 128B-swizzled shared memory, persistent CTAs, warp-specialised:
//   warp 0  : TMA producer (one lane)
//   warp 1  : MMA issuer   (one lane of the pair's leader CTA)
//   warp 2  : TMEM allocator
//   warps 4-7: epilogue (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
// Two variants of the same kernel (template parameter CG):
//   CG = 2  CTA pair (cluster of 2, cta_group::2): UMMA 256 x 256 x 16, each CTA holds 128 rows of A and half of
//           the W tile, so every operand byte fetched from L2 feeds twice the math of the single-CTA tile.  The
//           single-CTA kernel measured 0.98 PFLOP/s — exactly the L2->SM bandwidth bound of a 128x256 tile
//           (94 B/clk/SM needed at full MMA rate vs ~40 delivered); the pair needs 64 B/clk/SM.
//   CG = 1  single CTA, UMMA 128 x 256 x 16 (small problems, cross-check in the tests).
// The accumulator is double-buffered in TMEM (2 x 256 columns): the epilogue of tile i overlaps the main loop of
// tile i+1.
//
// C[b][m][n] = epi( sum_k A[b][m][k] * W[b][n][k] ): both operands K-major (row-major activations and nn.Linear
// weights) — every Linear of the reference DiT (bsi/models/dit.py:33-34,71-76,79-81,154,163-165).
// Roofline: tensor-bound, 2*M*N*K flop per launch.
#include <cuda.h>

#include <vector>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace bsi {

constexpr int BM = 128;  // rows of A per CTA
constexpr int BN = 256, BK = 64, UMMA_K = 16;
constexpr int kGemmThreads = 256;
constexpr int kTmemCols = 512;
constexpr int kABytes = BM * BK * 2;

template <int CG>
struct Cfg {
    static constexpr int kBRows = BN / CG;           // rows of the W tile this CTA loads
    static constexpr int kBBytes = kBRows * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = CG == 1 ? 4 : 6;  // 192 KB of operand staging either way
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

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

// Epilogue for one thread's row and a chunk of 32 consecutive columns starting at n0 (all modes but GATE_RESID).
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& ep, int batch, int row, int n0, const uint32_t (&acc)[32]) {
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

// Whole-tile epilogue of one warp: 32 rows (one per lane) x BN columns of TMEM buffer `taddr`.
// `release()` must be called once, after the last tcgen05.ld of the buffer has completed.
template <int EPI, class Release>
__device__ __forceinline__ void epilogue_tile(const EpiParams& ep, int batch, int row, int n_base, uint32_t taddr, int step, Release release) {
    if constexpr (EPI == BSI_EPI_GATE_RESID_F32) {
        // x = addcmul(x, gate, branch)  (bsi/models/dit.py:93-102): fp32 read-modify-write of the residual stream.
        // The residual row is prefetched one 64-column group ahead of the TMEM reads (16 independent 16-byte loads
        // per thread in flight) — with loads issued one at a time behind the stores this epilogue ran at 17 TFLOP/s.
        const bool valid = row < ep.M;
        float* out = reinterpret_cast<float*>(ep.C) + (long long)batch * ep.stride_c + (long long)row * ep.ldc + n_base;
        const float* gate = rowref_ptr(ep.gate, row / ep.rows_per_sample, step) + n_base;
        const float* bias = ep.bias ? ep.bias + (long long)batch * ep.stride_bias + n_base : nullptr;
        float4 res[2][16];
        auto load_group = [&](int g, float4(&r)[16]) {
            if (valid && n_base + g * 64 < ep.N) {
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = *reinterpret_cast<const float4*>(out + g * 64 + 4 * j);
            }
        };
        load_group(0, res[0]);
#pragma unroll
        for (int g = 0; g < BN / 64; ++g) {
            if (g + 1 < BN / 64) load_group(g + 1, res[(g + 1) & 1]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c0 = g * 64 + h * 32;
                uint32_t acc[32];
                ptx::tmem_ld_32x32b_x32(taddr + c0, acc);
                const bool live = valid && n_base + c0 < ep.N;
                float4 g4[8], b4[8];
                if (live) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        g4[j] = *reinterpret_cast<const float4*>(gate + c0 + 4 * j);
                        b4[j] = bias ? *reinterpret_cast<const float4*>(bias + c0 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                ptx::tmem_ld_wait();
                if (g == BN / 64 - 1 && h == 1) release();
                if (live) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 x4 = res[g & 1][h * 8 + j];
                        x4.x = fmaf(g4[j].x, __uint_as_float(acc[4 * j]) + b4[j].x, x4.x);
                        x4.y = fmaf(g4[j].y, __uint_as_float(acc[4 * j + 1]) + b4[j].y, x4.y);
                        x4.z = fmaf(g4[j].z, __uint_as_float(acc[4 * j + 2]) + b4[j].z, x4.z);
                        x4.w = fmaf(g4[j].w, __uint_as_float(acc[4 * j + 3]) + b4[j].w, x4.w);
                        *reinterpret_cast<float4*>(out + c0 + 4 * j) = x4;
                    }
                }
            }
        }
    } else {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t acc[32];
            ptx::tmem_ld_32x32b_x32(taddr + c0, acc);
            ptx::tmem_ld_wait();
            if (c0 + 32 >= BN) release();
            epilogue_chunk<EPI>(ep, batch, row, n_base + c0, acc);
        }
    }
}

template <int EPI, int CG>
__global__ void __launch_bounds__(kGemmThreads, 1)
    k_gemm_bf16(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const EpiParams ep,
                const int m_tiles, const int n_tiles, const int k_blocks, const int batch, const int a_shared) {
    using C = Cfg<CG>;
    constexpr int kStages = C::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * C::kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = batch * m_tiles * n_tiles;
    const int cta_rank = CG == 2 ? (int)ptx::cluster_ctarank() : 0;  // rank inside the CTA pair; 0 issues the MMAs
    const int worker = CG == 2 ? blockIdx.x / 2 : blockIdx.x;          // persistent worker = CTA or CTA pair
    const int num_workers = CG == 2 ? gridDim.x / 2 : gridDim.x;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);   // leader's arrive.expect_tx; TMA bytes of both CTAs land on the leader's barrier
            ptx::mbar_init(&empty_bar[s], 1);  // tcgen05.commit (multicast to both CTAs of a pair)
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&tmem_full[b], 1);        // tcgen05.commit after the last k-block
            ptx::mbar_init(&tmem_empty[b], 4 * CG);  // one arrive per epilogue warp of every CTA of the pair
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc<CG>(tmem_slot, kTmemCols);
        ptx::tmem_relinquish<CG>();
    }
    ptx::tc_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync_all();  // peer barriers must be initialised before remote arrives / TMA signals
    else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer: this CTA's 128 rows of A and its 256/CG rows of W per k-block
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = worker; tile < total_tiles; tile += num_workers) {
                const int n_t = tile % n_tiles, m_t = (tile / n_tiles) % m_tiles, b = tile / (n_tiles * m_tiles);
                const int row_a = (m_t * CG + cta_rank) * BM, row_w = n_t * BN + cta_rank * C::kBRows;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * C::kStageBytes;
                    if constexpr (CG == 1) {
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
                        ptx::tma_load_3d(sa, &map_a, &full_bar[stage], kb * BK, row_a, a_shared ? 0 : b);
                        ptx::tma_load_3d(sa + kABytes, &map_w, &full_bar[stage], kb * BK, row_w, b);
                    } else {
                        if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
                        ptx::tma_load_3d_2sm(sa, &map_a, &full_bar[stage], kb * BK, row_a, a_shared ? 0 : b);
                        ptx::tma_load_3d_2sm(sa + kABytes, &map_w, &full_bar[stage], kb * BK, row_w, b);
                    }
                    if (++stage == kStages) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && cta_rank == 0) {
            // ---------------- MMA issuer (leader CTA): D[tmem] += A[smem] * W[smem]^T, UMMA (128*CG) x 256 x 16
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM * CG, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = worker; tile < total_tiles; tile += num_workers, ++it) {
                const int buf = it & 1;
                ptx::mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * C::kStageBytes);
                    const uint64_t da = ptx::umma_desc_k_sw128(sa), db = ptx::umma_desc_k_sw128(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 32 B (16 bf16) inside the 128 B swizzle row: +2 in 16-byte address units
                        ptx::umma_bf16_ss<CG>(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    // free the smem stage (in both CTAs) once these MMAs have retired
                    if constexpr (CG == 1) ptx::umma_commit<1>(&empty_bar[stage]);
                    else ptx::umma_commit_mcast2(&empty_bar[stage], 0x3);
                    if (++stage == kStages) stage = 0, phase ^= 1;
                }
                if constexpr (CG == 1) ptx::umma_commit<1>(&tmem_full[buf]);  // accumulator complete
                else ptx::umma_commit_mcast2(&tmem_full[buf], 0x3);
            }
        }
    } else if (warp >= 4) {
        // ---------------- epilogue: warp q owns TMEM lanes [32q, 32q+32) = rows of this CTA's half of the tile
        const int q = warp - 4;
        const int step = ep.step_ptr ? *ep.step_ptr : 0;
        int it = 0;
        for (int tile = worker; tile < total_tiles; tile += num_workers, ++it) {
            const int n_t = tile % n_tiles, m_t = (tile / n_tiles) % m_tiles, b = tile / (n_tiles * m_tiles);
            const int buf = it & 1;
            ptx::mbar_wait(&tmem_full[buf], (it >> 1) & 1);
            ptx::tc_fence_after();
            const int row = (m_t * CG + cta_rank) * BM + q * 32 + lane;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;
            epilogue_tile<EPI>(ep, b, row, n_t * BN, taddr, step, [&]() {
                // all TMEM reads of this buffer are done: hand it back to the MMA warp before the remaining stores
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 1 || cta_rank == 0) ptx::mbar_arrive(&tmem_empty[buf]);
                    else ptx::mbar_arrive_cluster(&tmem_empty[buf], 0);
                }
            });
        }
    }

    ptx::tc_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync_all();
    else __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<CG>(tmem_base, kTmemCols);
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
static int g_force_cta_group = 0;  // 0 = automatic, 1 / 2 = forced (tests)

template <int EPI, int CG>
static int launch_gemm(const bsi_gemm_args* a, const EpiParams& ep, int a_shared, cudaStream_t stream) {
    using C = Cfg<CG>;
    static bool configured = false;
    if (!configured) {
        BSI_CUDA_OK(cudaFuncSetAttribute(k_gemm_bf16<EPI, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
        configured = true;
    }
    CUtensorMap ma, mw;
    int rc = make_operand_map(&ma, a->A, a->M, a->K, a->lda, a_shared ? 1 : a->batch, a->stride_a, BM);
    if (rc != BSI_OK) return rc;
    rc = make_operand_map(&mw, a->W, a->N, a->K, a->ldw, a->batch, a->stride_w, C::kBRows);
    if (rc != BSI_OK) return rc;
    const int m_tiles = (a->M + BM * CG - 1) / (BM * CG), n_tiles = (a->N + BN - 1) / BN, k_blocks = (a->K + BK - 1) / BK;
    const int total = a->batch * m_tiles * n_tiles;
    const int max_workers = sm_count() / CG;
    const int workers = total < max_workers ? total : max_workers;

    GemmRecord rec{};
    if (g_profile) {
        BSI_CUDA_OK(cudaEventCreate(&rec.start));
        BSI_CUDA_OK(cudaEventCreate(&rec.stop));
        rec.flops = 2.0 * a->batch * (double)a->M * (double)a->N * (double)a->K;
        BSI_CUDA_OK(cudaEventRecord(rec.start, stream));
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(workers * CG), cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    BSI_CUDA_OK(cudaLaunchKernelEx(&cfg, k_gemm_bf16<EPI, CG>, ma, mw, ep, m_tiles, n_tiles, k_blocks, (int)a->batch, a_shared));
    BSI_LAUNCH_OK("k_gemm_bf16");
    if (g_profile) {
        BSI_CUDA_OK(cudaEventRecord(rec.stop, stream));
        g_records.push_back(rec);
    }
    return BSI_OK;
}

template <int EPI>
static int dispatch_cta_group(const bsi_gemm_args* a, const EpiParams& ep, int a_shared, cudaStream_t stream) {
    // CTA pairs pay off as soon as there are at least two 128-row blocks; tiny problems keep the finer 128-row tiles
    const bool pair = g_force_cta_group ? g_force_cta_group == 2 : a->M > BM;
    return pair ? launch_gemm<EPI, 2>(a, ep, a_shared, stream) : launch_gemm<EPI, 1>(a, ep, a_shared, stream);
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
        BSI_CHECK_ARG(a->gate.base && a->rows_per_sample > 0 && a->N % 64 == 0 && (reinterpret_cast<uintptr_t>(a->gate.base) & 15) == 0,
                      "bsi_gemm_bf16: GATE_RESID needs a 16-byte aligned gate, rows_per_sample and N %% 64 == 0");
    if (a->epilogue == BSI_EPI_POS_F32) BSI_CHECK_ARG(a->pos && a->rows_per_sample > 0, "bsi_gemm_bf16: POS needs pos table");
    if (a->epilogue == BSI_EPI_UNPATCH_F32)
        BSI_CHECK_ARG(a->patch > 0 && a->grid_w > 0 && a->channels > 0 && a->rows_per_sample % a->grid_w == 0 &&
                          a->N == a->patch * a->patch * a->channels && a->batch == 1,
                      "bsi_gemm_bf16: UNPATCH geometry invalid");

    const int a_shared = (a->batch > 1 && a->stride_a == 0) ? 1 : 0;
    EpiParams ep;
    ep.C = a->C, ep.bias = a->bias, ep.M = a->M, ep.N = a->N, ep.ldc = a->ldc;
    ep.stride_c = a->stride_c, ep.stride_bias = a->stride_bias;
    ep.gate = a->gate, ep.step_ptr = a->step_ptr, ep.rows_per_sample = a->rows_per_sample > 0 ? a->rows_per_sample : 1;
    ep.pos = a->pos, ep.patch = a->patch, ep.grid_w = a->grid_w, ep.channels = a->channels;

    cudaStream_t st = (cudaStream_t)stream;
    switch (a->epilogue) {
        case BSI_EPI_BIAS_BF16: return dispatch_cta_group<BSI_EPI_BIAS_BF16>(a, ep, a_shared, st);
        case BSI_EPI_BIAS_GELU_BF16: return dispatch_cta_group<BSI_EPI_BIAS_GELU_BF16>(a, ep, a_shared, st);
        case BSI_EPI_BIAS_SILU_BF16: return dispatch_cta_group<BSI_EPI_BIAS_SILU_BF16>(a, ep, a_shared, st);
        case BSI_EPI_BIAS_F32: return dispatch_cta_group<BSI_EPI_BIAS_F32>(a, ep, a_shared, st);
        case BSI_EPI_GATE_RESID_F32: return dispatch_cta_group<BSI_EPI_GATE_RESID_F32>(a, ep, a_shared, st);
        case BSI_EPI_POS_F32: return dispatch_cta_group<BSI_EPI_POS_F32>(a, ep, a_shared, st);
        case BSI_EPI_UNPATCH_F32: return dispatch_cta_group<BSI_EPI_UNPATCH_F32>(a, ep, a_shared, st);
        default: set_error("bsi_gemm_bf16: unknown epilogue %d", a->epilogue); return BSI_ERR_INVALID_ARGUMENT;
    }
}

// Test hook: force the single-CTA (1) or CTA-pair (2) kernel, 0 = automatic choice.
extern "C" int bsi_gemm_force_cta_group(int32_t cg) {
    BSI_CHECK_ARG(cg >= 0 && cg <= 2, "bsi_gemm_force_cta_group: expected 0, 1 or 2");
    g_force_cta_group = cg;
    return BSI_OK;
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
