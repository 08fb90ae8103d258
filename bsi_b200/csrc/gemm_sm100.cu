// bf16 x bf16 -> fp32 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA into 128B-swizzled shared memory, persistent CTAs, warp-specialised:
//   warp 0  : TMA producer (one lane)
//   warp 1  : MMA issuer   (one lane of the pair's leader CTA)
//   warp 2  : TMEM allocator
//   warps 4-7: epilogue: tcgen05.ld (64 columns per wait, next batch in flight) -> fused epilogue in registers ->
//              128B-swizzled staging tile in shared memory -> TMA store.  The fp32 residual update
//              (x += gate * branch) also *loads* its tile by TMA, two chunks ahead of the math.
// Two variants of the same kernel (template parameter CG):
//   CG = 2  CTA pair (cluster of 2, cta_group::2): UMMA 256 x 256 x 16, each CTA holds 128 rows of A and half of
//           the W tile, so every operand byte fetched from L2 feeds twice the math of the single-CTA tile.
//   CG = 1  single CTA, UMMA 128 x 256 x 16 (small problems, cross-check in the tests).
// The accumulator is double-buffered in TMEM (2 x 256 columns): the epilogue of tile i overlaps the main loop of
// tile i+1.  (First version: per-thread scattered 16-byte global stores straight from registers kept the tensor
// pipe 51 % busy on the QKV shape and 30 % on the out-projection — profiles/ncu_gemm_full_r01_first.csv.)
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
constexpr int BK = 64, UMMA_K = 16;  // the N tile (256, or 128 for narrow convolutions) is the template parameter BN
constexpr int kTmemCols = 512;
constexpr int kABytes = BM * BK * 2;
constexpr int kEpiBufBytes = BM * 128;  // staging tile: 128 rows x 128 B (64 bf16 or 32 fp32 columns)

// KIND_AUX16: bf16 output combined with a bf16 auxiliary tile of the same geometry that arrives by TMA (GELU' of the saved pre-activation)
enum EpiKind { KIND_BF16 = 0, KIND_F32 = 1, KIND_RMW = 2, KIND_SCATTER = 3, KIND_AUX16 = 4 };
__host__ __device__ constexpr int epi_kind(int epi) {
    return epi == BSI_EPI_MUL_GELU_GRAD_BF16 ? KIND_AUX16
           : (epi <= BSI_EPI_BIAS_SILU_BF16 || epi == BSI_EPI_MOD_SILU_BF16 || epi == BSI_EPI_BIAS_GELU_DUAL_BF16) ? KIND_BF16
           : epi == BSI_EPI_GATE_RESID_F32 ? KIND_RMW
           : epi == BSI_EPI_UNPATCH_F32  ? KIND_SCATTER
                                         : KIND_F32;
}

// pipeline-depth experiments (python -m bsi_b200.build --variant NAME -DBSI_EXP_...): defaults are the product configuration
#ifndef BSI_EXP_PLAIN_STAGES
#define BSI_EXP_PLAIN_STAGES 5
#endif
// GELU / SiLU epilogues: 5 pipeline stages with one staging tile per warp group beat 4 stages with two (MLP-1 shape, M = 65536,
// back-to-back under the power cap: 484 vs 496-500 us, profiles/gemm_ab_r02.jsonl); plain epilogue: 5 vs 4 stages is within 1 %
#ifndef BSI_EXP_HEAVY_STAGES
#define BSI_EXP_HEAVY_STAGES 5
#endif
#ifndef BSI_EXP_HEAVY_BUFS
#define BSI_EXP_HEAVY_BUFS 1
#endif

constexpr int kReuseMaxW = 32;  // widest image row for which a convolution's A box (128 + 2W pixels) fits the stage

template <int EPI, int CG, int BN, bool CONV = false>
struct Cfg {
    static constexpr int kKind = epi_kind(EPI);
    static constexpr int kBRows = BN / CG;  // rows of the W tile this CTA loads
    static constexpr int kBBytes = kBRows * BK * 2;
    // Narrow (N <= 128) 3x3 convolutions are bound by L2 -> SM traffic, most of it the A tile re-fetched for every tap.  Their
    // stage therefore holds one A box with a one-row halo above and below ((128 + 2W) pixels) that serves the three taps
    // dy = -1, 0, +1 of one dx through three smem descriptors W pixels apart, plus those three taps' weight tiles.
    static constexpr bool kRowReuse = CONV && BN == 128;
    static constexpr int kAStage = kRowReuse ? (BM + 2 * kReuseMaxW) * 128 : kABytes;
    static constexpr int kBTiles = kRowReuse ? 3 : 1;
    static constexpr int kStageBytes = kAStage + kBTiles * kBBytes;
    // bf16 epilogues (bias / GELU / SiLU / modulation) are ALU-bound with one warp per scheduler (ncu: 25 % issue utilisation,
    // tensor pipe 63 % on the GELU shape): they get 8 epilogue warps, two per TMEM lane quarter, each pair splitting the columns
    // (the plain bias epilogue keeps 4 warps and the fifth pipeline stage: it already holds the tensor pipe at 87 %)
    static constexpr bool kDual = EPI == BSI_EPI_BIAS_GELU_DUAL_BF16;  // two bf16 outputs per tile: one staging tile each
    static constexpr bool kHeavy = EPI == BSI_EPI_BIAS_GELU_BF16 || EPI == BSI_EPI_BIAS_SILU_BF16 || EPI == BSI_EPI_MOD_SILU_BF16 || kDual;
    // (16 warps, one 64-column block each, measured no faster than 8: the GELU epilogue is bound by the fp32 pipe, not by latency)
    // (the GELU' epilogue of the training path is the heaviest of all -- ~20 instructions per element: 8 warps as well)
    static constexpr int kEpiWarps = (kHeavy || epi_kind(EPI) == KIND_AUX16) ? 8 : 4;
    static constexpr int kGroups = kEpiWarps / 4;                                   // groups of 4 warps, each owning BN / kGroups columns
    static constexpr int kBufsPerGroup = kDual ? 2 : kHeavy ? (kGroups == 4 ? 1 : BSI_EXP_HEAVY_BUFS) : 2;  // staging tiles per group (KIND_BF16)
    static constexpr int kThreads = 128 + 32 * kEpiWarps;
    static constexpr int kEpiBufs = (kKind == KIND_RMW || kKind == KIND_AUX16) ? 4 : kHeavy ? kGroups * kBufsPerGroup : (kKind == KIND_SCATTER ? 0 : 2);
    static constexpr int kVecBytes = 2 * 3 * BN * 4;  // bias, gate/scale and shift slices of the tile, double-buffered by tile parity
    static constexpr int kStages = kRowReuse ? (232448 - 1280 - 1024 /*kGnBytes*/ - kVecBytes - kEpiBufs * kEpiBufBytes) / kStageBytes
                                   : BN == 128 ? (CG == 2 ? 6 : 4)
                                               : (CG == 2 ? ((kKind == KIND_RMW || kKind == KIND_AUX16 || kDual) ? 4 : kHeavy ? BSI_EXP_HEAVY_STAGES : BSI_EXP_PLAIN_STAGES) : 3);
    static_assert(kStages >= 2, "pipeline needs two stages");
    // GroupNorm partial sums of a convolution's fp32 output (U-Net residual stream): [4 warps][32 groups][sum, sum of squares]
    static constexpr bool kGnStats = CONV && BN == 128 && kKind == KIND_RMW;
    static constexpr int kGnBytes = kGnStats ? 4 * 64 * 4 : 0;
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBufs * kEpiBufBytes + kVecBytes + 1024 /*align*/ + 256 /*barriers*/ + kGnBytes;
    static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory per CTA");
};

struct EpiParams {
    void* C;
    const float* bias;
    int M, N, ldc;
    long long stride_c, stride_bias;
    bsi_rowref gate;   // GATE_RESID: gate (base NULL = 1); MOD_SILU: scale
    bsi_rowref shift;  // MOD_SILU: shift
    const int* step_ptr;
    int rows_per_sample;
    const float* pos;
    int patch, grid_w, channels;
    float* gn_partial;  // CONV, N = 128, GATE_RESID: [M / 128][32][2] per-tile GroupNorm sums of the fp32 output (NULL = off)
};

// Implicit-GEMM 3x3 / 1x1 convolution over NHWC activations: the A tile of k-block (tap, channel block) is a 4-D TMA box
// {64 channels, W, 128/W rows, 1 image} shifted by the tap offset; out-of-image elements are zero-filled by TMA, which is
// exactly the zero padding of nn.Conv2d(padding=1).  Two sources implement the channel concat of the U-Net up path.
struct ConvGeom {
    int taps;      // 1 or 9
    int cb1, cbt;  // 64-channel blocks of source 1, and of both sources together
    int img_w, img_hw;
    int reuse;  // 1: k-groups (dx, channel block), one haloed A box + three weight tiles per stage (Cfg::kRowReuse)
};

__device__ __forceinline__ float gelu_tanh(float x) {
    // 0.5*x*(1+tanh(sqrt(2/pi)*(x+0.044715x^3)))  (nn.GELU(approximate="tanh"), bsi/models/dit.py:75) in 5 fp32 ops + 1 MUFU:
    // the epilogue of the MLP GEMM is bound by the fp32 pipe (2 clk per warp instruction), so every op counts
    const float x2 = x * x;
    const float u = x * fmaf(x2, 0.7978845608028654f * 0.044715f, 0.7978845608028654f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float gelu_tanh_grad(float x) {
    // d/dx gelu_tanh(x) = 0.5 (1 + t) + 0.5 x (1 - t^2) u'(x),  t = tanh(u), u = sqrt(2/pi) (x + 0.044715 x^3)
    const float x2 = x * x;
    const float u = x * fmaf(x2, 0.7978845608028654f * 0.044715f, 0.7978845608028654f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float du = fmaf(x2, 3.0f * 0.044715f * 0.7978845608028654f, 0.7978845608028654f);
    const float hx = 0.5f * x;
    return fmaf(hx * fmaf(-t, t, 1.0f), du, fmaf(0.5f, t, 0.5f));
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint4 ld_shared_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// Sum 16 per-lane values over the 32 lanes of a warp with 16 shuffles (each step halves the values a lane keeps); returns the total
// of value (lane >> 1), held by both lanes of the pair.  Fixed order: deterministic.
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool hi = lane & 16;
        const float send = hi ? v[i] : v[i + 8], keep = hi ? v[i + 8] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool hi = lane & 8;
        const float send = hi ? v[i] : v[i + 4], keep = hi ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const bool hi = lane & 4;
        const float send = hi ? v[i] : v[i + 2], keep = hi ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const bool hi = lane & 2;
        const float send = hi ? v[0] : v[1], keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// named barrier of one group of 4 epilogue warps (group 0 or 1)
__device__ __forceinline__ void epi_bar(int group = 0) { asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
// byte offset of 16-byte chunk `c` of row `r` inside a 128B-swizzled staging tile (what TMA SWIZZLE_128B expects)
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

template <int EPI, int CG, bool CONV, int BN>
__global__ void __launch_bounds__(Cfg<EPI, CG, BN, CONV>::kThreads, 1)
    k_gemm_bf16(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_c,
                const __grid_constant__ CUtensorMap map_r, const EpiParams ep, const ConvGeom geo, const int m_tiles, const int n_tiles,
                const int k_blocks, const int batch, const int a_shared) {
    using C = Cfg<EPI, CG, BN, CONV>;
    constexpr int kStages = C::kStages, kKind = C::kKind;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_buf = smem + kStages * C::kStageBytes;                       // kEpiBufs x 16 KB, 1024-aligned
    float* s_vec = reinterpret_cast<float*>(epi_buf + C::kEpiBufs * kEpiBufBytes);  // [parity][bias|gate][256]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_vec) + C::kVecBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* c_full = tmem_empty + 2;  // residual chunk landed (KIND_RMW), one per staging buffer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c_full + 4);
    float* s_gn = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);  // [4][64], only when C::kGnStats

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = batch * m_tiles * n_tiles;
    const int cta_rank = CG == 2 ? (int)ptx::cluster_ctarank() : 0;  // rank inside the CTA pair; 0 issues the MMAs
    const int worker = CG == 2 ? blockIdx.x / 2 : blockIdx.x;          // persistent worker = CTA or CTA pair
    const int num_workers = CG == 2 ? gridDim.x / 2 : gridDim.x;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_w);
        if (kKind != KIND_SCATTER) ptx::prefetch_tensormap(&map_c);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);   // leader's arrive.expect_tx; TMA bytes of both CTAs land on the leader's barrier
            ptx::mbar_init(&empty_bar[s], 1);  // tcgen05.commit (multicast to both CTAs of a pair)
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&tmem_full[b], 1);        // tcgen05.commit after the last k-block
            ptx::mbar_init(&tmem_empty[b], C::kEpiWarps * CG);  // one arrive per epilogue warp of every CTA of the pair
        }
        for (int b = 0; b < 4; ++b) ptx::mbar_init(&c_full[b], 1);
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
    pdl_prologue_done();  // everything above overlapped the previous kernel's tail; operands / residual / step counter are read below

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
                    if constexpr (CONV) {
                        const int img = row_a / geo.img_hw, y0 = (row_a - img * geo.img_hw) / geo.img_w;
                        if (C::kRowReuse && geo.reuse) {
                            // k-group (dx, channel block): rows y0-1 .. y0+128/W of the image shifted by dx, and the weights of taps
                            // (dy, dx), dy = -1, 0, +1 (tap-major K layout of bsi_pack_conv_weight)
                            const int dxi = kb / geo.cbt, cb = kb - dxi * geo.cbt;
                            const CUtensorMap* src = cb < geo.cb1 ? &map_a : &map_a2;
                            const int c0 = (cb < geo.cb1 ? cb : cb - geo.cb1) * BK;
                            const uint32_t bytes = (uint32_t)(BM + 2 * geo.img_w) * 128u + 3u * C::kBBytes;
                            if constexpr (CG == 1) {
                                ptx::mbar_arrive_expect_tx(&full_bar[stage], bytes);
                                ptx::tma_load_4d(sa, src, &full_bar[stage], c0, dxi - 1, y0 - 1, img);
                                for (int dyi = 0; dyi < 3; ++dyi)
                                    ptx::tma_load_3d(sa + C::kAStage + dyi * C::kBBytes, &map_w, &full_bar[stage], ((dyi * 3 + dxi) * geo.cbt + cb) * BK, row_w, b);
                            } else {
                                if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * bytes);
                                ptx::tma_load_4d_2sm(sa, src, &full_bar[stage], c0, dxi - 1, y0 - 1, img);
                                for (int dyi = 0; dyi < 3; ++dyi)
                                    ptx::tma_load_3d_2sm(sa + C::kAStage + dyi * C::kBBytes, &map_w, &full_bar[stage], ((dyi * 3 + dxi) * geo.cbt + cb) * BK, row_w, b);
                            }
                        } else {
                            const int tap = kb / geo.cbt, cb = kb - tap * geo.cbt;
                            const int dy = geo.taps == 9 ? tap / 3 - 1 : 0, dx = geo.taps == 9 ? tap % 3 - 1 : 0;
                            const CUtensorMap* src = cb < geo.cb1 ? &map_a : &map_a2;
                            const int c0 = (cb < geo.cb1 ? cb : cb - geo.cb1) * BK;
                            if constexpr (CG == 1) {
                                ptx::mbar_arrive_expect_tx(&full_bar[stage], kABytes + C::kBBytes);
                                ptx::tma_load_4d(sa, src, &full_bar[stage], c0, dx, y0 + dy, img);
                                ptx::tma_load_3d(sa + C::kAStage, &map_w, &full_bar[stage], kb * BK, row_w, b);
                            } else {
                                if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * (kABytes + C::kBBytes));
                                ptx::tma_load_4d_2sm(sa, src, &full_bar[stage], c0, dx, y0 + dy, img);
                                ptx::tma_load_3d_2sm(sa + C::kAStage, &map_w, &full_bar[stage], kb * BK, row_w, b);
                            }
                        }
                    } else if constexpr (CG == 1) {
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
                        ptx::tma_load_3d(sa, &map_a, &full_bar[stage], kb * BK, row_a, a_shared ? 0 : b);
                        ptx::tma_load_3d(sa + C::kAStage, &map_w, &full_bar[stage], kb * BK, row_w, b);
                    } else {
                        if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
                        ptx::tma_load_3d_2sm(sa, &map_a, &full_bar[stage], kb * BK, row_a, a_shared ? 0 : b);
                        ptx::tma_load_3d_2sm(sa + C::kAStage, &map_w, &full_bar[stage], kb * BK, row_w, b);
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
                    if (C::kRowReuse && geo.reuse) {
                        // tap dy reads the 128 pixels that start dy+1 image rows into the haloed box (W * 128 B: a multiple of 1024)
#pragma unroll
                        for (int dyi = 0; dyi < 3; ++dyi) {
                            const uint64_t da = ptx::umma_desc_k_sw128(sa + dyi * geo.img_w * 128);
                            const uint64_t db = ptx::umma_desc_k_sw128(sa + C::kAStage + dyi * C::kBBytes);
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k) ptx::umma_bf16_ss<CG>(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | dyi | k) != 0 ? 1u : 0u);
                        }
                    } else {
                        const uint64_t da = ptx::umma_desc_k_sw128(sa), db = ptx::umma_desc_k_sw128(sa + C::kAStage);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            // advance 32 B (16 bf16) inside the 128 B swizzle row: +2 in 16-byte address units
                            ptx::umma_bf16_ss<CG>(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
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
        // et: epilogue thread; er = row inside the CTA tile (TMEM lane); hf = column half handled by this warp (8-warp epilogues)
        const int q = warp & 3, et = threadIdx.x - 128, er = et & 127, hf = et >> 7;
        constexpr int kBatches = (BN / 64) / C::kGroups;  // 64-column batches per warp
        constexpr int kChunkCols = kKind == KIND_AUX16 ? 64 : 32;  // columns per 128-byte staging row: 32 fp32 or 64 bf16
        constexpr int kChunks = BN / kChunkCols;  // chunks of a tile (KIND_RMW / KIND_F32 / KIND_AUX16)
        const int step = ep.step_ptr ? *ep.step_ptr : 0;
        const uint32_t buf0 = ptx::smem_u32(epi_buf);
        const int my_tiles = worker < total_tiles ? (total_tiles - worker + num_workers - 1) / num_workers : 0;

        // KIND_RMW: residual chunk g (8 per tile, 32 fp32 columns each) of this CTA's tile sequence -> staging buffer g & 3
        auto issue_residual_load = [&](int g) {
            if (g >= my_tiles * kChunks) return;
            const int tile = worker + (g / kChunks) * num_workers, c = g % kChunks;
            const int n_t = tile % n_tiles, m_t = (tile / n_tiles) % m_tiles, b = tile / (n_tiles * m_tiles);
            ptx::mbar_arrive_expect_tx(&c_full[g & 3], kEpiBufBytes);
            ptx::tma_load_3d(epi_buf + (g & 3) * kEpiBufBytes, &map_r, &c_full[g & 3], n_t * BN + c * kChunkCols, (m_t * CG + cta_rank) * BM, b);
        };
        if constexpr (kKind == KIND_RMW) {
            if (et == 0) {
                issue_residual_load(0);
                issue_residual_load(1);
            }
        }
        // KIND_AUX16: each group of 4 warps (hf) owns two of the tile's four 64-column chunks and two staging tiles; gl is the group's
        // running chunk index, its auxiliary tile goes to staging tile hf*2 + (gl & 1)
        auto issue_aux_load = [&](int gl) {
            if (gl >= my_tiles * 2) return;
            const int tile = worker + (gl >> 1) * num_workers, c = hf * 2 + (gl & 1);
            const int n_t = tile % n_tiles, m_t = (tile / n_tiles) % m_tiles, b = tile / (n_tiles * m_tiles);
            uint64_t* bar = &c_full[hf * 2 + (gl & 1)];
            ptx::mbar_arrive_expect_tx(bar, kEpiBufBytes);
            ptx::tma_load_3d(epi_buf + (hf * 2 + (gl & 1)) * kEpiBufBytes, &map_r, bar, n_t * BN + c * 64, (m_t * CG + cta_rank) * BM, b);
        };
        if constexpr (kKind == KIND_AUX16) {
            if (er == 0) issue_aux_load(0);
        }

        int it = 0;
        for (int tile = worker; tile < total_tiles; tile += num_workers, ++it) {
            const int n_t = tile % n_tiles, m_t = (tile / n_tiles) % m_tiles, b = tile / (n_tiles * m_tiles);
            const int buf = it & 1;
            const int row_base = (m_t * CG + cta_rank) * BM, n_base = n_t * BN;
            const int row = row_base + er;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;

            // stage the bias (and gate) slice of this tile in shared memory: every thread needs all 256 columns
            float* s_bias = s_vec + (it & 1) * 3 * BN;
            float* s_gate = s_bias + BN;
            float* s_shift = s_gate + BN;
            {
                const float* bias = ep.bias ? ep.bias + (long long)b * ep.stride_bias : nullptr;
                // warp group hf stages exactly the columns [hf*BN/kGroups, (hf+1)*BN/kGroups) it reads later
                for (int sc = hf * (BN / C::kGroups) + er; sc < (hf + 1) * (BN / C::kGroups); sc += 128) {
                    const int n = n_base + sc;
                    s_bias[sc] = (bias && n < ep.N) ? bias[n] : 0.0f;
                    // all 128 rows of the CTA tile belong to one sample (rows_per_sample % 128 == 0, checked on the host)
                    const bool in_range = n < ep.N && row_base < ep.M;
                    if constexpr (kKind == KIND_RMW)
                        s_gate[sc] = !in_range ? 0.0f : (ep.gate.base ? rowref_ptr(ep.gate, row_base / ep.rows_per_sample, step)[n] : 1.0f);
                    if constexpr (EPI == BSI_EPI_MOD_SILU_BF16) {
                        s_gate[sc] = in_range ? rowref_ptr(ep.gate, row_base / ep.rows_per_sample, step)[n] : 0.0f;
                        s_shift[sc] = in_range ? rowref_ptr(ep.shift, row_base / ep.rows_per_sample, step)[n] : 0.0f;
                    }
                }
            }
            epi_bar(hf);
            ptx::mbar_wait(&tmem_full[buf], (it >> 1) & 1);
            ptx::tc_fence_after();

            uint32_t acc[2][2][32];  // [batch parity][half][32 columns]: 64 columns per tcgen05.wait::ld, next batch in flight
            const int jb0 = hf * kBatches;
            ptx::tmem_ld_32x32b_x32(taddr + jb0 * 64, acc[0][0]);
            ptx::tmem_ld_32x32b_x32(taddr + jb0 * 64 + 32, acc[0][1]);
#pragma unroll
            for (int jl = 0; jl < kBatches; ++jl) {
                const int jb = jb0 + jl;
                ptx::tmem_ld_wait();
                if (jl + 1 < kBatches) {
                    ptx::tmem_ld_32x32b_x32(taddr + (jb + 1) * 64, acc[(jl + 1) & 1][0]);
                    ptx::tmem_ld_32x32b_x32(taddr + (jb + 1) * 64 + 32, acc[(jl + 1) & 1][1]);
                } else {
                    // all TMEM reads of this buffer are done: hand it back to the MMA warp before the remaining stores
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (CG == 1 || cta_rank == 0) ptx::mbar_arrive(&tmem_empty[buf]);
                        else ptx::mbar_arrive_cluster(&tmem_empty[buf], 0);
                    }
                }
                const uint32_t(&a)[2][32] = acc[jl & 1];

                if constexpr (kKind == KIND_BF16 && C::kDual) {
                    // ---- 64 columns -> pre-activation tile (aux) and GELU tile (C), one staging tile and one TMA store each
                    const int sbuf = hf * 2;
                    const uint32_t sb_pre = buf0 + sbuf * kEpiBufBytes, sb_act = sb_pre + kEpiBufBytes;
                    if (er == 0) ptx::tma_store_wait_read<0>();  // both tiles of the previous batch have been read by their stores
                    epi_bar(hf);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            uint32_t pw[4], gw[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int col = jb * 64 + h * 32 + c * 8 + 2 * i;
                                const uint32_t pk = pack_bf16(__uint_as_float(a[h][c * 8 + 2 * i]) + s_bias[col], __uint_as_float(a[h][c * 8 + 2 * i + 1]) + s_bias[col + 1]);
                                pw[i] = pk;  // the activation is taken of the ROUNDED pre-activation: what the backward differentiates
                                gw[i] = pack_bf16(gelu_tanh(bf16_lo(pk)), gelu_tanh(bf16_hi(pk)));
                            }
                            st_shared_v4(sb_pre + swz(er, h * 4 + c), pw[0], pw[1], pw[2], pw[3]);
                            st_shared_v4(sb_act + swz(er, h * 4 + c), gw[0], gw[1], gw[2], gw[3]);
                        }
                    }
                    ptx::fence_proxy_async();
                    epi_bar(hf);
                    if (er == 0) {
                        ptx::tma_store_3d(&map_c, epi_buf + (sbuf + 1) * kEpiBufBytes, n_base + jb * 64, row_base, b);
                        ptx::tma_store_3d(&map_r, epi_buf + sbuf * kEpiBufBytes, n_base + jb * 64, row_base, b);
                        ptx::tma_store_commit();
                    }
                } else if constexpr (kKind == KIND_BF16) {
                    // ---- 64 bf16 columns -> one 128 x 128 B staging tile -> TMA store
                    const int sbuf = hf * C::kBufsPerGroup + (jl % C::kBufsPerGroup);  // staging tiles of this warp group
                    const uint32_t sb = buf0 + sbuf * kEpiBufBytes;
                    if (er == 0) ptx::tma_store_wait_read<C::kBufsPerGroup - 1>();  // the store that last read this buffer has drained
                    epi_bar(hf);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            float w[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int col = jb * 64 + h * 32 + c * 8 + i;
                                float t = __uint_as_float(a[h][c * 8 + i]) + s_bias[col];
                                if constexpr (EPI == BSI_EPI_BIAS_GELU_BF16) t = gelu_tanh(t);
                                if constexpr (EPI == BSI_EPI_BIAS_SILU_BF16) t = silu(t);
                                // FeatureModulation + SiLU: silu(shift + (1 + scale) * h)   (bsi/nn/residual_block.py:19-21,45-46)
                                if constexpr (EPI == BSI_EPI_MOD_SILU_BF16) t = silu(fmaf(s_gate[col] + 1.0f, t, s_shift[col]));
                                w[i] = t;
                            }
                            st_shared_v4(sb + swz(er, h * 4 + c), pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]), pack_bf16(w[4], w[5]),
                                         pack_bf16(w[6], w[7]));
                        }
                    }
                    ptx::fence_proxy_async();
                    epi_bar(hf);
                    if (er == 0) {
                        ptx::tma_store_3d(&map_c, epi_buf + sbuf * kEpiBufBytes, n_base + jb * 64, row_base, b);
                        ptx::tma_store_commit();
                    }
                } else if constexpr (kKind == KIND_F32) {
                    // ---- 2 x (32 fp32 columns -> staging tile -> TMA store)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int cidx = jb * 2 + h;  // 32-column chunk of the tile
                        const uint32_t sb = buf0 + (cidx & 1) * kEpiBufBytes;
                        if (et == 0) ptx::tma_store_wait_read<1>();
                        epi_bar();
                        const float* pos = nullptr;
                        if constexpr (EPI == BSI_EPI_POS_F32) pos = ep.pos + (long long)(row % ep.rows_per_sample) * ep.N + n_base + cidx * 32;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            float4 v;
                            v.x = __uint_as_float(a[h][c * 4 + 0]) + s_bias[cidx * 32 + c * 4 + 0];
                            v.y = __uint_as_float(a[h][c * 4 + 1]) + s_bias[cidx * 32 + c * 4 + 1];
                            v.z = __uint_as_float(a[h][c * 4 + 2]) + s_bias[cidx * 32 + c * 4 + 2];
                            v.w = __uint_as_float(a[h][c * 4 + 3]) + s_bias[cidx * 32 + c * 4 + 3];
                            if constexpr (EPI == BSI_EPI_POS_F32) {
                                if (n_base + cidx * 32 + c * 4 < ep.N) {
                                    const float4 p4 = *reinterpret_cast<const float4*>(pos + c * 4);
                                    v.x += p4.x, v.y += p4.y, v.z += p4.z, v.w += p4.w;
                                }
                            }
                            st_shared_v4(sb + swz(et, c), __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
                        }
                        ptx::fence_proxy_async();
                        epi_bar();
                        if (et == 0) {
                            ptx::tma_store_3d(&map_c, epi_buf + (cidx & 1) * kEpiBufBytes, n_base + cidx * 32, row_base, b);
                            ptx::tma_store_commit();
                        }
                    }
                } else if constexpr (kKind == KIND_RMW) {
                    // ---- x = addcmul(x, gate, branch) (bsi/models/dit.py:93-102): residual chunk arrives by TMA (issued two chunks
                    //      ahead), is updated in place in shared memory and goes back by TMA store
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int cidx = jb * 2 + h;
                        const int g = it * kChunks + cidx;  // running chunk index of this CTA
                        if (et == 0) {
                            ptx::tma_store_wait_read<1>();  // buffer (g+2)&3 was last read by the store of chunk g-2
                            issue_residual_load(g + 2);
                        }
                        ptx::mbar_wait(&c_full[g & 3], (g >> 2) & 1);
                        const uint32_t sb = buf0 + (g & 3) * kEpiBufBytes;
                        float gn[16];  // C::kGnStats: (sum, sum of squares) of this row's 8 four-channel groups of the chunk
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint32_t addr = sb + swz(et, c);
                            float4 x4 = ld_shared_f4(addr);
                            const float* gt = s_gate + cidx * 32 + c * 4;
                            const float* bs = s_bias + cidx * 32 + c * 4;
                            x4.x = fmaf(gt[0], __uint_as_float(a[h][c * 4 + 0]) + bs[0], x4.x);
                            x4.y = fmaf(gt[1], __uint_as_float(a[h][c * 4 + 1]) + bs[1], x4.y);
                            x4.z = fmaf(gt[2], __uint_as_float(a[h][c * 4 + 2]) + bs[2], x4.z);
                            x4.w = fmaf(gt[3], __uint_as_float(a[h][c * 4 + 3]) + bs[3], x4.w);
                            st_shared_v4(addr, __float_as_uint(x4.x), __float_as_uint(x4.y), __float_as_uint(x4.z), __float_as_uint(x4.w));
                            if constexpr (C::kGnStats) {
                                gn[2 * c] = (x4.x + x4.y) + (x4.z + x4.w);
                                gn[2 * c + 1] = fmaf(x4.x, x4.x, x4.y * x4.y) + fmaf(x4.z, x4.z, x4.w * x4.w);
                            }
                        }
                        if constexpr (C::kGnStats) {
                            if (ep.gn_partial) {  // uniform branch
                                const float tot = warp_reduce16(gn, lane);
                                if ((lane & 1) == 0) s_gn[q * 64 + cidx * 16 + (lane >> 1)] = tot;
                            }
                        }
                        ptx::fence_proxy_async();
                        epi_bar();
                        if (et == 0) {
                            ptx::tma_store_3d(&map_c, epi_buf + (g & 3) * kEpiBufBytes, n_base + cidx * 32, row_base, b);
                            ptx::tma_store_commit();
                        }
                    }
                } else if constexpr (kKind == KIND_AUX16) {
                    // ---- out = (acc + bias) * gelu'(pre): the group's pre-activation chunk (64 bf16 columns) arrives by TMA one chunk ahead,
                    //      is replaced in place by the product and goes out by TMA store
                    const int gl = it * 2 + jl, sbuf = hf * 2 + (gl & 1);
                    if (er == 0) {
                        ptx::tma_store_wait_read<0>();  // the other staging tile of this group has been read by the previous chunk's store
                        issue_aux_load(gl + 1);
                    }
                    ptx::mbar_wait(&c_full[sbuf], (gl >> 1) & 1);
                    const uint32_t sb = buf0 + sbuf * kEpiBufBytes;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint32_t addr = sb + swz(er, c);
                        const uint4 p4 = ld_shared_u4(addr);
                        const uint32_t pw[4] = {p4.x, p4.y, p4.z, p4.w};
                        uint32_t ow[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int col = jb * 64 + c * 8 + 2 * i;
                            const float v0 = __uint_as_float(a[c >> 2][(c & 3) * 8 + 2 * i]) + s_bias[col];
                            const float v1 = __uint_as_float(a[c >> 2][(c & 3) * 8 + 2 * i + 1]) + s_bias[col + 1];
                            ow[i] = pack_bf16(v0 * gelu_tanh_grad(bf16_lo(pw[i])), v1 * gelu_tanh_grad(bf16_hi(pw[i])));
                        }
                        st_shared_v4(addr, ow[0], ow[1], ow[2], ow[3]);
                    }
                    ptx::fence_proxy_async();
                    epi_bar(hf);
                    if (er == 0) {
                        ptx::tma_store_3d(&map_c, epi_buf + sbuf * kEpiBufBytes, n_base + jb * 64, row_base, b);
                        ptx::tma_store_commit();
                    }
                } else {
                    // ---- KIND_SCATTER: b (nh nw) (ph pw c) -> b c (nh ph) (nw pw)   (bsi/models/dit.py:166-172); N = p*p*c is tiny
                    if (row < ep.M) {
                        float* out = reinterpret_cast<float*>(ep.C);
                        const int T = ep.rows_per_sample, p = ep.patch, gw = ep.grid_w, ch = ep.channels;
                        const int bb = row / T, tok = row - bb * T;
                        const int gy = tok / gw, gx = tok - gy * gw;
                        const int Wimg = gw * p, Himg = (T / gw) * p;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int nl = jb * 64 + h * 32 + j, n = n_base + nl;
                                if (n < ep.N) {
                                    const int cc = n % ch, within = n / ch;
                                    const int py = within / p, px = within - py * p;
                                    out[(((long long)bb * ch + cc) * Himg + gy * p + py) * Wimg + gx * p + px] = __uint_as_float(a[h][j]) + s_bias[nl];
                                }
                            }
                        }
                    }
                }
            }
            if constexpr (C::kGnStats) {
                if (ep.gn_partial) {
                    // all four warps have written their 64 partial sums of this tile (the last chunk's epi_bar above orders them)
                    if (et < 64) {
                        const float t = (s_gn[et] + s_gn[64 + et]) + (s_gn[128 + et] + s_gn[192 + et]);
                        ep.gn_partial[(long long)(row_base / BM) * 64 + et] = t;
                    }
                }
            }
        }
        if (er == 0) ptx::tma_store_wait_all<0>();  // every staged tile has reached global memory
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

// [batch][rows][ld] tensor of `esize`-byte elements, box {128 B of columns, box_rows, 1}, 128-byte swizzle, zero fill out of bounds.
int make_tile_map(CUtensorMap* map, const void* base, int esize, int64_t rows, int64_t cols, int64_t ld, int64_t batch, int64_t batch_stride,
                  int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return BSI_ERR_CUDA;
    }
    const int per16 = 16 / esize;
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % per16) != 0 || (batch > 1 && (batch_stride % per16) != 0)) {
        set_error("GEMM operand must be 16-byte aligned with a pitch multiple of 16 bytes (ld=%lld)", (long long)ld);
        return BSI_ERR_INVALID_ARGUMENT;
    }
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(batch > 1 ? batch : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * esize, (cuuint64_t)(batch > 1 ? batch_stride : rows * ld) * esize};
    cuuint32_t box[3] = {(cuuint32_t)(128 / esize), (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld batch=%lld esize=%d)", (int)r, (long long)rows,
                  (long long)cols, (long long)ld, (long long)batch, esize);
        return BSI_ERR_CUDA;
    }
    return BSI_OK;
}

// bf16 NHWC [B][H][W][C] activation tensor, box {64 channels, W, 128/W rows, 1 image}, 128-byte swizzle, zero fill outside.
static int make_nhwc_map(CUtensorMap* map, const void* base, int64_t B, int64_t H, int64_t W, int64_t C, int halo_rows = 0) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return BSI_ERR_CUDA;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) || C % 64 != 0 || BM % W != 0 || H % (BM / W) != 0) {
        set_error("conv operand needs 16-byte alignment, C %% 64 == 0, W | 128 and H %% (128/W) == 0 (H=%lld W=%lld C=%lld)", (long long)H,
                  (long long)W, (long long)C);
        return BSI_ERR_INVALID_ARGUMENT;
    }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)W, (cuuint32_t)(BM / W + halo_rows), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (NHWC) failed with CUresult %d", (int)r);
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

// One GEMM / implicit-GEMM convolution problem, common to bsi_gemm_bf16 and bsi_conv_bf16.
struct Problem {
    const void *A = nullptr, *A2 = nullptr, *W = nullptr;
    void* C = nullptr;
    const float* resid = nullptr;  // RMW residual source (fp32, same geometry as C); nullptr = C itself
    const void* aux = nullptr;     // bf16 second output (GELU_DUAL) / auxiliary input (MUL_GELU_GRAD), geometry of C
    int M = 0, N = 0, K = 0, lda = 0, ldw = 0, ldc = 0, batch = 1, a_shared = 0;
    int64_t stride_a = 0, stride_w = 0, stride_c = 0;
    // convolution geometry (conv == true): A/A2 are NHWC [B][H][W][C1|C2]
    bool conv = false;
    int img_b = 0, img_h = 0, img_w = 0, c1 = 0, c2 = 0, taps = 1;
    EpiParams ep{};
};

template <int EPI, int CG, bool CONV, int BN>
static int launch_gemm(const Problem& p, cudaStream_t stream) {
    using C = Cfg<EPI, CG, BN, CONV>;
    BSI_ENSURE_SMEM((k_gemm_bf16<EPI, CG, CONV, BN>), C::kSmemBytes);
    CUtensorMap ma, ma2, mw, mc, mr;
    int rc;
    ConvGeom geo{};
    if (CONV) {
        // W * 128 B must keep the shifted descriptors 1024-byte aligned (W % 8 == 0) and the haloed box must fit the stage
        geo.reuse = C::kRowReuse && p.taps == 9 && p.img_w <= kReuseMaxW && p.img_w % 8 == 0;
        rc = make_nhwc_map(&ma, p.A, p.img_b, p.img_h, p.img_w, p.c1, geo.reuse ? 2 : 0);
        if (rc != BSI_OK) return rc;
        ma2 = ma;
        if (p.A2) {
            rc = make_nhwc_map(&ma2, p.A2, p.img_b, p.img_h, p.img_w, p.c2, geo.reuse ? 2 : 0);
            if (rc != BSI_OK) return rc;
        }
        geo.taps = p.taps, geo.cb1 = p.c1 / BK, geo.cbt = (p.c1 + (p.A2 ? p.c2 : 0)) / BK;
        geo.img_w = p.img_w, geo.img_hw = p.img_h * p.img_w;
    } else {
        rc = make_tile_map(&ma, p.A, 2, p.M, p.K, p.lda, p.a_shared ? 1 : p.batch, p.stride_a, BM);
        if (rc != BSI_OK) return rc;
        ma2 = ma;
    }
    rc = make_tile_map(&mw, p.W, 2, p.N, p.K, p.ldw, p.batch, p.stride_w, C::kBRows);
    if (rc != BSI_OK) return rc;
    if (C::kKind != KIND_SCATTER) {
        rc = make_tile_map(&mc, p.C, (C::kKind == KIND_BF16 || C::kKind == KIND_AUX16) ? 2 : 4, p.M, p.N, p.ldc, p.batch, p.stride_c, BM);
        if (rc != BSI_OK) return rc;
    } else {
        mc = ma;  // unused by the scatter epilogue
    }
    mr = mc;
    if (C::kKind == KIND_RMW && p.resid && p.resid != p.C) {
        rc = make_tile_map(&mr, p.resid, 4, p.M, p.N, p.ldc, p.batch, p.stride_c, BM);
        if (rc != BSI_OK) return rc;
    }
    if (C::kDual || C::kKind == KIND_AUX16) {
        rc = make_tile_map(&mr, p.aux, 2, p.M, p.N, p.ldc, p.batch, p.stride_c, BM);
        if (rc != BSI_OK) return rc;
    }
    const int m_tiles = (p.M + BM * CG - 1) / (BM * CG), n_tiles = (p.N + BN - 1) / BN;
    const int k_blocks = geo.reuse ? 3 * geo.cbt : (p.K + BK - 1) / BK;  // row-reuse stages carry three taps each
    const int total = p.batch * m_tiles * n_tiles;
    const int max_workers = sm_count() / CG;
    const int workers = total < max_workers ? total : max_workers;

    GemmRecord rec{};
    if (g_profile) {
        BSI_CUDA_OK(cudaEventCreate(&rec.start));
        BSI_CUDA_OK(cudaEventCreate(&rec.stop));
        rec.flops = 2.0 * p.batch * (double)p.M * (double)p.N * (double)p.K;
        BSI_CUDA_OK(cudaEventRecord(rec.start, stream));
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(workers * CG), cfg.blockDim = dim3(C::kThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes, cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    fill_pdl_attr(&attr[1]);
    cfg.attrs = attr, cfg.numAttrs = use_pdl() ? 2 : 1;
    BSI_CUDA_OK(cudaLaunchKernelEx(&cfg, k_gemm_bf16<EPI, CG, CONV, BN>, ma, ma2, mw, mc, mr, p.ep, geo, m_tiles, n_tiles, k_blocks, p.batch, p.a_shared));
    BSI_LAUNCH_OK("k_gemm_bf16");
    if (g_profile) {
        BSI_CUDA_OK(cudaEventRecord(rec.stop, stream));
        g_records.push_back(rec);
    }
    return BSI_OK;
}

template <int EPI, bool CONV>
static int dispatch_cta_group(const Problem& p, cudaStream_t stream) {
    // CTA pairs pay off as soon as there are at least two 128-row blocks; tiny problems keep the finer 128-row tiles
    const bool pair = g_force_cta_group ? g_force_cta_group == 2 : p.M > BM;
    if constexpr (CONV) {
        // the U-Net's convolutions have N = 128 outputs: a 128-wide tile halves the wasted MMA columns and epilogue work
        if (p.N <= 128) return pair ? launch_gemm<EPI, 2, CONV, 128>(p, stream) : launch_gemm<EPI, 1, CONV, 128>(p, stream);
    }
    return pair ? launch_gemm<EPI, 2, CONV, 256>(p, stream) : launch_gemm<EPI, 1, CONV, 256>(p, stream);
}

static int check_common(const Problem& p, int epilogue) {
    BSI_CHECK_ARG(p.M > 0 && p.N > 0 && p.K > 0 && p.batch >= 1, "gemm: bad shape M=%d N=%d K=%d batch=%d", p.M, p.N, p.K, p.batch);
    BSI_CHECK_ARG(p.N % (epilogue == BSI_EPI_UNPATCH_F32 ? 4 : 8) == 0, "gemm: N=%d must be a multiple of 8 (4 for UNPATCH)", p.N);
    const bool f32_out = epi_kind(epilogue) != KIND_BF16 && epi_kind(epilogue) != KIND_AUX16;
    if (epilogue != BSI_EPI_UNPATCH_F32) {
        BSI_CHECK_ARG(p.ldc >= p.N && p.ldc % (f32_out ? 4 : 8) == 0, "gemm: ldc=%d invalid for N=%d", p.ldc, p.N);
        BSI_CHECK_ARG((reinterpret_cast<uintptr_t>(p.C) & 15) == 0, "gemm: C must be 16-byte aligned");
    }
    if (epilogue == BSI_EPI_GATE_RESID_F32 || epilogue == BSI_EPI_MOD_SILU_BF16)
        BSI_CHECK_ARG(p.ep.rows_per_sample > 0 && p.ep.rows_per_sample % BM == 0, "gemm: epilogue %d needs rows_per_sample %% 128 == 0 (got %d)", epilogue,
                      p.ep.rows_per_sample);
    return BSI_OK;
}

}  // namespace bsi

using namespace bsi;

extern "C" int bsi_gemm_bf16(const bsi_gemm_args* a, void* stream) {
    BSI_CHECK_ARG(a && a->A && a->W && a->C, "bsi_gemm_bf16: null operand");
    BSI_CHECK_ARG(a->lda >= a->K && a->ldw >= a->K, "bsi_gemm_bf16: pitch smaller than K");
    BSI_CHECK_ARG((a->epilogue >= BSI_EPI_BIAS_BF16 && a->epilogue <= BSI_EPI_UNPATCH_F32) || a->epilogue == BSI_EPI_BIAS_GELU_DUAL_BF16 ||
                      a->epilogue == BSI_EPI_MUL_GELU_GRAD_BF16,
                  "bsi_gemm_bf16: unknown epilogue %d", a->epilogue);
    if (a->epilogue == BSI_EPI_BIAS_GELU_DUAL_BF16 || a->epilogue == BSI_EPI_MUL_GELU_GRAD_BF16)
        BSI_CHECK_ARG(a->aux && (reinterpret_cast<uintptr_t>(a->aux) & 15) == 0 && a->batch <= 1 && a->M > BM,
                      "bsi_gemm_bf16: epilogue %d needs a 16-byte aligned aux tensor, batch 1 and M > 128 (CTA-pair kernel)", a->epilogue);
    Problem p;
    p.A = a->A, p.W = a->W, p.C = a->C;
    p.M = a->M, p.N = a->N, p.K = a->K, p.lda = a->lda, p.ldw = a->ldw, p.ldc = a->ldc, p.batch = a->batch;
    p.stride_a = a->stride_a, p.stride_w = a->stride_w, p.stride_c = a->stride_c;
    p.a_shared = (a->batch > 1 && a->stride_a == 0) ? 1 : 0;
    p.aux = a->aux;
    EpiParams& ep = p.ep;
    ep.C = a->C, ep.bias = a->bias, ep.M = a->M, ep.N = a->N, ep.ldc = a->ldc;
    ep.stride_c = a->stride_c, ep.stride_bias = a->stride_bias;
    ep.gate = a->gate, ep.shift = bsi_rowref{}, ep.step_ptr = a->step_ptr, ep.rows_per_sample = a->rows_per_sample > 0 ? a->rows_per_sample : 1;
    ep.pos = a->pos, ep.patch = a->patch, ep.grid_w = a->grid_w, ep.channels = a->channels;
    ep.gn_partial = nullptr;
    int rc = check_common(p, a->epilogue);
    if (rc != BSI_OK) return rc;
    if (a->epilogue == BSI_EPI_GATE_RESID_F32) BSI_CHECK_ARG(a->gate.base, "bsi_gemm_bf16: GATE_RESID needs a gate");
    if (a->epilogue == BSI_EPI_POS_F32)
        BSI_CHECK_ARG(a->pos && a->rows_per_sample > 0 && (reinterpret_cast<uintptr_t>(a->pos) & 15) == 0, "bsi_gemm_bf16: POS needs a 16-byte aligned pos table");
    if (a->epilogue == BSI_EPI_UNPATCH_F32)
        BSI_CHECK_ARG(a->patch > 0 && a->grid_w > 0 && a->channels > 0 && a->rows_per_sample % a->grid_w == 0 &&
                          a->N == a->patch * a->patch * a->channels && a->batch == 1,
                      "bsi_gemm_bf16: UNPATCH geometry invalid");

    cudaStream_t st = (cudaStream_t)stream;
    switch (a->epilogue) {
        case BSI_EPI_BIAS_BF16: return dispatch_cta_group<BSI_EPI_BIAS_BF16, false>(p, st);
        case BSI_EPI_BIAS_GELU_BF16: return dispatch_cta_group<BSI_EPI_BIAS_GELU_BF16, false>(p, st);
        case BSI_EPI_BIAS_SILU_BF16: return dispatch_cta_group<BSI_EPI_BIAS_SILU_BF16, false>(p, st);
        case BSI_EPI_BIAS_F32: return dispatch_cta_group<BSI_EPI_BIAS_F32, false>(p, st);
        case BSI_EPI_GATE_RESID_F32: return dispatch_cta_group<BSI_EPI_GATE_RESID_F32, false>(p, st);
        case BSI_EPI_POS_F32: return dispatch_cta_group<BSI_EPI_POS_F32, false>(p, st);
        // training epilogues: CTA-pair kernel only (M > 128 is checked above; the single-CTA variant is not instantiated)
        case BSI_EPI_BIAS_GELU_DUAL_BF16: return launch_gemm<BSI_EPI_BIAS_GELU_DUAL_BF16, 2, false, 256>(p, st);
        case BSI_EPI_MUL_GELU_GRAD_BF16: return launch_gemm<BSI_EPI_MUL_GELU_GRAD_BF16, 2, false, 256>(p, st);
        default: return dispatch_cta_group<BSI_EPI_UNPATCH_F32, false>(p, st);
    }
}

extern "C" int bsi_conv_bf16(const bsi_conv_args* a, void* stream) {
    BSI_CHECK_ARG(a && a->X1 && a->W && a->Y, "bsi_conv_bf16: null operand");
    BSI_CHECK_ARG(a->taps == 1 || a->taps == 9, "bsi_conv_bf16: taps must be 1 (1x1) or 9 (3x3), got %d", a->taps);
    BSI_CHECK_ARG(a->B > 0 && a->H > 0 && a->Wd > 0 && a->C1 > 0 && a->C1 % 64 == 0 && (!a->X2 || (a->C2 > 0 && a->C2 % 64 == 0)),
                  "bsi_conv_bf16: bad geometry B=%d H=%d W=%d C1=%d C2=%d", a->B, a->H, a->Wd, a->C1, a->C2);
    Problem p;
    p.conv = true;
    p.A = a->X1, p.A2 = a->X2, p.W = a->W, p.C = a->Y, p.resid = a->resid;
    p.img_b = a->B, p.img_h = a->H, p.img_w = a->Wd, p.c1 = a->C1, p.c2 = a->X2 ? a->C2 : 0, p.taps = a->taps;
    p.M = a->B * a->H * a->Wd, p.N = a->N, p.K = a->taps * (p.c1 + p.c2);
    p.ldw = p.K, p.ldc = a->ldc, p.batch = 1;
    EpiParams& ep = p.ep;
    ep.C = a->Y, ep.bias = a->bias, ep.M = p.M, ep.N = p.N, ep.ldc = a->ldc, ep.stride_c = 0, ep.stride_bias = 0;
    ep.gate = a->scale, ep.shift = a->shift, ep.step_ptr = a->step_ptr, ep.rows_per_sample = a->H * a->Wd;
    ep.pos = nullptr, ep.patch = ep.grid_w = ep.channels = 0;
    ep.gn_partial = a->gn_partial;
    if (a->gn_partial)
        BSI_CHECK_ARG(a->epilogue == BSI_EPI_GATE_RESID_F32 && a->N == 128 && (a->H * a->Wd) % BM == 0,
                      "bsi_conv_bf16: GroupNorm partial sums need the GATE_RESID epilogue, N = 128 and H*W %% 128 == 0");
    int rc = check_common(p, a->epilogue);
    if (rc != BSI_OK) return rc;
    if (a->epilogue == BSI_EPI_MOD_SILU_BF16) BSI_CHECK_ARG(a->scale.base && a->shift.base, "bsi_conv_bf16: MOD_SILU needs scale and shift");
    cudaStream_t st = (cudaStream_t)stream;
    switch (a->epilogue) {
        case BSI_EPI_BIAS_BF16: return dispatch_cta_group<BSI_EPI_BIAS_BF16, true>(p, st);
        case BSI_EPI_BIAS_F32: return dispatch_cta_group<BSI_EPI_BIAS_F32, true>(p, st);
        case BSI_EPI_GATE_RESID_F32: return dispatch_cta_group<BSI_EPI_GATE_RESID_F32, true>(p, st);
        case BSI_EPI_MOD_SILU_BF16: return dispatch_cta_group<BSI_EPI_MOD_SILU_BF16, true>(p, st);
        default: set_error("bsi_conv_bf16: epilogue %d is not available for convolutions", a->epilogue); return BSI_ERR_UNSUPPORTED;
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
