// Fused non-causal attention for the DiT block on tcgen05 (bsi/models/dit.py:36-47), T = 256 tokens, head dim 64.
//   qkv [B*T][3*dim] bf16, columns (qkv, head, channel)  ->  out [B*T][dim] bf16, columns (head, channel)
// Persistent kernel, two CTAs resident per SM (256 TMEM columns each); a work item is one (128-query block, head,
// sample).  While one CTA is in its softmax (SFU-bound: exp2 per score) the other one runs its MMAs.
//   control warp : TMA loads of Q (128x64), K and V (256x64) into 128B-swizzled tiles, issued for item i+1 as soon as
//                  the MMAs of item i have consumed the buffers (Q,K after S; V after O), so loads never stall; issues
//                  S = Q K^T   UMMA 128x256x16 (x4, both operands K-major from smem)        -> TMEM columns [0,256)
//                  O = P V     UMMA 128x64x16 (x16, A = P from TMEM, B = V MN-major smem)  -> TMEM columns [64,128)
//   8 softmax warps: warp w owns query rows [32(w&3), +32) (its TMEM lane quarter) and keys [128(w>>2), +128); row
//                  max and row sum are combined between the two key halves through shared memory.  p = exp2((s - max) *
//                  scale*log2e) is written back as packed bf16 over S columns the same warp has already consumed:
//                  keys [0,128) -> columns [0,64), keys [128,256) -> columns [128,192); O lands in [64,128).
//                  Finally O / sum -> bf16 -> swizzled staging tile -> TMA store.
// The whole 128x256 score tile lives in TMEM: single-pass softmax statistics in fp32, no rescaling.
// Tensor-bound work 4*T*T*64 flop per (head, sample); the kernel's own ceiling is the SFU (16 ex2/clk/SM).
#include <cuda.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace bsi {

int make_tile_map(CUtensorMap* map, const void* base, int esize, int64_t rows, int64_t cols, int64_t ld, int64_t batch, int64_t batch_stride,
                  int box_rows);

namespace att {
constexpr int T = 256, HD = 64, QB = 128;
constexpr int kSoftmaxWarps = 8, kThreads = 32 * (kSoftmaxWarps + 1);  // + 1 control warp
constexpr int kTileBytes = QB * 128;                                   // 128 rows x 64 bf16
constexpr int kSmem = 6 * kTileBytes /*Q, K(2), V(2), out staging*/ + 4 * QB * 4 /*max, sum exchange*/ + 1024 /*align*/ + 128 /*barriers*/;
constexpr int kTmemCols = 256;
constexpr int kOCol = 64;  // O accumulator columns [64, 128)
}  // namespace att

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// LSE: also store log2(sum_j exp(scale * s_j)) per query row (training forward: spares the attention backward its statistics pass);
// the inference instantiation <false> is unchanged.
template <bool LSE>
__global__ void __launch_bounds__(att::kThreads, 2)
    k_attention_tc(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_out, const int dim, const int heads,
                   const int total_items, const float scale_log2, float* __restrict__ lse) {
    using namespace att;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                   // 128 x 64
    uint8_t* sK = smem + kTileBytes;      // 256 x 64 (two 128-row TMA boxes)
    uint8_t* sV = smem + 3 * kTileBytes;  // 256 x 64
    uint8_t* sO = smem + 5 * kTileBytes;  // output staging tile
    float* s_max = reinterpret_cast<float*>(smem + 6 * kTileBytes);  // [2 halves][128 rows]
    float* s_sum = s_max + 2 * QB;
    uint64_t* bar_qk = reinterpret_cast<uint64_t*>(s_sum + 2 * QB);
    uint64_t* bar_v = bar_qk + 1;
    uint64_t* bar_s = bar_qk + 2;      // S complete (tcgen05.commit): softmax may start, Q/K tiles may be refilled
    uint64_t* bar_p = bar_qk + 3;      // P written by the softmax warps
    uint64_t* bar_o = bar_qk + 4;      // O complete (tcgen05.commit): epilogue may start, V tile may be refilled
    uint64_t* bar_ofree = bar_qk + 5;  // O (and with it the whole TMEM tile) read out by the softmax warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == kSoftmaxWarps) {
        if (lane == 0) {
            ptx::prefetch_tensormap(&map_qkv);
            ptx::prefetch_tensormap(&map_out);
            ptx::mbar_init(bar_qk, 1);
            ptx::mbar_init(bar_v, 1);
            ptx::mbar_init(bar_s, 1);
            ptx::mbar_init(bar_p, kSoftmaxWarps);
            ptx::mbar_init(bar_o, 1);
            ptx::mbar_init(bar_ofree, kSoftmaxWarps);
            ptx::fence_mbar_init();
        }
        __syncwarp();
        ptx::tmem_alloc<1>(tmem_slot, kTmemCols);
        ptx::tmem_relinquish<1>();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_prologue_done();

    // work item -> (query block, head, sample); consecutive items share K/V in L2
    auto coords = [&](int item, int& qblk, int& h, int& row0) {
        qblk = item & 1;
        const int bh = item >> 1;
        h = bh % heads;
        row0 = (bh / heads) * T;
    };

    if (warp == kSoftmaxWarps) {
        if (lane == 0) {
            auto load_qk = [&](int item) {
                int qblk, h, row0;
                coords(item, qblk, h, row0);
                ptx::mbar_arrive_expect_tx(bar_qk, 3 * kTileBytes);
                ptx::tma_load_3d(sQ, &map_qkv, bar_qk, h * HD, row0 + qblk * QB, 0);
                ptx::tma_load_3d(sK, &map_qkv, bar_qk, dim + h * HD, row0, 0);
                ptx::tma_load_3d(sK + kTileBytes, &map_qkv, bar_qk, dim + h * HD, row0 + QB, 0);
            };
            auto load_v = [&](int item) {
                int qblk, h, row0;
                coords(item, qblk, h, row0);
                ptx::mbar_arrive_expect_tx(bar_v, 2 * kTileBytes);
                ptx::tma_load_3d(sV, &map_qkv, bar_v, 2 * dim + h * HD, row0, 0);
                ptx::tma_load_3d(sV + kTileBytes, &map_qkv, bar_v, 2 * dim + h * HD, row0 + QB, 0);
            };
            constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(QB, T);
            constexpr uint32_t idesc_o = ptx::umma_idesc_bf16(QB, HD, 0, 1);
            const uint64_t dq = ptx::umma_desc_k_sw128(ptx::smem_u32(sQ)), dk = ptx::umma_desc_k_sw128(ptx::smem_u32(sK));
            const uint32_t v0 = ptx::smem_u32(sV);

            if ((int)blockIdx.x < total_items) {
                load_qk(blockIdx.x);
                load_v(blockIdx.x);
            }
            int it = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
                const uint32_t ph = it & 1;
                const int next = item + gridDim.x;
                // ---- S = Q K^T : M = 128 queries, N = 256 keys, K = 64 channels (needs the previous item's O read out)
                ptx::mbar_wait(bar_qk, ph);
                ptx::mbar_wait(bar_ofree, ph ^ 1);
                ptx::tc_fence_after();
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) ptx::umma_bf16_ss<1>(tmem, dq + 2 * k, dk + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                ptx::umma_commit<1>(bar_s);
                // Q and K are free once S has been computed: prefetch the next item behind this item's softmax
                ptx::mbar_wait(bar_s, ph);
                if (next < total_items) load_qk(next);
                // ---- O = P V : M = 128 queries, N = 64 channels, K = 256 keys; A = P (bf16, 8 TMEM columns per 16 keys: keys
                //      [0,128) at columns [0,64), keys [128,256) at columns [128,192)), B = V as stored (key rows of 64
                //      contiguous channels = MN-major, 16 keys = 2048 B per k-step)
                ptx::mbar_wait(bar_p, ph);
                ptx::mbar_wait(bar_v, ph);
                ptx::tc_fence_after();
#pragma unroll
                for (int k = 0; k < T / 16; ++k) {
                    const uint64_t dv = ptx::umma_desc_mn_sw128(v0 + k * 2048, 8192, 1024);
                    const uint32_t pa = tmem + (k < 8 ? 8 * k : 128 + 8 * (k - 8));
                    ptx::umma_bf16_ts(tmem + kOCol, pa, dv, idesc_o, k != 0 ? 1u : 0u);
                }
                ptx::umma_commit<1>(bar_o);
                ptx::mbar_wait(bar_o, ph);
                if (next < total_items) load_v(next);
            }
        }
    } else {
        const int q = warp & 3, hf = warp >> 2;
        const int r = q * 32 + lane;  // query row inside the block == TMEM lane
        const uint32_t trow = tmem + (static_cast<uint32_t>(q * 32) << 16);
        const uint32_t scol = trow + hf * 128;  // this warp's 128 score columns (and, in their first half, its P columns)
        const uint32_t so = ptx::smem_u32(sO);
        int it = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
            const uint32_t ph = it & 1;
            int qblk, h, row0;
            coords(item, qblk, h, row0);
            ptx::mbar_wait(bar_s, ph);
            ptx::tc_fence_after();

            // ---- pass 1: partial row max over this warp's 128 keys (32 columns per load, next load in flight)
            uint32_t s[2][32];
            float mx = -INFINITY;
            ptx::tmem_ld_32x32b_x32(scol, s[0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                ptx::tmem_ld_wait();
                ptx::tmem_ld_32x32b_x32(scol + ((c + 1) & 3) * 32, s[(c + 1) & 1]);  // after the last chunk: chunk 0 again for pass 2
#pragma unroll
                for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(s[c & 1][j]));
            }
            s_max[hf * QB + r] = mx;
            softmax_bar();
            const float moff = fmaxf(mx, s_max[(hf ^ 1) * QB + r]) * scale_log2;

            // ---- pass 2: p = exp2(s*scale - max*scale), partial row sum, P -> TMEM as packed bf16
            float sum = 0.0f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                ptx::tmem_ld_wait();  // chunk c sits in s[c & 1]
                if (c + 1 < 4) ptx::tmem_ld_32x32b_x32(scol + (c + 1) * 32, s[(c + 1) & 1]);
                uint32_t p[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float p0 = ex2_approx(fmaf(__uint_as_float(s[c & 1][2 * j]), scale_log2, -moff));
                    const float p1 = ex2_approx(fmaf(__uint_as_float(s[c & 1][2 * j + 1]), scale_log2, -moff));
                    const uint32_t pk = pack_bf16(p0, p1);
                    // the row sum uses the bf16-rounded probabilities that the PV product will see
                    sum += __uint_as_float(pk << 16) + __uint_as_float(pk & 0xffff0000u);
                    p[j] = pk;
                }
                // P chunk c (32 keys = 16 packed columns) at [16c, 16c+16) of this warp's range: score columns it has already read
                ptx::tmem_st_32x32b_x16(scol + c * 16, p);
            }
            s_sum[hf * QB + r] = sum;
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_p);

            // ---- O / sum -> bf16 -> staging tile -> TMA store; this warp normalises channels [32hf, 32hf+32)
            ptx::mbar_wait(bar_o, ph);
            ptx::tc_fence_after();
            uint32_t o[32];
            ptx::tmem_ld_32x32b_x32(trow + kOCol + hf * 32, o);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_ofree);  // the TMEM tile may be overwritten by the next item's S
            if (threadIdx.x == 0) ptx::tma_store_wait_read<0>();  // previous item's store has drained the staging tile
            softmax_bar();  // also orders the s_sum exchange (written before bar_p above)
            const float total = sum + s_sum[(hf ^ 1) * QB + r];
            const float inv = 1.0f / total;
            if constexpr (LSE) {
                if (hf == 0) lse[((size_t)(row0 / T) * heads + h) * T + qblk * QB + r] = moff + log2f(total);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) w[i] = pack_bf16(__uint_as_float(o[c * 8 + 2 * i]) * inv, __uint_as_float(o[c * 8 + 2 * i + 1]) * inv);
                const uint32_t addr = so + (uint32_t)(r * 128 + (((hf * 4 + c) ^ (r & 7)) << 4));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
            }
            ptx::fence_proxy_async();
            softmax_bar();
            if (threadIdx.x == 0) {
                ptx::tma_store_3d(&map_out, sO, h * HD, row0 + qblk * QB, 0);
                ptx::tma_store_commit();
            }
        }
        if (threadIdx.x == 0) ptx::tma_store_wait_all<0>();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kSoftmaxWarps) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem, kTmemCols);
    }
}

int attention_tcgen05(void* out_bf16, const void* qkv_bf16, int B, int heads, float* lse, cudaStream_t stream) {
    using namespace att;
    const int dim = heads * HD;
    if (lse) BSI_ENSURE_SMEM(k_attention_tc<true>, kSmem);
    else BSI_ENSURE_SMEM(k_attention_tc<false>, kSmem);
    CUtensorMap mq, mo;
    int rc = make_tile_map(&mq, qkv_bf16, 2, (int64_t)B * T, 3 * dim, 3 * dim, 1, 0, QB);
    if (rc != BSI_OK) return rc;
    rc = make_tile_map(&mo, out_bf16, 2, (int64_t)B * T, dim, dim, 1, 0, QB);
    if (rc != BSI_OK) return rc;
    const float scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
    const int total = B * heads * (T / QB);
    const int resident = 2 * sm_count();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(total < resident ? total : resident), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = kSmem, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    fill_pdl_attr(&attr[0]);
    cfg.attrs = attr, cfg.numAttrs = use_pdl() ? 1 : 0;
    if (lse) BSI_CUDA_OK(cudaLaunchKernelEx(&cfg, k_attention_tc<true>, mq, mo, dim, heads, total, scale_log2, lse));
    else BSI_CUDA_OK(cudaLaunchKernelEx(&cfg, k_attention_tc<false>, mq, mo, dim, heads, total, scale_log2, (float*)nullptr));
    BSI_LAUNCH_OK("k_attention_tc");
    return BSI_OK;
}

}  // namespace bsi
