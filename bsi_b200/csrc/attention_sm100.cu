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

#include <cstdlib>

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

// ------------------------------------------------------------------------------------------------------------------------
// Second design of the same kernel (round 2).  The first one spends 27 % of its warp samples in mbarrier waits and 14 % in
// 256-thread named barriers (ncu, profiles r01): two warps share every query row, so row max and row sum cross shared memory
// behind CTA-wide barriers, the output goes through a CTA-wide staging tile, and the whole P V product waits for the last
// probability.  Here
//   * 4 softmax warps, ONE thread per query row over all 256 keys: no max / sum exchange, no named barrier at all;
//   * P V is issued in two halves: the MMAs over keys [0,128) run while the softmax still works on keys [128,256);
//   * every warp stages its own 32 output rows and issues its own TMA store (warp-level synchronisation only);
//   * the row sum is accumulated from the fp32 probabilities (one FADD per score instead of unpack + two);
//   * optionally POLY of every 8 exponentials are evaluated on the FMA pipe (Cody-Waite split + cubic, rel. error 1.6e-4,
//     a twentieth of the bf16 step P is rounded to) to take load off the 16-per-clock SFU.
namespace att2 {
constexpr int T = 256, HD = 64, QB = 128;
constexpr int kSoftmaxWarps = 4, kThreads = 32 * (kSoftmaxWarps + 1);
constexpr int kTileBytes = QB * 128;
constexpr int kSmem = 6 * kTileBytes /*Q, K(2), V(2), 4 x 4 KB staging*/ + 1024 /*align*/ + 128 /*barriers*/;
constexpr int kTmemCols = 256;
constexpr int kOCol = 64;
}  // namespace att2

__device__ __forceinline__ float ex2_poly(float t) {
    // 2^t for t <= 0 on the FMA pipe: t = n + f, f in [-0.5, 0.5]; 2^f by a cubic; the exponent is patched in with an integer add
    t = fmaxf(t, -125.0f);
    const float r = t + 12582912.0f;  // 1.5 * 2^23: the low mantissa bits now hold round(t)
    const float f = t - (r - 12582912.0f);
    const float p = fmaf(fmaf(fmaf(0.05360212177038193f, f, 0.24237291514873505f), f, 0.6935023665428162f), f, 0.9999481439590454f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

// cycles per phase of softmax warp 0, summed over items and CTAs (BSI_ATT_VARIANT=9: timing build of the kernel)
__device__ unsigned long long g_att_phase[16];  // [0,6): softmax warp 0; [8,14): control warp

// DROP: attention dropout of the training path (F.scaled_dot_product_attention(dropout_p), dit.py:43-44): the softmax is normalised by
// the full row sum, the probabilities that enter P V carry the stateless mask of common.cuh (same mask as the backward kernels).
template <bool LSE, int POLY, bool TIMING = false, bool SPIN = false, bool DROP = false>
__global__ void __launch_bounds__(att2::kThreads, 2)
    k_attention_tc2(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_out, const int dim, const int heads,
                    const int total_items, const float scale_log2, float* __restrict__ lse, const uint32_t drop_thresh, const uint32_t drop_seed,
                    const float drop_inv) {
    using namespace att2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = smem + kTileBytes;
    uint8_t* sV = smem + 3 * kTileBytes;
    uint8_t* sO = smem + 5 * kTileBytes;  // 4 warps x (32 rows x 128 B)
    uint64_t* bar_qk = reinterpret_cast<uint64_t*>(smem + 6 * kTileBytes);
    uint64_t* bar_v = bar_qk + 1;
    uint64_t* bar_s = bar_qk + 2;
    uint64_t* bar_p0 = bar_qk + 3;  // probabilities of keys [0,128) are in TMEM
    uint64_t* bar_p1 = bar_qk + 4;  // ... of keys [128,256)
    uint64_t* bar_o = bar_qk + 5;
    uint64_t* bar_ofree = bar_qk + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 7);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto bwait = [](uint64_t* bar, uint32_t parity) {
        if constexpr (SPIN) ptx::mbar_wait_spin(bar, parity);
        else ptx::mbar_wait(bar, parity);
    };
    if (warp == kSoftmaxWarps) {
        if (lane == 0) {
            ptx::prefetch_tensormap(&map_qkv);
            ptx::prefetch_tensormap(&map_out);
            ptx::mbar_init(bar_qk, 1);
            ptx::mbar_init(bar_v, 1);
            ptx::mbar_init(bar_s, 1);
            ptx::mbar_init(bar_p0, kSoftmaxWarps);
            ptx::mbar_init(bar_p1, kSoftmaxWarps);
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
            // the V descriptor of k-step k is this base + k * (2048 >> 4) in its address field: one add per MMA on the issuing thread
            // (a 128 x 64 x 16 MMA occupies the tensor pipe for ~51 clk, tools/mma_probe.cu; the issue path must stay below that)
            const uint64_t dv0 = ptx::umma_desc_mn_sw128(ptx::smem_u32(sV), 8192, 1024);
            if ((int)blockIdx.x < total_items) {
                load_qk(blockIdx.x);
                load_v(blockIdx.x);
            }
            int it = 0;
            long long cphase[6] = {0, 0, 0, 0, 0, 0}, cc = 0;
            auto ctick = [&](int i) {
                if constexpr (TIMING) {
                    const long long now = clock64();
                    cphase[i] += now - cc;
                    cc = now;
                }
            };
            if constexpr (TIMING) cc = clock64();
            for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
                const uint32_t ph = it & 1;
                const int next = item + gridDim.x;
                bwait(bar_qk, ph);
                bwait(bar_ofree, ph ^ 1);
                ptx::tc_fence_after();
                ctick(0);  // waiting for Q/K and for O to be read out
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) ptx::umma_bf16_ss<1>(tmem, dq + 2 * k, dk + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                ptx::umma_commit<1>(bar_s);
                bwait(bar_s, ph);
                ctick(1);  // S = Q K^T issue -> completion
                if (next < total_items) load_qk(next);
                // O = P V in two halves: keys [0,128) as soon as their probabilities are written, keys [128,256) after the rest
                bwait(bar_v, ph);
                bwait(bar_p0, ph);
                ptx::tc_fence_after();
                ctick(2);  // waiting for the first half of P
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    ptx::umma_bf16_ts(tmem + kOCol, tmem + 8 * k, dv0 + k * 128, idesc_o, k != 0 ? 1u : 0u);
                bwait(bar_p1, ph);
                ptx::tc_fence_after();
                ctick(3);  // first P V half issued; waiting for the second half of P
#pragma unroll
                for (int k = 8; k < 16; ++k)
                    ptx::umma_bf16_ts(tmem + kOCol, tmem + 128 + 8 * (k - 8), dv0 + k * 128, idesc_o, 1u);
                ptx::umma_commit<1>(bar_o);
                bwait(bar_o, ph);
                ctick(4);  // second P V half issue -> completion
                if (next < total_items) load_v(next);
            }
            if constexpr (TIMING) {
                for (int i = 0; i < 5; ++i) atomicAdd(&g_att_phase[8 + i], (unsigned long long)cphase[i]);
            }
        }
    } else {
        const int r = warp * 32 + lane;  // query row inside the block == TMEM lane
        const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        uint8_t* stage = sO + warp * 4096;
        const uint32_t so = ptx::smem_u32(stage);
        int it = 0;
        long long tphase[6] = {0, 0, 0, 0, 0, 0}, tc = 0;
        auto tick = [&](int i) {
            if constexpr (TIMING) {
                const long long now = clock64();
                tphase[i] += now - tc;
                tc = now;
            }
        };
        if constexpr (TIMING) tc = clock64();
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
            const uint32_t ph = it & 1;
            int qblk, h, row0;
            coords(item, qblk, h, row0);
            bwait(bar_s, ph);
            ptx::tc_fence_after();
            tick(0);  // waiting for S

            // ---- pass 1: row max over all 256 keys (32 columns per load, next load in flight)
            uint32_t s[2][32];
            float mx = -INFINITY;
            ptx::tmem_ld_32x32b_x32(trow, s[0]);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                ptx::tmem_ld_wait();
                ptx::tmem_ld_32x32b_x32(trow + ((c + 1) & 7) * 32, s[(c + 1) & 1]);  // after the last chunk: chunk 0 again for pass 2
#pragma unroll
                for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(s[c & 1][j]));
            }
            const float moff = mx * scale_log2;
            const uint32_t drop_sd = DROP ? mix32(drop_seed ^ ((uint32_t)(item >> 1) * 0x9E3779B9u)) : 0u;  // (sample, head) stream
            const uint32_t drop_q = (uint32_t)(qblk * QB + r) * T;
            tick(1);  // pass 1

            // ---- pass 2: p = exp2(s*scale - max*scale), row sum, P -> TMEM as packed bf16 over score columns already consumed:
            //      keys [0,128) -> columns [0,64), keys [128,256) -> columns [128,192); O accumulates in [64,128)
            float sum = 0.0f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                ptx::tmem_ld_wait();
                if (c + 1 < 8) ptx::tmem_ld_32x32b_x32(trow + (c + 1) * 32, s[(c + 1) & 1]);
                uint32_t p[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float t0 = fmaf(__uint_as_float(s[c & 1][2 * j]), scale_log2, -moff);
                    const float t1 = fmaf(__uint_as_float(s[c & 1][2 * j + 1]), scale_log2, -moff);
                    // POLY of every 8 scores take the FMA-pipe exponential
                    float p0 = ((2 * j) & 7) < POLY ? ex2_poly(t0) : ex2_approx(t0);
                    float p1 = ((2 * j + 1) & 7) < POLY ? ex2_poly(t1) : ex2_approx(t1);
                    sum += p0 + p1;
                    if constexpr (DROP) {
                        const uint32_t idx = drop_q + (uint32_t)(c * 32 + 2 * j);  // query * T + key
                        bool k0, k1;  // idx is even: one hash serves both keys
                        dropout_keep_pair(drop_sd, idx, drop_thresh, k0, k1);
                        p0 = k0 ? p0 * drop_inv : 0.0f;
                        p1 = k1 ? p1 * drop_inv : 0.0f;
                    }
                    p[j] = pack_bf16(p0, p1);
                }
                ptx::tmem_st_32x32b_x16(trow + (c < 4 ? c * 16 : 128 + (c - 4) * 16), p);
                if (c == 3 || c == 7) {
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(c == 3 ? bar_p0 : bar_p1);
                }
            }

            tick(2);  // pass 2
            // ---- O / sum -> bf16 -> this warp's 32 x 128 B staging tile -> TMA store
            bwait(bar_o, ph);
            ptx::tc_fence_after();
            tick(3);  // waiting for O
            uint32_t o[2][32];
            ptx::tmem_ld_32x32b_x32(trow + kOCol, o[0]);
            ptx::tmem_ld_32x32b_x32(trow + kOCol + 32, o[1]);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                ptx::mbar_arrive(bar_ofree);           // the TMEM tile may be overwritten by the next item's S
                ptx::tma_store_wait_read<0>();         // this warp's previous store has drained its staging tile
            }
            __syncwarp();
            const float inv = 1.0f / sum;
            if constexpr (LSE) lse[((size_t)(row0 / T) * heads + h) * T + qblk * QB + r] = moff + log2f(sum);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    w[i] = pack_bf16(__uint_as_float(o[c >> 2][(c & 3) * 8 + 2 * i]) * inv, __uint_as_float(o[c >> 2][(c & 3) * 8 + 2 * i + 1]) * inv);
                const uint32_t addr = so + (uint32_t)(lane * 128 + ((c ^ (lane & 7)) << 4));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
            }
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                ptx::tma_store_3d(&map_out, stage, h * HD, row0 + qblk * QB + warp * 32, 0);
                ptx::tma_store_commit();
            }
            tick(4);  // epilogue
        }
        if (lane == 0) ptx::tma_store_wait_all<0>();
        if constexpr (TIMING) {
            if (threadIdx.x == 0) {
                for (int i = 0; i < 5; ++i) atomicAdd(&g_att_phase[i], (unsigned long long)tphase[i]);
                atomicAdd(&g_att_phase[5], (unsigned long long)it);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kSoftmaxWarps) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem, kTmemCols);
    }
}

template <bool LSE, int POLY, bool TIMING = false, bool SPIN = false, bool DROP = false>
static int launch_attention2(const CUtensorMap& mq, const CUtensorMap& mo, int dim, int heads, int total, float scale_log2, float* lse, cudaStream_t stream,
                             uint32_t drop_thresh = 0, uint32_t drop_seed = 0, float drop_inv = 1.0f) {
    using namespace att2;
    BSI_ENSURE_SMEM((k_attention_tc2<LSE, POLY, TIMING, SPIN, DROP>), kSmem);
    const int resident = 2 * sm_count();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(total < resident ? total : resident), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = kSmem, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    fill_pdl_attr(&attr[0]);
    cfg.attrs = attr, cfg.numAttrs = use_pdl() ? 1 : 0;
    BSI_CUDA_OK(cudaLaunchKernelEx(&cfg, k_attention_tc2<LSE, POLY, TIMING, SPIN, DROP>, mq, mo, dim, heads, total, scale_log2, lse, drop_thresh, drop_seed, drop_inv));
    BSI_LAUNCH_OK("k_attention_tc2");
    return BSI_OK;
}

// 0 = first design (8 softmax warps), 1..3 = second design with 0 / 2 / 4 of every 8 exponentials on the FMA pipe, 4 = second design
// with polling mbarrier waits, 5 / 9 = timing builds (bsi_attention_debug_phases)
static int attention_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("BSI_ATT_VARIANT");
        v = e ? atoi(e) : 1;  // measured on B200 at B = 256: design 1 139 us, design 2 118 us (torch SDPA 118 us); FMA-pipe exponentials
                              // (2, 3) and polling barrier waits (4) are slower or equal -- profiles/attention_r02.jsonl
        if ((v < 0 || v > 5) && v != 9) v = 0;
    }
    return v;
}

bool attention_tcgen05_has_dropout() { return attention_variant() == 1; }

int attention_tcgen05(void* out_bf16, const void* qkv_bf16, int B, int heads, float* lse, cudaStream_t stream, float drop_p, uint32_t drop_seed) {
    using namespace att;
    const int dim = heads * HD;
    CUtensorMap mq, mo;
    int rc = make_tile_map(&mq, qkv_bf16, 2, (int64_t)B * T, 3 * dim, 3 * dim, 1, 0, QB);
    if (rc != BSI_OK) return rc;
    const float scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
    const int total = B * heads * (T / QB);
    if (const int variant = attention_variant()) {
        rc = make_tile_map(&mo, out_bf16, 2, (int64_t)B * T, dim, dim, 1, 0, 32);  // one store per warp: boxes of 32 rows
        if (rc != BSI_OK) return rc;
        if (drop_p > 0.0f)
            return launch_attention2<true, 0, false, false, true>(mq, mo, dim, heads, total, scale_log2, lse, stream, dropout_thresh(drop_p), drop_seed,
                                                                  1.0f / (1.0f - drop_p));
        if (variant == 4) return lse ? launch_attention2<true, 0, false, true>(mq, mo, dim, heads, total, scale_log2, lse, stream)
                                     : launch_attention2<false, 0, false, true>(mq, mo, dim, heads, total, scale_log2, nullptr, stream);
        if (variant == 5) return launch_attention2<false, 0, true, true>(mq, mo, dim, heads, total, scale_log2, nullptr, stream);
        if (lse) {
            if (variant == 1) return launch_attention2<true, 0>(mq, mo, dim, heads, total, scale_log2, lse, stream);
            if (variant == 2) return launch_attention2<true, 2>(mq, mo, dim, heads, total, scale_log2, lse, stream);
            return launch_attention2<true, 4>(mq, mo, dim, heads, total, scale_log2, lse, stream);
        }
        if (variant == 9) return launch_attention2<false, 0, true>(mq, mo, dim, heads, total, scale_log2, nullptr, stream);
        if (variant == 1) return launch_attention2<false, 0>(mq, mo, dim, heads, total, scale_log2, nullptr, stream);
        if (variant == 2) return launch_attention2<false, 2>(mq, mo, dim, heads, total, scale_log2, nullptr, stream);
        return launch_attention2<false, 4>(mq, mo, dim, heads, total, scale_log2, nullptr, stream);
    }
    if (lse) BSI_ENSURE_SMEM(k_attention_tc<true>, kSmem);
    else BSI_ENSURE_SMEM(k_attention_tc<false>, kSmem);
    rc = make_tile_map(&mo, out_bf16, 2, (int64_t)B * T, dim, dim, 1, 0, QB);
    if (rc != BSI_OK) return rc;
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

// Development aid: cycles per phase {wait S, pass 1, pass 2, wait O, epilogue, items, -, -} of softmax warp 0 and {wait Q/K + O free,
// S latency, wait P half 1, wait P half 2, P V tail latency} of the control warp at [8, 13), accumulated by the timing build of the kernel
// (BSI_ATT_VARIANT=9) since the last call; resets the counters.
extern "C" int bsi_attention_debug_phases(unsigned long long* out6) {
    BSI_CHECK_ARG(out6, "bsi_attention_debug_phases: null pointer");
    unsigned long long zero[16] = {0};
    BSI_CUDA_OK(cudaMemcpyFromSymbol(out6, bsi::g_att_phase, 14 * sizeof(unsigned long long)));
    BSI_CUDA_OK(cudaMemcpyToSymbol(bsi::g_att_phase, zero, sizeof(zero)));
    return BSI_OK;
}
