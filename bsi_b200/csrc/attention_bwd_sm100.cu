// Backward of the DiT attention on tcgen05 (autograd of bsi/models/dit.py:36-47), T = 256 tokens, head dim 64, training path.
//   qkv [B*T][3*dim] bf16, dout [B*T][dim] bf16, lse2[B][H][T] (log2-sum-exp saved by the forward), dsum[B][H][T] = rowsum(dO o O)
//   ->  dqkv [B*T][3*dim] bf16
// With P = exp2(scale*log2e * Q K^T - lse2), M = dropout mask / (1 - p):
//   dV = (P o M)^T dO,   dS = P o (M o (dO V^T) - D),   dQ = scale * dS K,   dK = scale * dS^T Q.
// Everything is formed TRANSPOSED (rows = keys), so that every thread owns one key row and the per-query constants (lse2, D) are
// per-column -- there is no row reduction and no exchange between threads in this kernel at all:
//   S^T  = K Q^T  and  dP^T = V dO^T   UMMA 128x128x16, both operands K-major from smem          -> TMEM [0,128) and [128,256)
//   softmax warps: P^T (bf16) back into TMEM over columns already read; dS^T (bf16) into a 128B-swizzled smem tile
//   dV  += P^T  dO    A = P^T from TMEM,           B = dO rows (MN-major smem)                    -> TMEM [256,320)
//   dK  += dS^T Q     A = the dS^T tile, K-major,  B = Q rows  (MN-major smem)                    -> TMEM [320,384)
//   dQ  += dS   K     A = the SAME tile read MN-major (its 128-byte lines are the query axis), B = K rows (MN-major)
//                                                                                                  -> TMEM [384,448) / [448,512)
// One work item = one (sample, head); four steps (key half x query half) of 128 x 128 scores each; dV / dK leave after each key
// half, dQ at the end.  One CTA per SM (the whole TMEM), 8 softmax warps + 1 control warp (TMA loads and all MMAs).
// Replaces the two mma.sync row-owner kernels of attention_bwd.cu, which recompute S and dP twice (17 % of a training step).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace bsi {

int make_tile_map(CUtensorMap* map, const void* base, int esize, int64_t rows, int64_t cols, int64_t ld, int64_t batch, int64_t batch_stride,
                  int box_rows);

namespace abt {
constexpr int T = 256, HD = 64, HALF = 128;
constexpr int kSoftmaxWarps = 8, kThreads = 32 * (kSoftmaxWarps + 1);
constexpr int kTile = HALF * 128;  // 128 rows x 128 B = 16 KB
constexpr int kStage = 32 * 128;  // one warp's read-out tile: 32 rows x 128 B
constexpr int kSmem = 10 * kTile /*Q, K, V, dO: 2 tiles each; dS^T: 2 blocks*/ + kSoftmaxWarps * kStage + 2 * 2 * T * 4 /*lse2, D: double-buffered*/ +
                      1024 /*align*/ + 128 /*barriers*/;
constexpr uint32_t kColS = 0, kColDP = 128, kColDV = 256, kColDK = 320, kColDQ = 384;
}  // namespace abt

__device__ __forceinline__ float ex2_approx_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// D[b][h][t] = sum_c dO[b,t,h,c] * O[b,t,h,c].  A group of 8 lanes owns one (token, head): 8 x 16 B = its 128-byte line of each tensor,
// so a warp reads four consecutive heads of one token (512 contiguous bytes per tensor); 3 shuffles finish the dot product.
__global__ void __launch_bounds__(256) k_attention_dsum(float* __restrict__ dsum, const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                                                        int64_t rows, int T, int heads) {
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 3;  // (token, head) index
    const int part = threadIdx.x & 7;
    float acc = 0.0f;
    const bool live = i < rows * heads;
    if (live) {
        const uint4 a = reinterpret_cast<const uint4*>(out + i * 64)[part], g = reinterpret_cast<const uint4*>(dout + i * 64)[part];
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc = fmaf(__uint_as_float(aw[j] << 16), __uint_as_float(gw[j] << 16), acc);
            acc = fmaf(__uint_as_float(aw[j] & 0xffff0000u), __uint_as_float(gw[j] & 0xffff0000u), acc);
        }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (live && part == 0) {
        const int64_t row = i / heads;
        const int h = (int)(i - row * heads);
        const int64_t b = row / T;
        dsum[(b * heads + h) * T + (row - b * T)] = acc;
    }
}

// timing build (BSI_ATT_BWD_VARIANT=9): cycles of softmax warp 0 {wait S/dP, softmax math, wait accumulators, read-out} and of the
// control warp {wait inputs, wait P, wait read-out, wait tile release} summed over items and CTAs; [15] = items
__device__ unsigned long long g_attbwd_phase[16];

template <bool DROP, bool TIMING = false>
__global__ void __launch_bounds__(abt::kThreads, 1)
    k_attention_bwd_tc(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do, const __grid_constant__ CUtensorMap map_dqkv,
                       const float* __restrict__ lse, const float* __restrict__ dsum, const int dim, const int heads, const int total_items,
                       const float scale_log2, const float scale, const uint32_t drop_thresh, const uint32_t drop_seed, const float drop_inv) {
    using namespace abt;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                  // [256 queries][128 B], two 128-row TMA boxes
    uint8_t* sK = smem + 2 * kTile;
    uint8_t* sV = smem + 4 * kTile;
    uint8_t* sdO = smem + 6 * kTile;
    uint8_t* sDS = smem + 8 * kTile;     // dS^T of the current step: [2 blocks of 64 queries][128 keys][128 B], 128B-swizzled
    uint8_t* sStage = smem + 10 * kTile;  // [8 warps][32 rows][128 B]: read-out tiles, stored by TMA
    float* sL = reinterpret_cast<float*>(sStage + kSoftmaxWarps * kStage);  // [2][T]
    float* sD = sL + 2 * T;                                   // [2][T]
    uint64_t* bar_in = reinterpret_cast<uint64_t*>(sD + 2 * T);  // first halves of Q, K, V, dO have landed (all that step 0 needs)
    uint64_t* bar_in2 = bar_in + 7;      // second halves
    uint64_t* bar_s = bar_in + 1;        // S^T and dP^T of the step are complete
    uint64_t* bar_p = bar_in + 2;        // P^T (TMEM) and dS^T (smem) of the step are written
    uint64_t* bar_step = bar_in + 3;     // the step's accumulating MMAs have retired (P^T columns and the dS^T tile are free)
    uint64_t* bar_acc = bar_in + 4;      // dV, dK of a key half (and, the second time, dQ) are complete
    uint64_t* bar_accfree = bar_in + 5;  // ... and have been read out
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_in + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == kSoftmaxWarps) {
        if (lane == 0) {
            ptx::prefetch_tensormap(&map_qkv);
            ptx::prefetch_tensormap(&map_do);
            ptx::prefetch_tensormap(&map_dqkv);
            ptx::mbar_init(bar_in, 1);
            ptx::mbar_init(bar_in2, 1);
            ptx::mbar_init(bar_s, 1);
            ptx::mbar_init(bar_p, kSoftmaxWarps);
            ptx::mbar_init(bar_step, 1);
            ptx::mbar_init(bar_acc, 1);
            ptx::mbar_init(bar_accfree, kSoftmaxWarps);
            ptx::fence_mbar_init();
        }
        __syncwarp();
        ptx::tmem_alloc<1>(tmem_slot, 512);
        ptx::tmem_relinquish<1>();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == kSoftmaxWarps) {
        if (lane == 0) {
            // The first halves of the four operands are dead after step 2 (key half 1 x query half 0) and are exactly what the next
            // item's step 0 needs: they are refilled before step 3 runs; the second halves follow when the item's last MMA has retired.
            auto load_half = [&](int item, int half) {
                const int h = item % heads, row0 = (item / heads) * T;
                uint64_t* bar = half ? bar_in2 : bar_in;
                ptx::mbar_arrive_expect_tx(bar, 4 * kTile);
                ptx::tma_load_3d(sQ + half * kTile, &map_qkv, bar, h * HD, row0 + half * HALF, 0);
                ptx::tma_load_3d(sK + half * kTile, &map_qkv, bar, dim + h * HD, row0 + half * HALF, 0);
                ptx::tma_load_3d(sV + half * kTile, &map_qkv, bar, 2 * dim + h * HD, row0 + half * HALF, 0);
                ptx::tma_load_3d(sdO + half * kTile, &map_do, bar, h * HD, row0 + half * HALF, 0);
            };
            constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(HALF, HALF);          // S^T, dP^T: K-major x K-major
            constexpr uint32_t idesc_kv = ptx::umma_idesc_bf16(HALF, HD, 0, 1);     // dV, dK: A K-major (TMEM / smem), B MN-major
            constexpr uint32_t idesc_q = ptx::umma_idesc_bf16(HALF, HD, 1, 1);      // dQ: A MN-major (the dS^T tile), B MN-major
            const uint32_t q0 = ptx::smem_u32(sQ), k0 = ptx::smem_u32(sK), v0 = ptx::smem_u32(sV), o0 = ptx::smem_u32(sdO), ds0 = ptx::smem_u32(sDS);
            if ((int)blockIdx.x < total_items) load_half(blockIdx.x, 0), load_half(blockIdx.x, 1);
            int it = 0;
            long long cph[4] = {0, 0, 0, 0}, cc = 0;
            auto ctick = [&](int i) {
                if constexpr (TIMING) {
                    const long long now = clock64();
                    cph[i] += now - cc;
                    cc = now;
                }
            };
            if constexpr (TIMING) cc = clock64();
            for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
                const int next = item + gridDim.x;
#pragma unroll 1
                for (int step = 0; step < 4; ++step) {
                    const int kh = step >> 1, qh = step & 1;
                    ctick(3);
                    if (step == 0) ptx::mbar_wait(bar_in, it & 1);
                    if (step == 1) ptx::mbar_wait(bar_in2, it & 1);
                    ptx::tc_fence_after();
                    ctick(0);
                    {
                        const uint64_t dk = ptx::umma_desc_k_sw128(k0 + kh * kTile), dq = ptx::umma_desc_k_sw128(q0 + qh * kTile);
                        const uint64_t dv = ptx::umma_desc_k_sw128(v0 + kh * kTile), dd = ptx::umma_desc_k_sw128(o0 + qh * kTile);
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k) ptx::umma_bf16_ss<1>(tmem + kColS, dk + 2 * k, dq + 2 * k, idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k) ptx::umma_bf16_ss<1>(tmem + kColDP, dv + 2 * k, dd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                        ptx::umma_commit<1>(bar_s);
                    }
                    ptx::mbar_wait(bar_p, step & 1);
                    ctick(1);
                    if (qh == 0) ptx::mbar_wait(bar_accfree, (kh & 1) ^ 1);  // the previous key half's dV / dK (and dQ) have been read out
                    ptx::tc_fence_after();
                    ctick(2);
                    // dV[kh] += P^T dO[qh]: 16 queries per k-step = 8 packed TMEM columns; queries [0,64) at [0,32), [64,128) at [64,96)
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        ptx::umma_bf16_ts(tmem + kColDV, tmem + kColS + (k < 4 ? 8 * k : 64 + 8 * (k - 4)),
                                          ptx::umma_desc_mn_sw128(o0 + qh * kTile + k * 2048, 8192, 1024), idesc_kv, (qh | k) != 0 ? 1u : 0u);
                    // dK[kh] += dS^T Q[qh]: the dS^T tile as a K-major A operand (two 64-query blocks of four k-steps)
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        ptx::umma_bf16_ss<1>(tmem + kColDK, ptx::umma_desc_k_sw128(ds0 + (k >> 2) * kTile) + 2 * (k & 3),
                                             ptx::umma_desc_mn_sw128(q0 + qh * kTile + k * 2048, 8192, 1024), idesc_kv, (qh | k) != 0 ? 1u : 0u);
                    // dQ[qh] += dS K[kh]: the same tile read MN-major (M = 128 queries = two 64-wide blocks kTile apart; 16 keys per k-step)
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        ptx::umma_bf16_ss<1>(tmem + kColDQ + qh * HD, ptx::umma_desc_mn_sw128(ds0 + k * 2048, kTile, 1024),
                                             ptx::umma_desc_mn_sw128(k0 + kh * kTile + k * 2048, 8192, 1024), idesc_q, (kh | k) != 0 ? 1u : 0u);
                    if (qh == 1) ptx::umma_commit<1>(bar_acc);
                    // The next step's S^T / dP^T MMAs overwrite the P^T columns the MMAs above read: tcgen05.mma operations of one
                    // thread execute in issue order, so no wait is needed for that.  Refilling operand tiles is different (TMA, the
                    // async proxy): wait for the MMAs that read them.
                    if (step >= 2 && next < total_items) {
                        ptx::umma_commit<1>(bar_step);
                        ptx::mbar_wait(bar_step, step & 1);
                        load_half(next, step == 2 ? 0 : 1);
                    }
                }
            }
            if constexpr (TIMING) {
                for (int i = 0; i < 4; ++i) atomicAdd(&g_attbwd_phase[8 + i], (unsigned long long)cph[i]);
            }
        }
    } else {
        const int q = warp & 3, hf = warp >> 2;
        const int kr = q * 32 + lane;  // key row inside the half == TMEM lane
        const uint32_t trow = tmem + (static_cast<uint32_t>(q * 32) << 16);
        const uint32_t ds_row = ptx::smem_u32(sDS) + hf * kTile + kr * 128;  // this thread's 128-byte line of the block of queries [64 hf, +64)
        const int tid = threadIdx.x;
        int it = 0;
        long long sph[4] = {0, 0, 0, 0}, sc = 0;
        auto tick = [&](int i) {
            if constexpr (TIMING) {
                const long long now = clock64();
                sph[i] += now - sc;
                sc = now;
            }
        };
        if constexpr (TIMING) sc = clock64();
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
            const int h = item % heads, b = item / heads;
            const size_t row0 = (size_t)b * T;
            // per-query constants of this (sample, head), double-buffered by item parity
            float* L = sL + (it & 1) * T;
            float* D = sD + (it & 1) * T;
            L[tid] = lse[((size_t)b * heads + h) * T + tid];
            D[tid] = dsum[((size_t)b * heads + h) * T + tid];
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const uint32_t sd = DROP ? mix32(drop_seed ^ ((uint32_t)(b * heads + h) * 0x9E3779B9u)) : 0u;
#pragma unroll 1
            for (int step = 0; step < 4; ++step) {
                const int kh = step >> 1, qh = step & 1;
                const uint32_t key = (uint32_t)(kh * HALF + kr);
                ptx::mbar_wait(bar_s, step & 1);
                ptx::tc_fence_after();
                tick(0);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t s[32], dp[32];
                    ptx::tmem_ld_32x32b_x32(trow + kColS + hf * 64 + c * 32, s);
                    ptx::tmem_ld_32x32b_x32(trow + kColDP + hf * 64 + c * 32, dp);
                    ptx::tmem_ld_wait();
                    const int qc0 = qh * HALF + hf * 64 + c * 32;  // first query column of the chunk
                    uint32_t pp[16], dsp[16];
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 l4 = *reinterpret_cast<const float4*>(L + qc0 + 4 * j4), d4 = *reinterpret_cast<const float4*>(D + qc0 + 4 * j4);
                        const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
                        float pv[4], gv[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int j = 4 * j4 + i;
                            const float p = ex2_approx_ftz(fmaf(__uint_as_float(s[j]), scale_log2, -lv[i]));
                            float m = 1.0f;
                            if constexpr (DROP) m = dropout_keep(sd, (uint32_t)(qc0 + j) * T + key, drop_thresh) ? drop_inv : 0.0f;
                            gv[i] = p * fmaf(__uint_as_float(dp[j]), m, -dv[i]);  // dS^T
                            pv[i] = p * m;                                        // dropped probabilities (dV)
                        }
                        pp[2 * j4] = pack_bf16(pv[0], pv[1]), pp[2 * j4 + 1] = pack_bf16(pv[2], pv[3]);
                        dsp[2 * j4] = pack_bf16(gv[0], gv[1]), dsp[2 * j4 + 1] = pack_bf16(gv[2], gv[3]);
                    }
                    // P^T chunk: 16 packed columns over score columns this warp has already consumed
                    ptx::tmem_st_32x32b_x16(trow + kColS + hf * 64 + c * 16, pp);
                    // dS^T chunk: 32 queries = 64 B = four 16-byte chunks of this thread's line, 128B-swizzled
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        const uint32_t addr = ds_row + (((c * 4 + ch) ^ (kr & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(dsp[4 * ch]), "r"(dsp[4 * ch + 1]), "r"(dsp[4 * ch + 2]),
                                     "r"(dsp[4 * ch + 3])
                                     : "memory");
                    }
                }
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar_p);
                tick(1);

                if (qh == 1) {
                    // ---- dK (warps with hf = 0) / dV (hf = 1) of this key half, and after the last step dQ: 64 channels = one 128-byte line
                    ptx::mbar_wait(bar_acc, kh & 1);
                    ptx::tc_fence_after();
                    tick(2);
                    // 64 accumulator columns of this warp's 32 rows -> bf16 -> the warp's swizzled staging tile -> one TMA store
                    // (per-thread 16-byte global stores of 128-byte rows 6 KB apart took 2 000 clk per tile)
                    const uint32_t stage = ptx::smem_u32(sStage + warp * kStage);
                    auto store_rows = [&](uint32_t col, float sc, int gcol, int grow) {
                        uint32_t o[2][32];
                        ptx::tmem_ld_32x32b_x32(trow + col, o[0]);
                        ptx::tmem_ld_32x32b_x32(trow + col + 32, o[1]);
                        ptx::tmem_ld_wait();
                        if (lane == 0) ptx::tma_store_wait_read<0>();  // this warp's previous store has drained the staging tile
                        __syncwarp();
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint32_t* oc = o[c >> 2] + (c & 3) * 8;
                            const uint32_t addr = stage + (uint32_t)(lane * 128 + ((c ^ (lane & 7)) << 4));
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16(__uint_as_float(oc[0]) * sc, __uint_as_float(oc[1]) * sc)),
                                         "r"(pack_bf16(__uint_as_float(oc[2]) * sc, __uint_as_float(oc[3]) * sc)),
                                         "r"(pack_bf16(__uint_as_float(oc[4]) * sc, __uint_as_float(oc[5]) * sc)),
                                         "r"(pack_bf16(__uint_as_float(oc[6]) * sc, __uint_as_float(oc[7]) * sc))
                                         : "memory");
                        }
                        ptx::fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            ptx::tma_store_3d(&map_dqkv, sStage + warp * kStage, gcol, grow, 0);
                            ptx::tma_store_commit();
                        }
                    };
                    const int krow = (int)row0 + kh * HALF + q * 32;
                    if (hf == 0) store_rows(kColDK, scale, dim + h * HD, krow);
                    else store_rows(kColDV, 1.0f, 2 * dim + h * HD, krow);
                    if (kh == 1) store_rows(kColDQ + hf * HD, scale, h * HD, (int)row0 + hf * HALF + q * 32);
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(bar_accfree);
                    tick(3);
                }
            }
        }
        if (lane == 0) ptx::tma_store_wait_all<0>();  // every staged tile has reached global memory
        if constexpr (TIMING) {
            if (threadIdx.x == 0) {
                for (int i = 0; i < 4; ++i) atomicAdd(&g_attbwd_phase[i], (unsigned long long)sph[i]);
                atomicAdd(&g_attbwd_phase[15], (unsigned long long)it);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kSoftmaxWarps) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem, 512);
    }
}


// ------------------------------------------------------------------------------------------------------------------------------
// Second version: the same transposed formulation in 64-query SUB-STEPS, software-pipelined so that the tensor core and the softmax
// warps work at the same time.  A sub-step j = (key half kh, query quarter qq) forms S^T / dP^T of 128 keys x 64 queries (2 x 64 TMEM
// columns); with two such buffers the control thread issues the score MMAs of sub-step j+1 BEFORE it waits for the softmax of j, and
// the accumulating MMAs of j-1 run under the softmax of j as well:
//     issue order   S(0) S(1) | acc(0) S(2) | acc(1) S(3) | ...        (tcgen05.mma of one thread execute in issue order, so S(j+2)
//     softmax            sm(0)   sm(1)        sm(2)                     overwriting the P^T columns acc(j) reads needs no barrier)
// TMEM: S^T/dP^T 2 x (64 + 64), dV 64, dK 64, dQ 2 x 64 = 512 columns.  dQ of a 64-query quarter is an M = 64 accumulator: it lives
// in 16 lanes of each 32-lane sub-partition, so two quarters interleave in one 64-column range (lane offset 16 for the odd quarter).
// The dS^T tile (128 keys x 64 queries, one 128-byte line per key) is double-buffered in the 32 KB that held the two query blocks
// of a 128-query step before; it is still the K-major A operand of dK and the MN-major A operand (M = 64) of dQ.
namespace abt {
constexpr uint32_t kColS2 = 0, kColDP2 = 64, kBuf2 = 128;  // buffer b: S^T at b*128, dP^T at b*128 + 64
}

template <bool DROP, bool TIMING = false>
__global__ void __launch_bounds__(abt::kThreads, 1)
    k_attention_bwd_tc2(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do, const __grid_constant__ CUtensorMap map_dqkv,
                        const __grid_constant__ CUtensorMap map_dq16, const float* __restrict__ lse, const float* __restrict__ dsum, const int dim,
                        const int heads, const int total_items, const float scale_log2, const float scale, const uint32_t drop_thresh,
                        const uint32_t drop_seed, const float drop_inv) {
    using namespace abt;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                  // [256 queries][128 B], two 128-row TMA boxes
    uint8_t* sK = smem + 2 * kTile;
    uint8_t* sV = smem + 4 * kTile;
    uint8_t* sdO = smem + 6 * kTile;
    uint8_t* sDS = smem + 8 * kTile;     // dS^T: [2 buffers][128 keys][64 queries = 128 B], 128B-swizzled
    uint8_t* sStage = smem + 10 * kTile;  // [8 warps][32 rows][128 B]: read-out tiles, stored by TMA
    float* sL = reinterpret_cast<float*>(sStage + kSoftmaxWarps * kStage);  // [2][T]
    float* sD = sL + 2 * T;                                   // [2][T]
    uint64_t* bar_in = reinterpret_cast<uint64_t*>(sD + 2 * T);  // rows [0,128) of Q, K, V, dO have landed
    uint64_t* bar_in2 = bar_in + 1;      // rows [128,256)
    uint64_t* bar_s = bar_in + 2;        // [2] S^T and dP^T of a sub-step are complete (buffer = sub-step parity)
    uint64_t* bar_p = bar_in + 4;        // [2] P^T (TMEM) and dS^T (smem) of a sub-step are written
    uint64_t* bar_step = bar_in + 6;     // the MMAs issued so far have retired (operand tiles may be refilled)
    uint64_t* bar_acc = bar_in + 7;      // dV, dK of a key half (and, the second time, dQ) are complete
    uint64_t* bar_accfree = bar_in + 8;  // ... and have been read out
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_in + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == kSoftmaxWarps) {
        if (lane == 0) {
            ptx::prefetch_tensormap(&map_qkv);
            ptx::prefetch_tensormap(&map_do);
            ptx::prefetch_tensormap(&map_dqkv);
            ptx::prefetch_tensormap(&map_dq16);
            ptx::mbar_init(bar_in, 1);
            ptx::mbar_init(bar_in2, 1);
            for (int i = 0; i < 2; ++i) ptx::mbar_init(bar_s + i, 1), ptx::mbar_init(bar_p + i, kSoftmaxWarps);
            ptx::mbar_init(bar_step, 1);
            ptx::mbar_init(bar_acc, 1);
            ptx::mbar_init(bar_accfree, kSoftmaxWarps);
            ptx::fence_mbar_init();
        }
        __syncwarp();
        ptx::tmem_alloc<1>(tmem_slot, 512);
        ptx::tmem_relinquish<1>();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int my_items = (int)blockIdx.x < total_items ? (total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == kSoftmaxWarps) {
        if (lane == 0 && my_items > 0) {
            auto load_half = [&](int item, int half) {
                const int h = item % heads, row0 = (item / heads) * T;
                uint64_t* bar = half ? bar_in2 : bar_in;
                ptx::mbar_arrive_expect_tx(bar, 4 * kTile);
                ptx::tma_load_3d(sQ + half * kTile, &map_qkv, bar, h * HD, row0 + half * HALF, 0);
                ptx::tma_load_3d(sK + half * kTile, &map_qkv, bar, dim + h * HD, row0 + half * HALF, 0);
                ptx::tma_load_3d(sV + half * kTile, &map_qkv, bar, 2 * dim + h * HD, row0 + half * HALF, 0);
                ptx::tma_load_3d(sdO + half * kTile, &map_do, bar, h * HD, row0 + half * HALF, 0);
            };
            constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(HALF, 64);            // S^T, dP^T: 128 keys x 64 queries, K-major x K-major
            constexpr uint32_t idesc_kv = ptx::umma_idesc_bf16(HALF, HD, 0, 1);     // dV, dK: A K-major (TMEM / smem), B MN-major
            constexpr uint32_t idesc_q = ptx::umma_idesc_bf16(64, HD, 1, 1);        // dQ: M = 64 queries, A MN-major (the dS^T tile), B MN-major
            const uint32_t q0 = ptx::smem_u32(sQ), k0 = ptx::smem_u32(sK), v0 = ptx::smem_u32(sV), o0 = ptx::smem_u32(sdO), ds0 = ptx::smem_u32(sDS);
            load_half(blockIdx.x, 0), load_half(blockIdx.x, 1);
            // One thread issues ~190 MMAs per item and most of them take only 32-54 clk on the tensor pipe (tools/mma_probe.cu), so the
            // issue path has to be a handful of instructions per MMA: every shared-memory descriptor is a base built ONCE plus a
            // byte offset >> 4 added to its low (address) field -- all tiles sit below 256 KB, the 14-bit field cannot carry out.
            // (Rebuilding descriptors per MMA made this thread, not the tensor pipe, the bottleneck: 72 clk per MMA.)
            const uint64_t kK = ptx::umma_desc_k_sw128(k0), kQ = ptx::umma_desc_k_sw128(q0), kV = ptx::umma_desc_k_sw128(v0), kO = ptx::umma_desc_k_sw128(o0);
            const uint64_t kDS = ptx::umma_desc_k_sw128(ds0);
            const uint64_t mO = ptx::umma_desc_mn_sw128(o0, 8192, 1024), mQ = ptx::umma_desc_mn_sw128(q0, 8192, 1024);
            const uint64_t mK = ptx::umma_desc_mn_sw128(k0, 8192, 1024), mDS = ptx::umma_desc_mn_sw128(ds0, kTile, 1024);
            constexpr uint32_t kTile16 = kTile >> 4, kQuarter16 = (64 * 128) >> 4, kStep16 = 2048 >> 4;  // descriptor address units of 16 B
            long long cph[4] = {0, 0, 0, 0}, cc = 0;
            auto ctick = [&](int i) {
                if constexpr (TIMING) {
                    const long long now = clock64();
                    cph[i] += now - cc;
                    cc = now;
                }
            };
            if constexpr (TIMING) cc = clock64();
            const int total_sub = my_items * 8;
            auto issue_scores = [&](int g) {
                const int it = g >> 3, j = g & 7, kh = j >> 2, qq = j & 3;
                if (j == 0) ptx::mbar_wait(bar_in, it & 1);
                if (j == 2) ptx::mbar_wait(bar_in2, it & 1);
                ptx::tc_fence_after();
                const uint32_t buf = tmem + (g & 1) * kBuf2;
                const uint64_t dk = kK + (uint32_t)kh * kTile16, dq = kQ + (uint32_t)qq * kQuarter16;
                const uint64_t dv = kV + (uint32_t)kh * kTile16, dd = kO + (uint32_t)qq * kQuarter16;
                // S^T and dP^T interleaved: two independent accumulators
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    ptx::umma_bf16_ss<1>(buf + kColS2, dk + 2 * k, dq + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                    ptx::umma_bf16_ss<1>(buf + kColDP2, dv + 2 * k, dd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                }
                ptx::umma_commit<1>(bar_s + (g & 1));
            };
            issue_scores(0);
            uint32_t step_phase = 0;
            for (int g = 0; g < total_sub; ++g) {
                const int it = g >> 3, j = g & 7, kh = j >> 2, qq = j & 3;
                const int next = (int)blockIdx.x + (it + 1) * (int)gridDim.x;
                ctick(3);
                if (g + 1 < total_sub) issue_scores(g + 1);
                ctick(0);
                ptx::mbar_wait(bar_p + (g & 1), (g >> 1) & 1);
                ctick(1);
                if (j == 0) ptx::mbar_wait(bar_accfree, 1);  // the previous item's dV / dK / dQ have been read out
                if (j == 4) ptx::mbar_wait(bar_accfree, 0);  // this item's first dV / dK
                ptx::tc_fence_after();
                ctick(2);
                const uint32_t buf = tmem + (g & 1) * kBuf2;
                const uint64_t bO = mO + (uint32_t)qq * kQuarter16, bQ = mQ + (uint32_t)qq * kQuarter16, bK = mK + (uint32_t)kh * kTile16;
                const uint64_t aDSk = kDS + (uint32_t)(g & 1) * kTile16, aDSm = mDS + (uint32_t)(g & 1) * kTile16;
                const uint32_t dq_acc = tmem + kColDQ + (qq >> 1) * HD + ((uint32_t)(qq & 1) * 16u << 16);
                const uint32_t first = qq == 0 ? 0u : 1u, first_q = kh == 0 ? 0u : 1u;
                // dV[kh] += P^T dO[quarter]   (A = P^T from TMEM: 16 queries per k-step = 8 packed columns; queries [0,32) at [0,16), [32,64) at [32,48))
                // dK[kh] += dS^T Q[quarter]   (A = the dS^T tile, K-major: four k-steps of 16 queries)
                // dQ[quarter] += dS K[kh]     (A = the same tile MN-major, M = 64 queries = one 128-byte line; eight k-steps of 16 keys)
                // interleaved over the three accumulators
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    ptx::umma_bf16_ts(tmem + kColDV, buf + kColS2 + (k < 2 ? 8 * k : 32 + 8 * (k - 2)), bO + k * kStep16, idesc_kv, k == 0 ? first : 1u);
                    ptx::umma_bf16_ss<1>(tmem + kColDK, aDSk + 2 * k, bQ + k * kStep16, idesc_kv, k == 0 ? first : 1u);
                    ptx::umma_bf16_ss<1>(dq_acc, aDSm + (2 * k) * kStep16, bK + (2 * k) * kStep16, idesc_q, k == 0 ? first_q : 1u);
                    ptx::umma_bf16_ss<1>(dq_acc, aDSm + (2 * k + 1) * kStep16, bK + (2 * k + 1) * kStep16, idesc_q, 1u);
                }
                if (qq == 3) ptx::umma_commit<1>(bar_acc);
                // Refilling operand tiles (TMA, the async proxy) has to wait for the MMAs that read them: rows [0,128) of Q / dO and
                // K, V of key half 0 are dead after sub-step 5, the rest after sub-step 7.
                if ((j == 5 || j == 7) && next < total_items) {
                    ptx::umma_commit<1>(bar_step);
                    ptx::mbar_wait(bar_step, step_phase);
                    step_phase ^= 1;
                    load_half(next, j == 5 ? 0 : 1);
                }
            }
            if constexpr (TIMING) {
                for (int i = 0; i < 4; ++i) atomicAdd(&g_attbwd_phase[8 + i], (unsigned long long)cph[i]);
            }
        }
    } else {
        const int q = warp & 3, hf = warp >> 2;
        const int kr = q * 32 + lane;  // key row inside the half == TMEM lane
        const uint32_t trow = tmem + (static_cast<uint32_t>(q * 32) << 16);
        const int tid = threadIdx.x;
        long long sph[4] = {0, 0, 0, 0}, sc = 0;
        auto tick = [&](int i) {
            if constexpr (TIMING) {
                const long long now = clock64();
                sph[i] += now - sc;
                sc = now;
            }
        };
        if constexpr (TIMING) sc = clock64();
        for (int it = 0; it < my_items; ++it) {
            const int item = (int)blockIdx.x + it * (int)gridDim.x;
            const int h = item % heads, b = item / heads;
            const size_t row0 = (size_t)b * T;
            // per-query constants of this (sample, head), double-buffered by item parity
            float* L = sL + (it & 1) * T;
            float* D = sD + (it & 1) * T;
            L[tid] = lse[((size_t)b * heads + h) * T + tid];
            D[tid] = dsum[((size_t)b * heads + h) * T + tid];
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const uint32_t sd = DROP ? mix32(drop_seed ^ ((uint32_t)(b * heads + h) * 0x9E3779B9u)) : 0u;
#pragma unroll 1
            for (int j = 0; j < 8; ++j) {
                const int g = it * 8 + j, kh = j >> 2, qq = j & 3;
                const uint32_t key = (uint32_t)(kh * HALF + kr);
                const uint32_t buf = trow + (g & 1) * kBuf2;
                const uint32_t ds_row = ptx::smem_u32(sDS) + (g & 1) * kTile + kr * 128;  // this thread's 128-byte line of the sub-step's dS^T tile
                ptx::mbar_wait(bar_s + (g & 1), (g >> 1) & 1);
                ptx::tc_fence_after();
                tick(0);
                {
                    uint32_t s[32], dp[32];
                    ptx::tmem_ld_32x32b_x32(buf + kColS2 + hf * 32, s);
                    ptx::tmem_ld_32x32b_x32(buf + kColDP2 + hf * 32, dp);
                    ptx::tmem_ld_wait();
                    const int qc0 = qq * 64 + hf * 32;  // first query column of this thread's 32
                    uint32_t pp[16], dsp[16];
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 l4 = *reinterpret_cast<const float4*>(L + qc0 + 4 * j4), d4 = *reinterpret_cast<const float4*>(D + qc0 + 4 * j4);
                        const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
                        float pv[4], gv[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int jj = 4 * j4 + i;
                            const float p = ex2_approx_ftz(fmaf(__uint_as_float(s[jj]), scale_log2, -lv[i]));
                            float m = 1.0f;
                            if constexpr (DROP) m = dropout_keep(sd, (uint32_t)(qc0 + jj) * T + key, drop_thresh) ? drop_inv : 0.0f;
                            gv[i] = p * fmaf(__uint_as_float(dp[jj]), m, -dv[i]);  // dS^T
                            pv[i] = p * m;                                         // dropped probabilities (dV)
                        }
                        pp[2 * j4] = pack_bf16(pv[0], pv[1]), pp[2 * j4 + 1] = pack_bf16(pv[2], pv[3]);
                        dsp[2 * j4] = pack_bf16(gv[0], gv[1]), dsp[2 * j4 + 1] = pack_bf16(gv[2], gv[3]);
                    }
                    // P^T: 16 packed columns over score columns this warp itself has consumed ([0,16) for hf = 0, [32,48) for hf = 1)
                    ptx::tmem_st_32x32b_x16(buf + kColS2 + hf * 32, pp);
                    // dS^T: 32 queries = 64 B = four 16-byte chunks of this thread's line, 128B-swizzled
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        const uint32_t addr = ds_row + (((hf * 4 + ch) ^ (kr & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(dsp[4 * ch]), "r"(dsp[4 * ch + 1]), "r"(dsp[4 * ch + 2]),
                                     "r"(dsp[4 * ch + 3])
                                     : "memory");
                    }
                }
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar_p + (g & 1));
                tick(1);

                if (qq == 3) {
                    // ---- dK (warps with hf = 0) / dV (hf = 1) of this key half, and after the last sub-step dQ
                    ptx::mbar_wait(bar_acc, kh & 1);
                    ptx::tc_fence_after();
                    tick(2);
                    const uint32_t stage = ptx::smem_u32(sStage + warp * kStage);
                    // 64 accumulator columns of this warp's 32 lanes -> bf16 -> the warp's swizzled staging tile (row = lane)
                    auto stage_rows = [&](uint32_t col, float sc) {
                        uint32_t o[2][32];
                        ptx::tmem_ld_32x32b_x32(trow + col, o[0]);
                        ptx::tmem_ld_32x32b_x32(trow + col + 32, o[1]);
                        ptx::tmem_ld_wait();
                        if (lane == 0) ptx::tma_store_wait_read<0>();  // this warp's previous store has drained the staging tile
                        __syncwarp();
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint32_t* oc = o[c >> 2] + (c & 3) * 8;
                            const uint32_t addr = stage + (uint32_t)(lane * 128 + ((c ^ (lane & 7)) << 4));
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16(__uint_as_float(oc[0]) * sc, __uint_as_float(oc[1]) * sc)),
                                         "r"(pack_bf16(__uint_as_float(oc[2]) * sc, __uint_as_float(oc[3]) * sc)),
                                         "r"(pack_bf16(__uint_as_float(oc[4]) * sc, __uint_as_float(oc[5]) * sc)),
                                         "r"(pack_bf16(__uint_as_float(oc[6]) * sc, __uint_as_float(oc[7]) * sc))
                                         : "memory");
                        }
                        ptx::fence_proxy_async();
                        __syncwarp();
                    };
                    const int krow = (int)row0 + kh * HALF + q * 32;
                    stage_rows(hf == 0 ? kColDK : kColDV, hf == 0 ? scale : 1.0f);
                    if (lane == 0) {
                        ptx::tma_store_3d(&map_dqkv, sStage + warp * kStage, (hf == 0 ? dim : 2 * dim) + h * HD, krow, 0);
                        ptx::tma_store_commit();
                    }
                    if (kh == 1) {
                        // dQ: lanes [0,16) of this sub-partition hold queries 16 q + lane of quarter 2 hf, lanes [16,32) those of quarter 2 hf + 1
                        stage_rows(kColDQ + hf * HD, scale);
                        if (lane == 0) {
                            ptx::tma_store_3d(&map_dq16, sStage + warp * kStage, h * HD, (int)row0 + (2 * hf) * 64 + q * 16, 0);
                            ptx::tma_store_3d(&map_dq16, sStage + warp * kStage + 16 * 128, h * HD, (int)row0 + (2 * hf + 1) * 64 + q * 16, 0);
                            ptx::tma_store_commit();
                        }
                    }
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(bar_accfree);
                    tick(3);
                }
            }
        }
        if (lane == 0) ptx::tma_store_wait_all<0>();  // every staged tile has reached global memory
        if constexpr (TIMING) {
            if (threadIdx.x == 0) {
                for (int i = 0; i < 4; ++i) atomicAdd(&g_attbwd_phase[i], (unsigned long long)sph[i]);
                atomicAdd(&g_attbwd_phase[15], (unsigned long long)my_items);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kSoftmaxWarps) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem, 512);
    }
}

int attention_backward_tcgen05(void* dqkv_bf16, const float* lse, float* dsum, const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, int B,
                               int heads, float drop_p, uint32_t drop_seed, cudaStream_t stream) {
    using namespace abt;
    const int dim = heads * HD;
    const int64_t rows = (int64_t)B * T;
    k_attention_dsum<<<(unsigned)((rows * heads * 8 + 255) / 256), 256, 0, stream>>>(dsum, (const __nv_bfloat16*)out_bf16, (const __nv_bfloat16*)dout_bf16, rows, T,
                                                                               heads);
    BSI_LAUNCH_OK("k_attention_dsum");
    CUtensorMap mq, md;
    int rc = make_tile_map(&mq, qkv_bf16, 2, rows, 3 * dim, 3 * dim, 1, 0, HALF);
    if (rc != BSI_OK) return rc;
    rc = make_tile_map(&md, dout_bf16, 2, rows, dim, dim, 1, 0, HALF);
    if (rc != BSI_OK) return rc;
    CUtensorMap mg;
    rc = make_tile_map(&mg, dqkv_bf16, 2, rows, 3 * dim, 3 * dim, 1, 0, 32);  // read-out: one 32-row box per warp
    if (rc != BSI_OK) return rc;
    const float scale = 1.0f / sqrtf((float)HD), scale_log2 = 1.4426950408889634f * scale;
    const uint32_t thresh = dropout_thresh(drop_p);
    const float drop_inv = drop_p > 0.0f ? 1.0f / (1.0f - drop_p) : 1.0f;
    const int total = B * heads;
    const int grid = total < sm_count() ? total : sm_count();
    // BSI_ATT_BWD_VARIANT: default = the pipelined 64-query sub-step kernel; 1 = the first (128-query step) kernel; 9 = timing build
    static const int variant = [] { const char* e = getenv("BSI_ATT_BWD_VARIANT"); return e ? atoi(e) : 2; }();
    if (variant == 1) {
        if (thresh) {
            BSI_ENSURE_SMEM(k_attention_bwd_tc<true>, kSmem);
            k_attention_bwd_tc<true><<<grid, kThreads, kSmem, stream>>>(mq, md, mg, lse, dsum, dim, heads, total, scale_log2, scale, thresh, drop_seed, drop_inv);
        } else {
            BSI_ENSURE_SMEM(k_attention_bwd_tc<false>, kSmem);
            k_attention_bwd_tc<false><<<grid, kThreads, kSmem, stream>>>(mq, md, mg, lse, dsum, dim, heads, total, scale_log2, scale, 0u, 0u, 1.0f);
        }
        BSI_LAUNCH_OK("k_attention_bwd_tc");
        return BSI_OK;
    }
    CUtensorMap mg16;
    rc = make_tile_map(&mg16, dqkv_bf16, 2, rows, 3 * dim, 3 * dim, 1, 0, 16);  // dQ read-out: 16 rows of one query quarter per store
    if (rc != BSI_OK) return rc;
    if (variant == 9) {
        BSI_ENSURE_SMEM((k_attention_bwd_tc2<false, true>), kSmem);
        k_attention_bwd_tc2<false, true><<<grid, kThreads, kSmem, stream>>>(mq, md, mg, mg16, lse, dsum, dim, heads, total, scale_log2, scale, 0u, 0u, 1.0f);
    } else if (thresh) {
        BSI_ENSURE_SMEM(k_attention_bwd_tc2<true>, kSmem);
        k_attention_bwd_tc2<true><<<grid, kThreads, kSmem, stream>>>(mq, md, mg, mg16, lse, dsum, dim, heads, total, scale_log2, scale, thresh, drop_seed, drop_inv);
    } else {
        BSI_ENSURE_SMEM(k_attention_bwd_tc2<false>, kSmem);
        k_attention_bwd_tc2<false><<<grid, kThreads, kSmem, stream>>>(mq, md, mg, mg16, lse, dsum, dim, heads, total, scale_log2, scale, 0u, 0u, 1.0f);
    }
    BSI_LAUNCH_OK("k_attention_bwd_tc");
    return BSI_OK;
}

}  // namespace bsi

// Development aid: phase counters of the timing build (BSI_ATT_BWD_VARIANT=9) since the last call; resets them.
extern "C" int bsi_attention_backward_debug_phases(unsigned long long* out16) {
    BSI_CHECK_ARG(out16, "bsi_attention_backward_debug_phases: null pointer");
    unsigned long long zero[16] = {0};
    BSI_CUDA_OK(cudaMemcpyFromSymbol(out16, bsi::g_attbwd_phase, sizeof(zero)));
    BSI_CUDA_OK(cudaMemcpyToSymbol(bsi::g_attbwd_phase, zero, sizeof(zero)));
    return BSI_OK;
}
