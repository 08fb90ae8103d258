// Non-causal multi-head attention over packed QKV for the DiT block (bsi/models/dit.py:36-47):
//   qkv [B*T][3*dim] bf16, columns (qkv, head, channel)  ->  out [B*T][dim] bf16, columns (head, channel)
// which removes the reference's two permute+contiguous copies (dit.py:39-41,46).
// One CTA per (128 query rows, head, sample); K and V of the head (T x 64) stay in shared memory,
// scores are kept in registers with an online softmax over 64-key chunks (fp32 statistics).
// Tensor-bound: 4*T*T*64 flop per (head, sample).  v1 uses warp-level mma.sync m16n8k16.
#include "common.cuh"

namespace bsi {

constexpr int kAttThreads = 256;  // 8 warps x 16 query rows
constexpr int kHd = 64;           // head dim
constexpr int kQRows = 128;

__device__ __forceinline__ uint32_t sw_off(int row, int chunk) { return (uint32_t)(row * 128 + (((chunk) ^ (row & 7)) << 4)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// DROP: attention dropout of the training path (F.scaled_dot_product_attention(dropout_p), dit.py:43-44): the softmax is normalised
// over all keys, then each probability is kept with probability 1 - p and rescaled; the mask is a pure function of
// (seed, head-of-sample, query, key) so that the backward kernels (attention_bwd.cu) regenerate it.
template <bool DROP>
__global__ void __launch_bounds__(kAttThreads, 2)
    k_attention_mma(__nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ qkv, int T, int dim, float scale_log2, uint32_t drop_thresh,
                    uint32_t drop_seed, float drop_inv, float* __restrict__ lse) {
    extern __shared__ __align__(128) uint8_t att_smem[];
    const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(att_smem), sK = sQ + kQRows * 128, sV = sK + T * 128;
    const int q0 = blockIdx.x * kQRows, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t ld = 3 * (size_t)dim;
    const __nv_bfloat16* base = qkv + (size_t)b * T * ld + h * kHd;

    // ---- stage Q (128 x 64), K and V (T x 64) with 16-byte cp.async into swizzled rows
    for (int i = tid; i < kQRows * 8; i += kAttThreads) {
        int r = i >> 3, c = i & 7;
        cp_async16(sQ + sw_off(r, c), base + (size_t)(q0 + r) * ld + c * 8);
    }
    for (int i = tid; i < T * 8; i += kAttThreads) {
        int r = i >> 3, c = i & 7;
        cp_async16(sK + sw_off(r, c), base + (size_t)r * ld + dim + c * 8);
        cp_async16(sV + sw_off(r, c), base + (size_t)r * ld + 2 * dim + c * 8);
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int r0 = warp * 16;
    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8, chunk = ks * 2 + (lane >> 4);
        ldsm_x4(sQ + sw_off(row, chunk), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }

    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.0f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.0f, 0.0f};

    for (int kc = 0; kc < T / 64; ++kc) {
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                int key = kc * 64 + np * 16 + (lane & 7) + (lane >> 4) * 8, chunk = ks * 2 + ((lane >> 3) & 1);
                uint32_t b0, b1, b2, b3;
                ldsm_x4(sK + sw_off(key, chunk), b0, b1, b2, b3);
                mma_bf16(s[2 * np], qf[ks], b0, b1);
                mma_bf16(s[2 * np + 1], qf[ks], b2, b3);
            }
        }
        // online softmax; thread holds rows g (=lane/4) [c0,c1] and g+8 [c2,c3]
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
        float corr[2], msc[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float m_new = fmaxf(m_run[r], mx[r]);
            corr[r] = exp2f((m_run[r] - m_new) * scale_log2);
            m_run[r] = m_new;
            msc[r] = m_new * scale_log2;
        }
        float rs[2] = {0.0f, 0.0f};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = exp2f(fmaf(s[j][0], scale_log2, -msc[0]));
            s[j][1] = exp2f(fmaf(s[j][1], scale_log2, -msc[0]));
            s[j][2] = exp2f(fmaf(s[j][2], scale_log2, -msc[1]));
            s[j][3] = exp2f(fmaf(s[j][3], scale_log2, -msc[1]));
            rs[0] += s[j][0] + s[j][1];
            rs[1] += s[j][2] + s[j][3];
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
        if constexpr (DROP) {
            const uint32_t sd = mix32(drop_seed ^ ((uint32_t)(b * gridDim.y + h) * 0x9E3779B9u));
            const uint32_t qa = (uint32_t)(q0 + warp * 16 + (lane >> 2)) * T, qb = qa + 8u * T;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t key = kc * 64 + j * 8 + 2 * (lane & 3);
                s[j][0] = dropout_keep(sd, qa + key, drop_thresh) ? s[j][0] * drop_inv : 0.0f;
                s[j][1] = dropout_keep(sd, qa + key + 1, drop_thresh) ? s[j][1] * drop_inv : 0.0f;
                s[j][2] = dropout_keep(sd, qb + key, drop_thresh) ? s[j][2] * drop_inv : 0.0f;
                s[j][3] = dropout_keep(sd, qb + key + 1, drop_thresh) ? s[j][3] * drop_inv : 0.0f;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[j][0] *= corr[0], o[j][1] *= corr[0];
            o[j][2] *= corr[1], o[j][3] *= corr[1];
        }
        // O += P V
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t pa[4] = {pack_bf16(s[2 * kk][0], s[2 * kk][1]), pack_bf16(s[2 * kk][2], s[2 * kk][3]),
                              pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
            for (int dn = 0; dn < 4; ++dn) {
                int key = kc * 64 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, chunk = dn * 2 + (lane >> 4);
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(sV + sw_off(key, chunk), b0, b1, b2, b3);
                mma_bf16(o[2 * dn], pa, b0, b1);
                mma_bf16(o[2 * dn + 1], pa, b2, b3);
            }
        }
    }
    // finish: row sums across the 4 lanes of a quad, normalise, stage through this warp's Q rows, store 16 B per lane
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
    __syncwarp();
    const int g = lane >> 2, t4 = lane & 3;
    if (lse && t4 == 0) {  // log2 of sum_j exp(scale * s_j) per query row: lets the backward skip its statistics pass
        float* lrow = lse + ((size_t)b * gridDim.y + h) * T + q0 + r0;
        lrow[g] = fmaf(m_run[0], scale_log2, log2f(l_run[0]));
        lrow[g + 8] = fmaf(m_run[1], scale_log2, log2f(l_run[1]));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        // column pair (j*8 + 2*t4, +1) lives in 16-byte chunk j at byte offset 4*t4
        uint32_t lo = pack_bf16(o[j][0] * inv0, o[j][1] * inv0), hi = pack_bf16(o[j][2] * inv1, o[j][3] * inv1);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sQ + sw_off(r0 + g, j) + 4 * t4), "r"(lo) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sQ + sw_off(r0 + g + 8, j) + 4 * t4), "r"(hi) : "memory");
    }
    __syncwarp();
    __nv_bfloat16* obase = out + ((size_t)b * T + q0) * dim + h * kHd;
#pragma unroll
    for (int i = lane; i < 16 * 8; i += 32) {
        int r = r0 + (i >> 3), c = i & 7;
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sQ + sw_off(r, c)));
        *reinterpret_cast<uint4*>(obase + (size_t)r * dim + c * 8) = v;
    }
}

int attention_tcgen05(void* out_bf16, const void* qkv_bf16, int B, int heads, float* lse, cudaStream_t stream, float drop_p = 0.0f,
                      uint32_t drop_seed = 0);  // attention_sm100.cu
bool attention_tcgen05_has_dropout();  // the second design of the kernel is selected (the first one has no dropout variant)
static int g_force_legacy_attention = 0;

}  // namespace bsi

using namespace bsi;

// Test hook: 1 forces the warp-level mma.sync kernel even where the tcgen05 kernel applies, 0 restores the default.
extern "C" int bsi_attention_force_legacy(int32_t on) {
    g_force_legacy_attention = on ? 1 : 0;
    return BSI_OK;
}

extern "C" int bsi_attention_bf16(void* out_bf16, const void* qkv_bf16, int32_t B, int32_t T, int32_t heads, int32_t head_dim,
                                  void* stream) {
    BSI_CHECK_ARG(out_bf16 && qkv_bf16 && B > 0 && heads > 0, "bsi_attention_bf16: bad arguments");
    if (head_dim != kHd || T % kQRows != 0 || T > 512) {
        set_error("bsi_attention_bf16: only head_dim=64 and T in {128,256,384,512} are implemented (got head_dim=%d T=%d)", head_dim, T);
        return BSI_ERR_UNSUPPORTED;
    }
    if (T == 256 && !g_force_legacy_attention) return attention_tcgen05(out_bf16, qkv_bf16, B, heads, nullptr, (cudaStream_t)stream);
    const int dim = heads * head_dim;
    const int smem = (kQRows + 2 * T) * 128;
    BSI_ENSURE_SMEM(k_attention_mma<false>, smem);
    const float scale_log2 = 1.4426950408889634f / sqrtf((float)head_dim);
    dim3 grid(T / kQRows, heads, B);
    k_attention_mma<false><<<grid, kAttThreads, smem, (cudaStream_t)stream>>>((__nv_bfloat16*)out_bf16, (const __nv_bfloat16*)qkv_bf16, T, dim,
                                                                             scale_log2, 0u, 0u, 1.0f, nullptr);
    BSI_LAUNCH_OK("k_attention_mma");
    return BSI_OK;
}

// Training forward without dropout: the attention output plus the per-row log2-sum-exp the backward needs (lse_valid fast path).
extern "C" int bsi_attention_lse_bf16(void* out_bf16, float* lse_out, const void* qkv_bf16, int32_t B, int32_t T, int32_t heads, int32_t head_dim,
                                      void* stream) {
    BSI_CHECK_ARG(out_bf16 && lse_out && qkv_bf16 && B > 0 && heads > 0, "bsi_attention_lse_bf16: bad arguments");
    if (head_dim != kHd || T % kQRows != 0 || T > 512) {
        set_error("bsi_attention_lse_bf16: only head_dim=64 and T in {128,256,384,512} are implemented (got head_dim=%d T=%d)", head_dim, T);
        return BSI_ERR_UNSUPPORTED;
    }
    if (T == 256 && !g_force_legacy_attention) return attention_tcgen05(out_bf16, qkv_bf16, B, heads, lse_out, (cudaStream_t)stream);
    const int dim = heads * head_dim;
    const int smem = (kQRows + 2 * T) * 128;
    BSI_ENSURE_SMEM(k_attention_mma<false>, smem);
    const float scale_log2 = 1.4426950408889634f / sqrtf((float)head_dim);
    dim3 grid(T / kQRows, heads, B);
    k_attention_mma<false><<<grid, kAttThreads, smem, (cudaStream_t)stream>>>((__nv_bfloat16*)out_bf16, (const __nv_bfloat16*)qkv_bf16, T, dim, scale_log2,
                                                                             0u, 0u, 1.0f, lse_out);
    BSI_LAUNCH_OK("k_attention_mma");
    return BSI_OK;
}

extern "C" int bsi_attention_dropout_bf16(void* out_bf16, float* lse_out, const void* qkv_bf16, int32_t B, int32_t T, int32_t heads, int32_t head_dim,
                                          float drop_p, uint32_t drop_seed, void* stream) {
    BSI_CHECK_ARG(out_bf16 && qkv_bf16 && B > 0 && heads > 0 && drop_p > 0.0f && drop_p < 1.0f, "bsi_attention_dropout_bf16: bad arguments");
    if (head_dim != kHd || T % kQRows != 0 || T > 512) {
        set_error("bsi_attention_dropout_bf16: only head_dim=64 and T in {128,256,384,512} are implemented (got head_dim=%d T=%d)", head_dim, T);
        return BSI_ERR_UNSUPPORTED;
    }
    if (T == 256 && !g_force_legacy_attention && lse_out && attention_tcgen05_has_dropout())
        return attention_tcgen05(out_bf16, qkv_bf16, B, heads, lse_out, (cudaStream_t)stream, drop_p, drop_seed);
    const int dim = heads * head_dim;
    const int smem = (kQRows + 2 * T) * 128;
    BSI_ENSURE_SMEM(k_attention_mma<true>, smem);
    const float scale_log2 = 1.4426950408889634f / sqrtf((float)head_dim);
    dim3 grid(T / kQRows, heads, B);
    k_attention_mma<true><<<grid, kAttThreads, smem, (cudaStream_t)stream>>>((__nv_bfloat16*)out_bf16, (const __nv_bfloat16*)qkv_bf16, T, dim, scale_log2,
                                                                            dropout_thresh(drop_p), drop_seed, 1.0f / (1.0f - drop_p), lse_out);
    BSI_LAUNCH_OK("k_attention_mma<dropout>");
    return BSI_OK;
}
