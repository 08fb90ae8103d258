// Backward of the DiT attention (autograd of bsi/models/dit.py:36-47) on the packed layout, for the training path:
//   qkv [B*T][3*dim] bf16, out [B*T][dim] bf16 (saved forward output), dout [B*T][dim] bf16  ->  dqkv [B*T][3*dim] bf16
// With P = softmax(Q K^T / sqrt(d)), D_i = sum_c dO_ic O_ic:   dV = P^T dO,   dS = P o (dO V^T - D),   dQ = dS K / sqrt(d),   dK = dS^T Q / sqrt(d).
// Two row-owner kernels (warp-level mma.sync m16n8k16, fp32 accumulation), so that no cross-warp reduction or atomics are needed
// and the result is deterministic:
//   k_attention_bwd_q   CTA = 128 query rows of one (head, sample): pass 1 recomputes the softmax statistics (log-sum-exp, written to
//                       lse[B][H][T] together with D), pass 2 recomputes P per 64-key chunk, forms dS and accumulates dQ += dS K.
//   k_attention_bwd_kv  CTA = 128 key rows: per 64-query chunk S^T = K Q^T, P^T, dP^T = V dO^T, dS^T, and dV += P^T dO, dK += dS^T Q.
// S and dP are computed twice (8 instead of 5 products); the library flash-attention backward this replaces also recomputes S.
#include <cstdlib>

#include "common.cuh"

namespace bsi {

namespace ab {
constexpr int kThreads = 256;  // 8 warps x 16 rows
constexpr int kHd = 64, kRows = 128;

__device__ __forceinline__ uint32_t sw_off(int row, int chunk) { return (uint32_t)(row * 128 + (((chunk) ^ (row & 7)) << 4)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// stage `rows` rows of 64 bf16 (128 B) from a row-major matrix with pitch ld into a swizzled smem tile
__device__ __forceinline__ void stage(uint32_t dst, const __nv_bfloat16* src, size_t ld, int rows) {
    for (int i = threadIdx.x; i < rows * 8; i += kThreads) {
        const int r = i >> 3, c = i & 7;
        cp_async16(dst + sw_off(r, c), src + (size_t)r * ld + c * 8);
    }
}
// A-operand fragments (16 rows x 64) of the warp's rows r0.. of a staged tile
__device__ __forceinline__ void load_frags(uint32_t (&f)[4][4], uint32_t tile, int r0, int lane) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8, chunk = ks * 2 + (lane >> 4);
        ldsm_x4(tile + sw_off(row, chunk), f[ks][0], f[ks][1], f[ks][2], f[ks][3]);
    }
}
// s[16 x 64] = A[16 x 64] . B[col0 .. col0+64)[64]^T   (B rows are the output columns)
__device__ __forceinline__ void mm_abt(float (&s)[8][4], const uint32_t (&a)[4][4], uint32_t tileB, int col0, int lane) {
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            const int key = col0 + np * 16 + (lane & 7) + (lane >> 4) * 8, chunk = ks * 2 + ((lane >> 3) & 1);
            uint32_t b0, b1, b2, b3;
            ldsm_x4(tileB + sw_off(key, chunk), b0, b1, b2, b3);
            mma_bf16(s[2 * np], a[ks], b0, b1);
            mma_bf16(s[2 * np + 1], a[ks], b2, b3);
        }
    }
}
// acc[16 x 64] += P[16 x 64] . B[row0 .. row0+64)[64]   (P from registers, bf16-rounded; B rows are the contraction)
__device__ __forceinline__ void mm_pb(float (&acc)[8][4], const float (&p)[8][4], uint32_t tileB, int row0, int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const uint32_t pa[4] = {pack_bf16(p[2 * kk][0], p[2 * kk][1]), pack_bf16(p[2 * kk][2], p[2 * kk][3]),
                                pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]), pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3])};
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
            const int key = row0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, chunk = dn * 2 + (lane >> 4);
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(tileB + sw_off(key, chunk), b0, b1, b2, b3);
            mma_bf16(acc[2 * dn], pa, b0, b1);
            mma_bf16(acc[2 * dn + 1], pa, b2, b3);
        }
    }
}
// acc (16 rows x 64, scaled) -> bf16 -> staged through the warp's own rows of `tile` -> global rows with pitch ld
__device__ __forceinline__ void store_rows(const float (&acc)[8][4], float scale, uint32_t tile, int r0, int lane, __nv_bfloat16* dst, size_t ld) {
    __syncwarp();
    const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t lo = pack_bf16(acc[j][0] * scale, acc[j][1] * scale), hi = pack_bf16(acc[j][2] * scale, acc[j][3] * scale);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile + sw_off(r0 + g, j) + 4 * t4), "r"(lo) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile + sw_off(r0 + g + 8, j) + 4 * t4), "r"(hi) : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int i = lane; i < 16 * 8; i += 32) {
        const int r = r0 + (i >> 3), c = i & 7;
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(tile + sw_off(r, c)));
        *reinterpret_cast<uint4*>(dst + (size_t)r * ld + c * 8) = v;
    }
    __syncwarp();
}
}  // namespace ab

__global__ void __launch_bounds__(ab::kThreads, 1)
    k_attention_bwd_q(__nv_bfloat16* __restrict__ dqkv, float* __restrict__ lse, float* __restrict__ dsum, const __nv_bfloat16* __restrict__ qkv,
                      const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout, int T, int dim, int heads, float scale_log2,
                      float scale, uint32_t drop_thresh, uint32_t drop_seed, float drop_inv, int have_lse) {
    using namespace ab;
    extern __shared__ __align__(128) uint8_t ab_smem[];
    const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(ab_smem), sdO = sQ + kRows * 128, sO = sdO + kRows * 128, sK = sO + kRows * 128,
                   sV = sK + T * 128;
    float* sD = reinterpret_cast<float*>(ab_smem + 3 * kRows * 128 + 2 * T * 128);
    const int q0 = blockIdx.x * kRows, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t ld = 3 * (size_t)dim;
    const __nv_bfloat16* base = qkv + (size_t)b * T * ld + h * kHd;
    stage(sQ, base + (size_t)q0 * ld, ld, kRows);
    stage(sdO, dout + ((size_t)b * T + q0) * dim + h * kHd, dim, kRows);
    stage(sO, out + ((size_t)b * T + q0) * dim + h * kHd, dim, kRows);
    stage(sK, base + dim, ld, T);
    stage(sV, base + 2 * dim, ld, T);
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    {  // D_i = sum_c dO_ic * O_ic: two threads per row, 32 channels each
        const int row = tid >> 1, half = tid & 1;
        float d = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint4 a, o;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(sdO + sw_off(row, half * 4 + c)));
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w) : "r"(sO + sw_off(row, half * 4 + c)));
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                d = fmaf(__uint_as_float(aw[i] << 16), __uint_as_float(ow[i] << 16), d);
                d = fmaf(__uint_as_float(aw[i] & 0xffff0000u), __uint_as_float(ow[i] & 0xffff0000u), d);
            }
        }
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        if (half == 0) {
            sD[row] = d;
            dsum[((size_t)b * heads + h) * T + q0 + row] = d;
        }
    }
    __syncthreads();

    const int r0 = warp * 16, g = lane >> 2;
    uint32_t qf[4][4], dof[4][4];
    load_frags(qf, sQ, r0, lane);
    load_frags(dof, sdO, r0, lane);

    // ---- pass 1: log-sum-exp of the warp's 16 rows (thread: rows g and g + 8) -- skipped when the forward kernel saved it
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.0f, 0.0f};
    for (int kc = 0; kc < (have_lse ? 0 : T / 64); ++kc) {
        float s[8][4];
        mm_abt(s, qf, sK, kc * 64, lane);
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
            const float m_new = fmaxf(m_run[r], mx[r]);
            l_run[r] *= exp2f((m_run[r] - m_new) * scale_log2);
            m_run[r] = m_new;
        }
        float rs[2] = {0.0f, 0.0f};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            rs[0] += exp2f((s[j][0] - m_run[0]) * scale_log2) + exp2f((s[j][1] - m_run[0]) * scale_log2);
            rs[1] += exp2f((s[j][2] - m_run[1]) * scale_log2) + exp2f((s[j][3] - m_run[1]) * scale_log2);
        }
        l_run[0] += rs[0], l_run[1] += rs[1];
    }
    float lse2[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
        lse2[r] = fmaf(m_run[r], scale_log2, log2f(l_run[r]));  // log2 of sum_j exp(scale * s_j)
    }
    if (have_lse) {
        lse2[0] = lse[((size_t)b * heads + h) * T + q0 + r0 + g];
        lse2[1] = lse[((size_t)b * heads + h) * T + q0 + r0 + g + 8];
    } else if ((lane & 3) == 0) {
        lse[((size_t)b * heads + h) * T + q0 + r0 + g] = lse2[0];
        lse[((size_t)b * heads + h) * T + q0 + r0 + g + 8] = lse2[1];
    }
    const float dr[2] = {sD[r0 + g], sD[r0 + g + 8]};

    // ---- pass 2: dQ += (P o (dO V^T - D)) K
    float dq[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.0f;
    for (int kc = 0; kc < T / 64; ++kc) {
        float s[8][4], dp[8][4];
        mm_abt(s, qf, sK, kc * 64, lane);
        mm_abt(dp, dof, sV, kc * 64, lane);
        if (drop_thresh) {  // dP = mask/(1-p) o (dO V^T): the forward dropped these probabilities (attention.cu, k_attention_mma<true>)
            const uint32_t sd = mix32(drop_seed ^ ((uint32_t)(b * heads + h) * 0x9E3779B9u));
            const uint32_t qa = (uint32_t)(q0 + r0 + g) * T, qb = qa + 8u * T;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t key = kc * 64 + j * 8 + 2 * (lane & 3);
                dp[j][0] = dropout_keep(sd, qa + key, drop_thresh) ? dp[j][0] * drop_inv : 0.0f;
                dp[j][1] = dropout_keep(sd, qa + key + 1, drop_thresh) ? dp[j][1] * drop_inv : 0.0f;
                dp[j][2] = dropout_keep(sd, qb + key, drop_thresh) ? dp[j][2] * drop_inv : 0.0f;
                dp[j][3] = dropout_keep(sd, qb + key + 1, drop_thresh) ? dp[j][3] * drop_inv : 0.0f;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = exp2f(fmaf(s[j][0], scale_log2, -lse2[0])) * (dp[j][0] - dr[0]);
            s[j][1] = exp2f(fmaf(s[j][1], scale_log2, -lse2[0])) * (dp[j][1] - dr[0]);
            s[j][2] = exp2f(fmaf(s[j][2], scale_log2, -lse2[1])) * (dp[j][2] - dr[1]);
            s[j][3] = exp2f(fmaf(s[j][3], scale_log2, -lse2[1])) * (dp[j][3] - dr[1]);
        }
        mm_pb(dq, s, sK, kc * 64, lane);
    }
    store_rows(dq, scale, sQ, r0, lane, dqkv + ((size_t)b * T + q0) * ld + h * kHd, ld);
}

__global__ void __launch_bounds__(ab::kThreads, 1)
    k_attention_bwd_kv(__nv_bfloat16* __restrict__ dqkv, const float* __restrict__ lse, const float* __restrict__ dsum, const __nv_bfloat16* __restrict__ qkv,
                       const __nv_bfloat16* __restrict__ dout, int T, int dim, int heads, float scale_log2, float scale, uint32_t drop_thresh,
                       uint32_t drop_seed, float drop_inv) {
    using namespace ab;
    extern __shared__ __align__(128) uint8_t ab_smem[];
    const uint32_t sK = (uint32_t)__cvta_generic_to_shared(ab_smem), sV = sK + kRows * 128, sQ = sV + kRows * 128, sdO = sQ + T * 128;
    float* sL = reinterpret_cast<float*>(ab_smem + 2 * kRows * 128 + 2 * T * 128);
    float* sD = sL + T;
    const int k0 = blockIdx.x * kRows, h = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t ld = 3 * (size_t)dim;
    const __nv_bfloat16* base = qkv + (size_t)b * T * ld + h * kHd;
    stage(sK, base + dim + (size_t)k0 * ld, ld, kRows);
    stage(sV, base + 2 * dim + (size_t)k0 * ld, ld, kRows);
    stage(sQ, base, ld, T);
    stage(sdO, dout + (size_t)b * T * dim + h * kHd, dim, T);
    for (int i = tid; i < T; i += kThreads) {
        sL[i] = lse[((size_t)b * heads + h) * T + i];
        sD[i] = dsum[((size_t)b * heads + h) * T + i];
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int r0 = warp * 16, t4 = lane & 3;
    uint32_t kf[4][4], vf[4][4];
    load_frags(kf, sK, r0, lane);
    load_frags(vf, sV, r0, lane);
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.0f;
    const uint32_t sd = mix32(drop_seed ^ ((uint32_t)(b * heads + h) * 0x9E3779B9u));
    for (int qc = 0; qc < T / 64; ++qc) {
        float s[8][4], dp[8][4];
        mm_abt(s, kf, sQ, qc * 64, lane);     // S^T: rows = this warp's keys, columns = queries of the chunk
        mm_abt(dp, vf, sdO, qc * 64, lane);   // dP^T = V dO^T
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = qc * 64 + j * 8 + 2 * t4;
            const float l0 = sL[c], l1 = sL[c + 1], d0 = sD[c], d1 = sD[c + 1];
            s[j][0] = exp2f(fmaf(s[j][0], scale_log2, -l0)), s[j][1] = exp2f(fmaf(s[j][1], scale_log2, -l1));
            s[j][2] = exp2f(fmaf(s[j][2], scale_log2, -l0)), s[j][3] = exp2f(fmaf(s[j][3], scale_log2, -l1));
            float m00 = 1.0f, m01 = 1.0f, m10 = 1.0f, m11 = 1.0f;  // mask/(1-p) of (key row g | g+8, query column c | c+1)
            if (drop_thresh) {
                const uint32_t ka = (uint32_t)(k0 + r0 + (lane >> 2)), kb = ka + 8u;
                m00 = dropout_keep(sd, (uint32_t)c * T + ka, drop_thresh) ? drop_inv : 0.0f;
                m01 = dropout_keep(sd, (uint32_t)(c + 1) * T + ka, drop_thresh) ? drop_inv : 0.0f;
                m10 = dropout_keep(sd, (uint32_t)c * T + kb, drop_thresh) ? drop_inv : 0.0f;
                m11 = dropout_keep(sd, (uint32_t)(c + 1) * T + kb, drop_thresh) ? drop_inv : 0.0f;
            }
            dp[j][0] = s[j][0] * (dp[j][0] * m00 - d0), dp[j][1] = s[j][1] * (dp[j][1] * m01 - d1);
            dp[j][2] = s[j][2] * (dp[j][2] * m10 - d0), dp[j][3] = s[j][3] * (dp[j][3] * m11 - d1);
            s[j][0] *= m00, s[j][1] *= m01, s[j][2] *= m10, s[j][3] *= m11;  // dV uses the dropped probabilities
        }
        mm_pb(dv, s, sdO, qc * 64, lane);  // dV += P^T dO
        mm_pb(dk, dp, sQ, qc * 64, lane);  // dK += dS^T Q
    }
    __nv_bfloat16* dst = dqkv + ((size_t)b * T + k0) * ld + h * kHd;
    store_rows(dk, scale, sK, r0, lane, dst + dim, ld);
    store_rows(dv, 1.0f, sV, r0, lane, dst + 2 * dim, ld);
}

int attention_backward_tcgen05(void* dqkv_bf16, const float* lse, float* dsum, const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, int B,
                               int heads, float drop_p, uint32_t drop_seed, cudaStream_t stream);  // attention_bwd_sm100.cu

}  // namespace bsi

using namespace bsi;

extern "C" int bsi_attention_backward_bf16(void* dqkv_bf16, float* lse_ws, float* dsum_ws, const void* qkv_bf16, const void* out_bf16, const void* dout_bf16,
                                           int32_t B, int32_t T, int32_t heads, int32_t head_dim, float drop_p, uint32_t drop_seed, int32_t lse_valid,
                                           void* stream) {
    BSI_CHECK_ARG(dqkv_bf16 && lse_ws && dsum_ws && qkv_bf16 && out_bf16 && dout_bf16 && B > 0 && heads > 0 && drop_p >= 0.0f && drop_p < 1.0f,
                  "bsi_attention_backward_bf16: bad arguments");
    const uint32_t drop_thresh = dropout_thresh(drop_p);
    const float drop_inv = drop_p > 0.0f ? 1.0f / (1.0f - drop_p) : 1.0f;
    if (head_dim != ab::kHd || T % ab::kRows != 0 || T > 512) {
        set_error("bsi_attention_backward_bf16: only head_dim=64 and T in {128,256,384,512} are implemented (got head_dim=%d T=%d)", head_dim, T);
        return BSI_ERR_UNSUPPORTED;
    }
    // T = 256 with the forward's saved statistics: the tcgen05 kernel (BSI_ATT_BWD_VARIANT=0 keeps the mma.sync row-owner kernels)
    static const bool use_tc = [] { const char* e = getenv("BSI_ATT_BWD_VARIANT"); return !(e && e[0] == '0'); }();
    if (use_tc && T == 256 && lse_valid)
        return attention_backward_tcgen05(dqkv_bf16, lse_ws, dsum_ws, qkv_bf16, out_bf16, dout_bf16, B, heads, drop_p, drop_seed, (cudaStream_t)stream);
    const int dim = heads * head_dim;
    const float scale = 1.0f / sqrtf((float)head_dim), scale_log2 = 1.4426950408889634f * scale;
    dim3 grid(T / ab::kRows, heads, B);
    const int smem_q = 3 * ab::kRows * 128 + 2 * T * 128 + ab::kRows * 4, smem_kv = 2 * ab::kRows * 128 + 2 * T * 128 + 2 * T * 4;
    BSI_ENSURE_SMEM(k_attention_bwd_q, smem_q);
    k_attention_bwd_q<<<grid, ab::kThreads, smem_q, (cudaStream_t)stream>>>((__nv_bfloat16*)dqkv_bf16, lse_ws, dsum_ws, (const __nv_bfloat16*)qkv_bf16,
                                                                           (const __nv_bfloat16*)out_bf16, (const __nv_bfloat16*)dout_bf16, T, dim, heads,
                                                                           scale_log2, scale, drop_thresh, drop_seed, drop_inv, lse_valid);
    BSI_LAUNCH_OK("k_attention_bwd_q");
    BSI_ENSURE_SMEM(k_attention_bwd_kv, smem_kv);
    k_attention_bwd_kv<<<grid, ab::kThreads, smem_kv, (cudaStream_t)stream>>>((__nv_bfloat16*)dqkv_bf16, lse_ws, dsum_ws, (const __nv_bfloat16*)qkv_bf16,
                                                                             (const __nv_bfloat16*)dout_bf16, T, dim, heads, scale_log2, scale, drop_thresh, drop_seed,
                                                                             drop_inv);
    BSI_LAUNCH_OK("k_attention_bwd_kv");
    return BSI_OK;
}
