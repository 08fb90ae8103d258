// fp32-accurate ("exact") mode of the native DiT (reference evaluation precision: fp32 weights and activations,
// bsi/lightning/plugins.py:7-24, config/train.yaml:38): the tensor cores still do the work, but every GEMM operand is split into
// two bf16 terms, x = hi + lo with hi = bf16(x), lo = bf16(x - hi), and the product is formed as
//     A W^T  ~=  A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T           (the dropped lo*lo term is 2^-16 relative)
// by ONE ordinary bf16 GEMM over a three times longer K: activations are stored as rows [hi | lo | hi], weights as [hi | hi | lo].
// The bf16 products are exact in the fp32 accumulator, so the result carries fp32-level error (~1e-6) instead of bf16's 4e-3.
// This file holds what that mode needs besides the GEMM: the splitting kernel (with the activations fused), an fp32 attention,
// and fp32-output variants of the operand builders.  Everything here is bandwidth- or FMA-bound SIMT code: the mode trades a 3-5x
// longer forward for reference-grade numbers (ELBO evaluation, the 1e-5 trajectory tier), it is not the throughput path.
#include "common.cuh"

namespace bsi {

constexpr int kExThreads = 256;
static inline int ex_grid(int64_t items) {
    int64_t need = (items + kExThreads - 1) / kExThreads, cap = (int64_t)sm_count() * 8;
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

__device__ __forceinline__ float ex_act(float x, int act) {
    if (act == 1) {  // nn.GELU(approximate="tanh") (dit.py:75), library tanhf
        const float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
        return 0.5f * x * (1.0f + tanhf(u));
    }
    if (act == 2) return x / (1.0f + expf(-x));  // SiLU (dit.py:80)
    return x;
}

// out[r][3*bp]: activation layout [hi | lo | hi] (weight_layout = 0) or weight layout [hi | hi | lo] (1); bp = pitch of one block >= cols,
// padding columns are zero.  in[r][c] at in + r*ld_in + c.
__global__ void __launch_bounds__(kExThreads) k_split3(__nv_bfloat16* __restrict__ out, const float* __restrict__ in, int64_t rows, int cols, int64_t ld_in, int bp,
                                                       int weight_layout, int act) {
    const int64_t total = rows * bp;
    for (int64_t i = (int64_t)blockIdx.x * kExThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kExThreads) {
        const int64_t r = i / bp;
        const int c = (int)(i - r * bp);
        const float v = c < cols ? ex_act(in[r * ld_in + c], act) : 0.0f;
        const __nv_bfloat16 hi = __float2bfloat16(v);
        const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
        __nv_bfloat16* o = out + r * 3 * (int64_t)bp + c;
        o[0] = hi;
        o[bp] = weight_layout ? hi : lo;
        o[2 * bp] = weight_layout ? lo : hi;
    }
}

// fp32 attention over packed fp32 QKV (dit.py:36-47): CTA = 32 queries of one (sample, head); scores of the 32 x T tile in shared memory.
//   phase 1: thread <-> key column(s): s[q][k] = <Q[q], K[k]>           (K tile in smem, padded rows: conflict-free)
//   phase 2: warp <-> rows: softmax in place (library expf)
//   phase 3: thread <-> (channel, query group): O[q][c] = sum_k P[q][k] V[k][c]   (V tile overwrites the K tile)
constexpr int kAfQ = 32, kAfHd = 64;
__global__ void __launch_bounds__(kExThreads) k_attention_f32(float* __restrict__ out, const float* __restrict__ qkv, int T, int dim, float scale) {
    extern __shared__ float af_smem[];
    float* sQ = af_smem;                    // [32][64]
    float* sKV = sQ + kAfQ * kAfHd;         // [T][65]
    float* sS = sKV + (size_t)T * 65;       // [32][T]
    const int q0 = blockIdx.x * kAfQ, h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
    const size_t ld = 3 * (size_t)dim;
    const float* base = qkv + (size_t)b * T * ld + h * kAfHd;
    for (int i = tid; i < kAfQ * kAfHd; i += kExThreads) sQ[i] = base[(size_t)(q0 + i / kAfHd) * ld + (i % kAfHd)];
    for (int i = tid; i < T * kAfHd; i += kExThreads) sKV[(i / kAfHd) * 65 + (i % kAfHd)] = base[dim + (size_t)(i / kAfHd) * ld + (i % kAfHd)];
    __syncthreads();
    for (int k = tid; k < T; k += kExThreads) {
        float acc[kAfQ];
#pragma unroll
        for (int q = 0; q < kAfQ; ++q) acc[q] = 0.0f;
        for (int c = 0; c < kAfHd; ++c) {
            const float kv = sKV[k * 65 + c];
#pragma unroll
            for (int q = 0; q < kAfQ; ++q) acc[q] = fmaf(sQ[q * kAfHd + c], kv, acc[q]);
        }
#pragma unroll
        for (int q = 0; q < kAfQ; ++q) sS[q * T + k] = acc[q] * scale;
    }
    __syncthreads();
    {
        const int warp = tid >> 5, lane = tid & 31;
        for (int q = warp; q < kAfQ; q += kExThreads / 32) {
            float mx = -INFINITY;
            for (int k = lane; k < T; k += 32) mx = fmaxf(mx, sS[q * T + k]);
            mx = warp_max(mx);
            float sum = 0.0f;
            for (int k = lane; k < T; k += 32) {
                const float p = expf(sS[q * T + k] - mx);
                sS[q * T + k] = p;
                sum += p;
            }
            const float inv = 1.0f / warp_sum(sum);
            for (int k = lane; k < T; k += 32) sS[q * T + k] *= inv;
        }
    }
    // V tile over the K tile (row pitch 64 now: consecutive channels -> consecutive banks)
    __syncthreads();
    for (int i = tid; i < T * kAfHd; i += kExThreads) sKV[i] = base[2 * dim + (size_t)(i / kAfHd) * ld + (i % kAfHd)];
    __syncthreads();
    const int c = tid & 63, qg = tid >> 6;  // 4 query groups x 8 queries
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.0f;
    for (int k = 0; k < T; ++k) {
        const float v = sKV[k * kAfHd + c];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(sS[(qg + 4 * j) * T + k], v, o[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) out[((size_t)b * T + q0 + qg + 4 * j) * dim + h * kAfHd + c] = o[j];
}

// fp32 patch-embed operand with the reference's exact angle arithmetic (fourier_features.py:24-36: coefs fp32(2 pi 2^n), args = addcmul(offset,
// coef, x), sin): [B*T][P] fp32, one thread per pixel (dit.py:149-153,228-231 + c_in scaling, bsi.py:385).
__global__ void __launch_bounds__(kExThreads)
    k_patch_operand_f32(float* __restrict__ A, const float* __restrict__ mu, bsi_rowref scale, const int32_t* __restrict__ step_ptr, int B, int C, int H, int W, int p,
                        int n_min, int n_max, int lda) {
    const int step = step_ptr ? *step_ptr : 0;
    const int nfreq = n_max >= n_min ? n_max - n_min + 1 : 0;
    const int cin = C * (1 + 2 * nfreq);
    const int gw = W / p;
    const int64_t HW = (int64_t)H * W, total = (int64_t)B * HW;
    const int T = (H / p) * gw;
    for (int64_t i = (int64_t)blockIdx.x * kExThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kExThreads) {
        const int64_t b = i / HW;
        const int pix = (int)(i - b * HW);
        const int y = pix / W, x = pix - y * W;
        float* dst = A + ((int64_t)b * T + (y / p) * gw + (x / p)) * lda + (int64_t)((y % p) * p + (x % p)) * cin;
        const float sc = rowref_at(scale, b, step);
        for (int c = 0; c < C; ++c) {
            const float v = __fmul_rn(sc, mu[(b * C + c) * HW + pix]);
            dst[c] = v;
            for (int f = 0; f < nfreq; ++f) {
                const float coef = 6.283185307179586f * (float)(1 << (n_min + f));
                dst[C + c * 2 * nfreq + 2 * f] = sinf(__fmul_rn(coef, v));
                dst[C + c * 2 * nfreq + 2 * f + 1] = sinf(__fmaf_rn(coef, v, 1.5707963267948966f));
            }
        }
    }
}

}  // namespace bsi

using namespace bsi;

extern "C" {

int bsi_split3_bf16(void* out_bf16, const float* in, int64_t rows, int32_t cols, int64_t ld_in, int32_t block_pitch, int32_t weight_layout, int32_t act,
                    void* stream) {
    BSI_CHECK_ARG(out_bf16 && in && rows > 0 && cols > 0 && ld_in >= cols && block_pitch >= cols && block_pitch % 8 == 0 && act >= 0 && act <= 2,
                  "bsi_split3_bf16: bad arguments (rows=%lld cols=%d ld_in=%lld block_pitch=%d act=%d)", (long long)rows, cols, (long long)ld_in, block_pitch, act);
    k_split3<<<ex_grid(rows * block_pitch), kExThreads, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)out_bf16, in, rows, cols, ld_in, block_pitch, weight_layout, act);
    BSI_LAUNCH_OK("k_split3");
    return BSI_OK;
}

int bsi_attention_f32(float* out, const float* qkv, int32_t B, int32_t T, int32_t heads, int32_t head_dim, void* stream) {
    BSI_CHECK_ARG(out && qkv && B > 0 && heads > 0, "bsi_attention_f32: bad arguments");
    BSI_CHECK_ARG(head_dim == kAfHd && T % kAfQ == 0 && T > 0 && T <= 512, "bsi_attention_f32: head_dim must be 64 and T a multiple of 32 up to 512 (got %d, %d)", head_dim, T);
    BSI_CHECK_ARG(B <= 65535 && heads <= 65535, "bsi_attention_f32: grid out of range");
    const int smem = (kAfQ * kAfHd + T * 65 + kAfQ * T) * (int)sizeof(float);
    BSI_ENSURE_SMEM(k_attention_f32, smem);
    k_attention_f32<<<dim3(T / kAfQ, heads, B), kExThreads, smem, (cudaStream_t)stream>>>(out, qkv, T, heads * head_dim, 1.0f / sqrtf((float)head_dim));
    BSI_LAUNCH_OK("k_attention_f32");
    return BSI_OK;
}

int bsi_dit_patch_operand_f32(float* A, const float* mu, bsi_rowref scale, const int32_t* step_ptr, int32_t B, int32_t C, int32_t H, int32_t Wd, int32_t patch,
                              int32_t n_min, int32_t n_max, int32_t lda, void* stream) {
    BSI_CHECK_ARG(A && mu && scale.base && B > 0 && C > 0 && patch > 0 && H % patch == 0 && Wd % patch == 0, "bsi_dit_patch_operand_f32: bad arguments");
    BSI_CHECK_ARG(n_max < n_min || (n_min >= 0 && n_max < 24), "Fourier exponents out of range");
    const int nfreq = n_max >= n_min ? n_max - n_min + 1 : 0;
    BSI_CHECK_ARG(lda >= patch * patch * C * (1 + 2 * nfreq), "operand pitch %d too small", lda);
    k_patch_operand_f32<<<ex_grid((int64_t)B * H * Wd), kExThreads, 0, (cudaStream_t)stream>>>(A, mu, scale, step_ptr, B, C, H, Wd, patch, n_min, n_max, lda);
    BSI_LAUNCH_OK("k_patch_operand_f32");
    return BSI_OK;
}

}  // extern "C"
