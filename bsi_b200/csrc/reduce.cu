// Warp-shuffle reductions for the ELBO / loss terms (HBM-bound).
//   recon:  out[r] = -sum_d log clamp(cdf(edge[idx+1]) - cdf(edge[idx]), 1e-20)   (bsi/bsi.py:230-247)
//   sqerr:  out[r] =  sum_d (x - x_hat)^2                                          (bsi/bsi.py:273,288,309)
// x_hat = c_skip[r]*mu + c_out[r]*f is formed in registers and never written to HBM
// (bsi/bsi.py:381-386).  Algorithmic bytes per element: mu 4 + f 4 (+ x 4, which stays in L2
// across the n Monte-Carlo replicas of the same batch row) -> one float per row out.
// One CTA per row r keeps the summation order fixed (deterministic results).
#include "common.cuh"

namespace bsi {

constexpr int kRThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < (kRThreads / 32) ? red[lane] : 0.0f;
        t = warp_sum(t);
        if (lane == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
}

__device__ __forceinline__ float combine(bool has_mu, float cs, float co, float m, float f) {
    // addcmul(c_skip*mu, c_out, f): torch evaluates addcmul as one FMA on the rounded c_skip*mu
    return has_mu ? __fmaf_rn(co, f, __fmul_rn(cs, m)) : f;
}

__device__ __forceinline__ int bucket_of_r(float x, float lo_edge, float dx, int k) {
    float q = __fdiv_rn(__fsub_rn(x, lo_edge), dx);
    q = fminf(fmaxf(q, -1.0f), (float)k);
    int i = (int)q;
    return min(max(i, 0), k - 1);
}

// torch.distributions.Normal.cdf: 0.5 * (1 + erf((v - loc) * scale.reciprocal() / sqrt(2))), split in its two halves
__device__ __forceinline__ float cdf_arg(float v, float loc, float inv_scale) {
    return __fdiv_rn(__fmul_rn(__fsub_rn(v, loc), inv_scale), 1.4142135623730951f);
}
__device__ __forceinline__ float cdf_of_erf(float e) { return __fmul_rn(0.5f, __fadd_rn(1.0f, e)); }
// |z| >= 4: erfc(4) = 1.5e-8 is a quarter of the fp32 spacing below 1, so erff(z) is exactly +-1 and the cdf exactly 0 or 1
constexpr float kErfSaturated = 4.0f;

template <bool VEC>
__global__ void __launch_bounds__(kRThreads)
    k_recon_reduce(float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ mu, const float* __restrict__ f,
                   const float* __restrict__ c_skip, const float* __restrict__ c_out, const float* __restrict__ edges, int k,
                   float lo_edge, float dx, float inv_scale, int64_t B, int64_t D) {
    extern __shared__ float s_edges[];  // k+1 boundaries
    __shared__ float red[kRThreads / 32];
    for (int i = threadIdx.x; i <= k; i += blockDim.x) s_edges[i] = edges[i];
    __syncthreads();
    const int64_t r = blockIdx.x, b = r % B;
    const bool has_mu = mu != nullptr;
    const float cs = has_mu ? c_skip[r] : 0.0f, co = has_mu ? c_out[r] : 1.0f;
    const float4* x4 = reinterpret_cast<const float4*>(x + b * D);
    const float4* m4 = has_mu ? reinterpret_cast<const float4*>(mu + r * D) : nullptr;
    const float4* f4 = reinterpret_cast<const float4*>(f + r * D);
    float acc = 0.0f;
    // VEC: four elements per 16-byte access; otherwise (element count not a multiple of 4: rows are not 16-byte aligned) one
    const int64_t items = VEC ? (D >> 2) : D;
    for (int64_t q = threadIdx.x; q < items; q += blockDim.x) {
        float xs[4], fs[4], ms[4] = {0.f, 0.f, 0.f, 0.f};
        if constexpr (VEC) {
            const float4 xv = x4[q], fv = __ldcs(f4 + q), mv = has_mu ? __ldcs(m4 + q) : make_float4(0, 0, 0, 0);
            xs[0] = xv.x, xs[1] = xv.y, xs[2] = xv.z, xs[3] = xv.w, fs[0] = fv.x, fs[1] = fv.y, fs[2] = fv.z, fs[3] = fv.w;
            ms[0] = mv.x, ms[1] = mv.y, ms[2] = mv.z, ms[3] = mv.w;
        } else {
            xs[0] = x[b * D + q], fs[0] = f[r * D + q], ms[0] = has_mu ? mu[r * D + q] : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < (VEC ? 4 : 1); ++j) {
            float xh = combine(has_mu, cs, co, ms[j], fs[j]);
            int idx = bucket_of_r(xs[j], lo_edge, dx, k);
            // The bin is 11 sigma wide (image_8bit with alpha_R = 2e6), so at most one of its two edges lies within the 4 sqrt(2)
            // sigma of x_hat where erf is not saturated -- except in a thin sliver around the bin centre.  One erff per element
            // serves whichever edge needs it; the other cdf is the exact constant.  Same values as evaluating both.
            const float zl = cdf_arg(s_edges[idx], xh, inv_scale), zr = cdf_arg(s_edges[idx + 1], xh, inv_scale);
            const bool nl = idx != 0 && fabsf(zl) < kErfSaturated, nr = idx != k - 1 && fabsf(zr) < kErfSaturated;
            const float e = erff(nl ? zl : zr);
            float left = idx == 0 ? 0.0f : (nl ? cdf_of_erf(e) : (zl < 0.0f ? 0.0f : 1.0f));
            float right = idx == k - 1 ? 1.0f : (nr ? cdf_of_erf(e) : (zr < 0.0f ? 0.0f : 1.0f));
            if (nl && nr) right = cdf_of_erf(erff(zr));  // both edges unsaturated (x_hat within 0.08 sigma-units of the centre)
            acc -= logf(fmaxf(__fsub_rn(right, left), 1e-20f));
        }
    }
    float total = block_sum(acc, red);
    if (threadIdx.x == 0) out[r] = total;
}

template <bool VEC>
__global__ void __launch_bounds__(kRThreads)
    k_sqerr_reduce(float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ mu, const float* __restrict__ f,
                   const float* __restrict__ c_skip, const float* __restrict__ c_out, int64_t B, int64_t D) {
    __shared__ float red[kRThreads / 32];
    const int64_t r = blockIdx.x, b = r % B;
    const bool has_mu = mu != nullptr;
    const float cs = has_mu ? c_skip[r] : 0.0f, co = has_mu ? c_out[r] : 1.0f;
    const float4* x4 = reinterpret_cast<const float4*>(x + b * D);
    const float4* m4 = has_mu ? reinterpret_cast<const float4*>(mu + r * D) : nullptr;
    const float4* f4 = reinterpret_cast<const float4*>(f + r * D);
    float acc = 0.0f;
    if constexpr (VEC) {
        for (int64_t q = threadIdx.x; q < (D >> 2); q += blockDim.x) {
            float4 xv = x4[q];
            float4 fv = __ldcs(f4 + q);
            float4 mv = has_mu ? __ldcs(m4 + q) : make_float4(0, 0, 0, 0);
            float xs[4] = {xv.x, xv.y, xv.z, xv.w}, fs[4] = {fv.x, fv.y, fv.z, fv.w}, ms[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float d = __fsub_rn(xs[j], combine(has_mu, cs, co, ms[j], fs[j]));
                acc = fmaf(d, d, acc);
            }
        }
    } else {
        for (int64_t e = threadIdx.x; e < D; e += blockDim.x) {
            float d = __fsub_rn(x[b * D + e], combine(has_mu, cs, co, has_mu ? mu[r * D + e] : 0.0f, f[r * D + e]));
            acc = fmaf(d, d, acc);
        }
    }
    float total = block_sum(acc, red);
    if (threadIdx.x == 0) out[r] = total;
}

// d/df of  w[r] * sum_d (x - (c_skip*mu + c_out*f))^2  =  -2 * w[r] * c_out[r] * (x - x_hat)
// (autograd of bsi/bsi.py:309-310 w.r.t. the denoiser output).  16 B/elem.
// scalar variant for element counts that are not a multiple of 4
__global__ void __launch_bounds__(kRThreads)
    k_sqerr_backward_s(float* __restrict__ grad_f, const float* __restrict__ w, const float* __restrict__ x, const float* __restrict__ mu,
                       const float* __restrict__ f, const float* __restrict__ c_skip, const float* __restrict__ c_out, int64_t B, int64_t D) {
    const int e = blockIdx.y * kRThreads + threadIdx.x;
    if (e >= D) return;
    const uint32_t r32 = blockIdx.x;
    const int64_t r = r32, b = r32 % (uint32_t)B;
    const bool has_mu = mu != nullptr;
    const float cs = has_mu ? c_skip[r] : 0.0f, co = has_mu ? c_out[r] : 1.0f;
    grad_f[r * D + e] = -2.0f * w[r] * co * (x[b * D + e] - combine(has_mu, cs, co, has_mu ? mu[r * D + e] : 0.0f, f[r * D + e]));
}

__global__ void __launch_bounds__(kRThreads)
    k_sqerr_backward(float* __restrict__ grad_f, const float* __restrict__ w, const float* __restrict__ x,
                     const float* __restrict__ mu, const float* __restrict__ f, const float* __restrict__ c_skip,
                     const float* __restrict__ c_out, int64_t R, int64_t B, int64_t D) {
    // blockIdx.x = row, blockIdx.y = chunk of 256 quads: no 64-bit division per element
    const int q = blockIdx.y * kRThreads + threadIdx.x;
    if (q >= (int)(D >> 2)) return;
    const uint32_t r32 = blockIdx.x;
    const int64_t r = r32, b = r32 % (uint32_t)B;
    const bool has_mu = mu != nullptr;
    const float cs = has_mu ? c_skip[r] : 0.0f, co = has_mu ? c_out[r] : 1.0f;
    const float g = -2.0f * w[r] * co;
    float4 xv = *reinterpret_cast<const float4*>(x + b * D + (int64_t)q * 4);
    float4 fv = __ldcs(reinterpret_cast<const float4*>(f + r * D) + q);
    float4 mv = has_mu ? __ldcs(reinterpret_cast<const float4*>(mu + r * D) + q) : make_float4(0, 0, 0, 0);
    float4 o;
    o.x = g * (xv.x - combine(has_mu, cs, co, mv.x, fv.x));
    o.y = g * (xv.y - combine(has_mu, cs, co, mv.y, fv.y));
    o.z = g * (xv.z - combine(has_mu, cs, co, mv.z, fv.z));
    o.w = g * (xv.w - combine(has_mu, cs, co, mv.w, fv.w));
    *reinterpret_cast<float4*>(grad_f + r * D + (int64_t)q * 4) = o;
}

}  // namespace bsi

using namespace bsi;

extern "C" {

int bsi_recon_reduce(float* out, const float* x, const float* mu, const float* f, const float* c_skip, const float* c_out,
                     const float* edges, int32_t k, float lo_edge, float dx, float inv_scale, int64_t R, int64_t B, int64_t D,
                     void* stream) {
    BSI_CHECK_ARG(out && x && f && edges && R > 0 && B > 0, "bsi_recon_reduce: null pointer or empty batch");
    BSI_CHECK_ARG(!mu || (c_skip && c_out), "bsi_recon_reduce: mu given without c_skip/c_out");
    BSI_CHECK_ARG(k >= 2 && k <= 1024, "bsi_recon_reduce: k=%d outside [2,1024]", k);
    BSI_CHECK_ARG(D > 0, "data numel per sample (%lld) must be positive", (long long)D);
    BSI_CHECK_ARG(R <= 0x7fffffff, "too many rows");
    if (D % 4 == 0)
        k_recon_reduce<true><<<(unsigned)R, kRThreads, (k + 1) * sizeof(float), (cudaStream_t)stream>>>(out, x, mu, f, c_skip, c_out, edges, k, lo_edge, dx,
                                                                                                       inv_scale, B, D);
    else
        k_recon_reduce<false><<<(unsigned)R, kRThreads, (k + 1) * sizeof(float), (cudaStream_t)stream>>>(out, x, mu, f, c_skip, c_out, edges, k, lo_edge, dx,
                                                                                                        inv_scale, B, D);
    BSI_LAUNCH_OK("k_recon_reduce");
    return BSI_OK;
}

int bsi_sqerr_reduce(float* out, const float* x, const float* mu, const float* f, const float* c_skip, const float* c_out,
                     int64_t R, int64_t B, int64_t D, void* stream) {
    BSI_CHECK_ARG(out && x && f && R > 0 && B > 0, "bsi_sqerr_reduce: null pointer or empty batch");
    BSI_CHECK_ARG(!mu || (c_skip && c_out), "bsi_sqerr_reduce: mu given without c_skip/c_out");
    BSI_CHECK_ARG(D > 0, "data numel per sample (%lld) must be positive", (long long)D);
    BSI_CHECK_ARG(R <= 0x7fffffff, "too many rows");
    if (D % 4 == 0) k_sqerr_reduce<true><<<(unsigned)R, kRThreads, 0, (cudaStream_t)stream>>>(out, x, mu, f, c_skip, c_out, B, D);
    else k_sqerr_reduce<false><<<(unsigned)R, kRThreads, 0, (cudaStream_t)stream>>>(out, x, mu, f, c_skip, c_out, B, D);
    BSI_LAUNCH_OK("k_sqerr_reduce");
    return BSI_OK;
}

int bsi_sqerr_backward(float* grad_f, const float* w, const float* x, const float* mu, const float* f, const float* c_skip,
                       const float* c_out, int64_t R, int64_t B, int64_t D, void* stream) {
    BSI_CHECK_ARG(grad_f && w && x && f && R > 0 && B > 0, "bsi_sqerr_backward: null pointer or empty batch");
    BSI_CHECK_ARG(!mu || (c_skip && c_out), "bsi_sqerr_backward: mu given without c_skip/c_out");
    BSI_CHECK_ARG(D > 0, "data numel per sample (%lld) must be positive", (long long)D);
    if (D % 4 != 0) {
        const int64_t chunks_s = (D + kRThreads - 1) / kRThreads;
        BSI_CHECK_ARG(R <= 0x7fffffffLL && B <= 0x7fffffffLL && chunks_s <= 65535, "bsi_sqerr_backward: shape out of range");
        k_sqerr_backward_s<<<dim3((unsigned)R, (unsigned)chunks_s), kRThreads, 0, (cudaStream_t)stream>>>(grad_f, w, x, mu, f, c_skip, c_out, B, D);
        BSI_LAUNCH_OK("k_sqerr_backward_s");
        return BSI_OK;
    }
    const int64_t chunks = ((D >> 2) + kRThreads - 1) / kRThreads;
    BSI_CHECK_ARG(R <= 0x7fffffffLL && B <= 0x7fffffffLL && chunks <= 65535, "bsi_sqerr_backward: shape out of range");
    k_sqerr_backward<<<dim3((unsigned)R, (unsigned)chunks), kRThreads, 0, (cudaStream_t)stream>>>(grad_f, w, x, mu, f, c_skip, c_out, R, B, D);
    BSI_LAUNCH_OK("k_sqerr_backward");
    return BSI_OK;
}

}  // extern "C"
