// HBM-bound elementwise kernels of the BSI hot path: sampler init/step, q-sample, EDM combine,
// bucketize, casts, time embedding and the patch-embed operand builder.
// Each kernel states its algorithmic bytes per element; all use 128-bit accesses on the
// contiguous axis and a grid of (SM count x resident CTAs) with a grid-stride loop.
#include <atomic>
#include <cstdlib>

#include "common.cuh"

namespace bsi {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool use_pdl() {
    static int cached = -1;
    if (cached < 0) {
        // measured on B200 (imagenet64-dit, k = 64): the captured sampler graph ran 18 % SLOWER with programmatic edges
        // (3.51 s vs 2.98 s per step), the eager forward 1 % slower — so PDL is opt-in (BSI_PDL=1), not the default
        const char* e = getenv("BSI_PDL");
        cached = (e && e[0] == '1') ? 1 : 0;
    }
    return cached == 1;
}
void fill_pdl_attr(cudaLaunchAttribute* attr) {
    attr->id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr->val.programmaticStreamSerializationAllowed = 1;
}
int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

constexpr int kThreads = 256;
static inline int grid_for(int64_t work_items, int ctas_per_sm = 8) {
    int64_t need = (work_items + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4_stream(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Rounding points follow the reference's eager torch ops: torch.addcmul(a, b, c) is a single FMA (measured on both
// the CPU and the CUDA backend), every other op of a chain rounds separately; the explicit _rn intrinsics pin this
// down so nvcc cannot contract differently.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float addcmul_rn(float a, float b, float c) { return __fmaf_rn(b, c, a); }  // torch.addcmul: one FMA (CPU and CUDA)
__device__ __forceinline__ float add_mul_sep(float a, float b, float c) { return __fadd_rn(a, __fmul_rn(b, c)); }  // a + b*c as two eager ops

// by-value key, or the {seed, sample_base} pair a replayed graph reads from device memory
__device__ __forceinline__ bsi_noise resolve_noise(bsi_noise nz) {
    if (nz.key_ptr) nz.seed = nz.key_ptr[0], nz.sample_base = nz.key_ptr[1];
    return nz;
}
__device__ __forceinline__ float4 noise4(const bsi_noise& nz, int step, int64_t sample, int64_t quad, int64_t D) {
    if (nz.eps) return ld4_stream(nz.eps + sample * D + quad * 4);
    return philox_normal4(nz.seed, nz.sample_base + (uint64_t)sample, (uint32_t)(nz.draw + step), (uint32_t)quad);
}

// Row kernels below use a 2-D grid: blockIdx.x = sample / row, blockIdx.y = chunk of 256 float4 quads inside the row, so there is no
// 64-bit division per element (the first version spent as many instructions on `i / qpr` and `r % B` as on Philox).
struct RowGrid {
    dim3 grid;
};
static inline RowGrid row_grid(int64_t rows, int64_t D) {
    const int64_t qpr = D >> 2;
    RowGrid g;
    g.grid = dim3((unsigned)rows, (unsigned)((qpr + kThreads - 1) / kThreads), 1);
    return g;
}
#define BSI_CHECK_ROW_GRID(rows, D)                                                                                   \
    BSI_CHECK_ARG((rows) <= 0x7fffffffLL && (((D) >> 2) + kThreads - 1) / kThreads <= 65535, "row grid out of range: %lld rows of %lld", \
                  (long long)(rows), (long long)(D))

// ------------------------------------------------------------------ sampler init (bsi/bsi.py:325-327)
// bytes/elem: 4 written (+4 read when noise is injected)
__global__ void __launch_bounds__(kThreads) k_sample_init(float* __restrict__ mu, const float* __restrict__ sigma0_ptr,
                                                          bsi_noise nz_arg, int64_t n, int64_t D) {
    const bsi_noise nz = resolve_noise(nz_arg);
    const float s0 = sigma0_ptr[0];
    const int64_t s = blockIdx.x;
    const int q = blockIdx.y * kThreads + threadIdx.x;
    if (q >= (int)(D >> 2)) return;
    float4 e = noise4(nz, 0, s, q, D);
    st4(mu + s * D + q * 4, make_float4(s0 * e.x, s0 * e.y, s0 * e.z, s0 * e.w));
}

// ------------------------------------------------------------------ fused sampler step (bsi/bsi.py:331-335, 381-386)
// bytes/elem: read mu 4 + read f 4 + write mu' 4 = 12 (+4 injected eps, +8 when history outputs are requested)
template <bool kPrecond>
__global__ void __launch_bounds__(kThreads)
    k_step_fused(float* __restrict__ mu, const float* __restrict__ f, const float* __restrict__ coef,
                 const int32_t* __restrict__ step_ptr, int32_t step_arg, bsi_noise nz_arg, float* __restrict__ x_hat_out,
                 float* __restrict__ y_out, int64_t n, int64_t D) {
    const int q = blockIdx.y * kThreads + threadIdx.x;
    if (q >= (int)(D >> 2)) return;
    const int64_t s = blockIdx.x, off = s * D + (int64_t)q * 4;
    // the two streaming loads are issued before the (long) Philox chain so that their latency hides under it
    const float4 m = ld4(mu + off);
    const float4 fo = ld4_stream(f + off);
    const bsi_noise nz = resolve_noise(nz_arg);
    const int step = step_ptr ? *step_ptr : step_arg;
    const float* c = coef + (int64_t)step * 8;
    const float c_skip = c[0], c_out = c[1], sigma = c[2], alpha = c[3], lam = c[4], lam_next = c[5];
    const float4 e = noise4(nz, step, s, q, D);
    const float mv[4] = {m.x, m.y, m.z, m.w}, fv[4] = {fo.x, fo.y, fo.z, fo.w}, ev[4] = {e.x, e.y, e.z, e.w};
    float xh[4], yv[4], out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        // x_hat = addcmul(c_skip*mu, c_out, f); y = x_hat + rsqrt(alpha)*eps; mu' = (alpha*y + lam*mu)/lam_next
        xh[j] = kPrecond ? addcmul_rn(mul_rn(c_skip, mv[j]), c_out, fv[j]) : fv[j];
        yv[j] = add_mul_sep(xh[j], sigma, ev[j]);
        out[j] = __fdiv_rn(__fadd_rn(mul_rn(alpha, yv[j]), mul_rn(lam, mv[j])), lam_next);
    }
    st4(mu + off, make_float4(out[0], out[1], out[2], out[3]));
    if (x_hat_out) st4(x_hat_out + off, make_float4(xh[0], xh[1], xh[2], xh[3]));
    if (y_out) st4(y_out + off, make_float4(yv[0], yv[1], yv[2], yv[3]));
}

__global__ void k_step_advance(int32_t* step_ptr) { *step_ptr += 1; }

// ------------------------------------------------------------------ EDM combine / row scaling (bsi/bsi.py:381-386)
// combine: 12 B/elem; scale: 8 B/elem
__global__ void __launch_bounds__(kThreads)
    k_edm_combine(float* __restrict__ x_hat, const float* __restrict__ mu, const float* __restrict__ f, bsi_rowref c_skip,
                  bsi_rowref c_out, const int32_t* __restrict__ step_ptr, int64_t n, int64_t D) {
    const int q = blockIdx.y * kThreads + threadIdx.x;
    if (q >= (int)(D >> 2)) return;
    const int64_t s = blockIdx.x, off = s * D + (int64_t)q * 4;
    const int step = step_ptr ? *step_ptr : 0;
    const float cs = rowref_at(c_skip, s, step), co = rowref_at(c_out, s, step);
    const float4 m = ld4_stream(mu + off), fo = ld4_stream(f + off);
    st4(x_hat + off, make_float4(addcmul_rn(mul_rn(cs, m.x), co, fo.x), addcmul_rn(mul_rn(cs, m.y), co, fo.y),
                                 addcmul_rn(mul_rn(cs, m.z), co, fo.z), addcmul_rn(mul_rn(cs, m.w), co, fo.w)));
}
__global__ void __launch_bounds__(kThreads) k_scale_rows(float* __restrict__ out, const float* __restrict__ in, bsi_rowref sc,
                                                         const int32_t* __restrict__ step_ptr, int64_t n, int64_t D) {
    const int q = blockIdx.y * kThreads + threadIdx.x;
    if (q >= (int)(D >> 2)) return;
    const int64_t s = blockIdx.x, off = s * D + (int64_t)q * 4;
    const int step = step_ptr ? *step_ptr : 0;
    const float c = rowref_at(sc, s, step);
    const float4 m = ld4_stream(in + off);
    st4(out + off, make_float4(c * m.x, c * m.y, c * m.z, c * m.w));
}

// ------------------------------------------------------------------ q(mu|x,lambda) (bsi/bsi.py:405-420)
// bytes/elem: read x 4 (L2-resident across the n replicas) + write mu 4 (+4 model_in, +4 injected eps)
__global__ void __launch_bounds__(kThreads)
    k_q_sample(float* __restrict__ mu, float* __restrict__ model_in, const float* __restrict__ x, const float* __restrict__ gamma,
               const float* __restrict__ sigma, const float* __restrict__ c_in, bsi_noise nz_arg, int64_t R, int64_t B, int64_t D) {
    const int q = blockIdx.y * kThreads + threadIdx.x;
    if (q >= (int)(D >> 2)) return;
    const uint32_t r32 = blockIdx.x;
    const int64_t r = r32, b = r32 % (uint32_t)B, off = r * D + (int64_t)q * 4;
    const float4 xv = ld4(x + b * D + (int64_t)q * 4);
    const bsi_noise nz = resolve_noise(nz_arg);
    const float g = gamma[r], sg = sigma[r];
    const float4 e = noise4(nz, 0, r, q, D);
    // addcmul(gamma*x, sigma, eps)
    const float4 m = make_float4(addcmul_rn(mul_rn(g, xv.x), sg, e.x), addcmul_rn(mul_rn(g, xv.y), sg, e.y),
                                 addcmul_rn(mul_rn(g, xv.z), sg, e.z), addcmul_rn(mul_rn(g, xv.w), sg, e.w));
    st4(mu + off, m);
    if (model_in) {
        const float ci = c_in[r];
        st4(model_in + off, make_float4(ci * m.x, ci * m.y, ci * m.z, ci * m.w));
    }
}

// ------------------------------------------------------------------ scalar variants of the row kernels
// Data shapes whose element count is not a multiple of 4 (the reference accepts any data_shape): rows are not 16-byte aligned, so
// these kernels use 4-byte accesses, one element per thread.  Element e still takes component e & 3 of Philox quad e >> 2, i.e. the
// noise of an element does not depend on which variant runs.  Same arithmetic (rounding points) as the vector kernels.
__device__ __forceinline__ float noise1(const bsi_noise& nz, int step, int64_t sample, int e, int64_t D) {
    if (nz.eps) return __ldcs(nz.eps + sample * D + e);
    const float4 v = philox_normal4(nz.seed, nz.sample_base + (uint64_t)sample, (uint32_t)(nz.draw + step), (uint32_t)(e >> 2));
    const int c = e & 3;
    return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w;
}
static inline dim3 row_grid_scalar(int64_t rows, int64_t D) { return dim3((unsigned)rows, (unsigned)((D + kThreads - 1) / kThreads), 1); }
#define BSI_CHECK_ROW_GRID_SCALAR(rows, D) \
    BSI_CHECK_ARG((D) > 0 && (rows) <= 0x7fffffffLL && ((D) + kThreads - 1) / kThreads <= 65535, "row grid out of range: %lld rows of %lld", (long long)(rows), (long long)(D))

__global__ void __launch_bounds__(kThreads) k_sample_init_s(float* __restrict__ mu, const float* __restrict__ sigma0_ptr, bsi_noise nz_arg, int64_t D) {
    const int e = blockIdx.y * kThreads + threadIdx.x;
    if (e >= D) return;
    const bsi_noise nz = resolve_noise(nz_arg);
    const int64_t s = blockIdx.x;
    mu[s * D + e] = sigma0_ptr[0] * noise1(nz, 0, s, e, D);
}
template <bool kPrecond>
__global__ void __launch_bounds__(kThreads)
    k_step_fused_s(float* __restrict__ mu, const float* __restrict__ f, const float* __restrict__ coef, const int32_t* __restrict__ step_ptr, int32_t step_arg,
                   bsi_noise nz_arg, float* __restrict__ x_hat_out, float* __restrict__ y_out, int64_t D) {
    const int e = blockIdx.y * kThreads + threadIdx.x;
    if (e >= D) return;
    const int64_t s = blockIdx.x, off = s * D + e;
    const bsi_noise nz = resolve_noise(nz_arg);
    const int step = step_ptr ? *step_ptr : step_arg;
    const float* c = coef + (int64_t)step * 8;
    const float m = mu[off], fo = f[off], eps = noise1(nz, step, s, e, D);
    const float xh = kPrecond ? addcmul_rn(mul_rn(c[0], m), c[1], fo) : fo;
    const float y = add_mul_sep(xh, c[2], eps);
    mu[off] = __fdiv_rn(__fadd_rn(mul_rn(c[3], y), mul_rn(c[4], m)), c[5]);
    if (x_hat_out) x_hat_out[off] = xh;
    if (y_out) y_out[off] = y;
}
__global__ void __launch_bounds__(kThreads)
    k_edm_combine_s(float* __restrict__ x_hat, const float* __restrict__ mu, const float* __restrict__ f, bsi_rowref c_skip, bsi_rowref c_out,
                    const int32_t* __restrict__ step_ptr, int64_t D) {
    const int e = blockIdx.y * kThreads + threadIdx.x;
    if (e >= D) return;
    const int64_t s = blockIdx.x, off = s * D + e;
    const int step = step_ptr ? *step_ptr : 0;
    x_hat[off] = addcmul_rn(mul_rn(rowref_at(c_skip, s, step), mu[off]), rowref_at(c_out, s, step), f[off]);
}
__global__ void __launch_bounds__(kThreads) k_scale_rows_s(float* __restrict__ out, const float* __restrict__ in, bsi_rowref sc, const int32_t* __restrict__ step_ptr,
                                                           int64_t D) {
    const int e = blockIdx.y * kThreads + threadIdx.x;
    if (e >= D) return;
    const int64_t s = blockIdx.x, off = s * D + e;
    out[off] = rowref_at(sc, s, step_ptr ? *step_ptr : 0) * in[off];
}
__global__ void __launch_bounds__(kThreads)
    k_q_sample_s(float* __restrict__ mu, float* __restrict__ model_in, const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ sigma,
                 const float* __restrict__ c_in, bsi_noise nz_arg, int64_t B, int64_t D) {
    const int e = blockIdx.y * kThreads + threadIdx.x;
    if (e >= D) return;
    const uint32_t r32 = blockIdx.x;
    const int64_t r = r32, b = r32 % (uint32_t)B, off = r * D + e;
    const bsi_noise nz = resolve_noise(nz_arg);
    const float m = addcmul_rn(mul_rn(gamma[r], x[b * D + e]), sigma[r], noise1(nz, 0, r, e, D));
    mu[off] = m;
    if (model_in) model_in[off] = c_in[r] * m;
}

// ------------------------------------------------------------------ bucketize (bsi/bsi.py:32-35)
__device__ __forceinline__ int bucket_of(float x, float lo_edge, float dx, int k) {
    // fp32 subtract, true division, truncation toward zero, clamp — same op order as the reference.
    float q = __fdiv_rn(__fsub_rn(x, lo_edge), dx);
    // float->int64 conversion of the reference saturates far out of range; clamp first to stay defined.
    q = fminf(fmaxf(q, -1.0f), (float)k);
    int i = (int)q;  // trunc toward zero
    return min(max(i, 0), k - 1);
}
__global__ void __launch_bounds__(kThreads) k_bucketize(const float* __restrict__ x, int64_t* __restrict__ o64,
                                                        uint8_t* __restrict__ o8, float lo_edge, float dx, int k, int64_t numel) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
        int b = bucket_of(x[i], lo_edge, dx, k);
        if (o64) o64[i] = b;
        if (o8) o8[i] = (uint8_t)b;
    }
}

// ------------------------------------------------------------------ to_8bit_image (bsi/bsi.py:41-48): 4 B read, 1 B written per element
__global__ void __launch_bounds__(kThreads) k_to_u8(uint8_t* __restrict__ out, const float* __restrict__ x, float lo, float span, int64_t numel) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
        // ((x - min) / (max - min) * 255).clamp(0, 255).to(uint8): same op order, truncation toward zero
        float v = __fmul_rn(__fdiv_rn(__fsub_rn(x[i], lo), span), 255.0f);
        out[i] = (uint8_t)(int)fminf(fmaxf(v, 0.0f), 255.0f);
    }
}

// ------------------------------------------------------------------ fp32 -> bf16 cast with pitch
__global__ void __launch_bounds__(kThreads) k_cast_bf16(__nv_bfloat16* __restrict__ out, const float* __restrict__ in,
                                                        int64_t rows, int64_t cols, int64_t ld) {
    const int64_t total = rows * ld;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / ld, c = i - r * ld;
        out[i] = __float2bfloat16(c < cols ? in[r * cols + c] : 0.0f);
    }
}

// ------------------------------------------------------------------ time embedding (bsi/models/pos_emb.py:77-84)
__global__ void __launch_bounds__(kThreads)
    k_time_embed(__nv_bfloat16* __restrict__ o16, float* __restrict__ o32, const float* __restrict__ t,
                 const float* __restrict__ scale, const float* __restrict__ bias, int64_t rows, int size) {
    const int64_t total = rows * size;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / size;
        int j = (int)(i - r * size);
        float v = sinf(addcmul_rn(bias[j], scale[j], t[r]));  // addcmul(bias, scale, t).sin()
        if (o16) o16[i] = __float2bfloat16(v);
        if (o32) o32[i] = v;
    }
}

// ------------------------------------------------------------------ patch-embed operand
// (bsi/models/dit.py:149-153,228-231 + bsi/nn/fourier_features.py:24-36), fused with the c_in scaling.
// One thread per (sample, pixel): reads C floats, writes Cin bf16 (contiguous in the operand row).
// bytes/pixel: 4*C read + 2*Cin written.
__global__ void __launch_bounds__(kThreads)
    k_patch_operand(__nv_bfloat16* __restrict__ A, const float* __restrict__ mu, bsi_rowref scale,
                    const int32_t* __restrict__ step_ptr, int B, int C, int H, int W, int p, int n_min, int n_max, int lda) {
    const int step = step_ptr ? *step_ptr : 0;
    const int nfreq = n_max >= n_min ? n_max - n_min + 1 : 0;
    const int cin = C * (1 + 2 * nfreq);
    const int gw = W / p;
    const int64_t HW = (int64_t)H * W, total = (int64_t)B * HW;
    const int T = (H / p) * gw;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t b = i / HW;
        int pix = (int)(i - b * HW);
        int y = pix / W, x = pix - y * W;
        int tok = (y / p) * gw + (x / p);
        int within = (y % p) * p + (x % p);
        __nv_bfloat16* dst = A + ((int64_t)b * T + tok) * lda + (int64_t)within * cin;
        const float sc = rowref_at(scale, b, step);
        for (int c = 0; c < C; ++c) {
            float v = sc * mu[(b * C + c) * HW + pix];
            dst[c] = __float2bfloat16(v);
            for (int f = 0; f < nfreq; ++f) {
                // coefs = 2*pi*2^n as an fp32 buffer; args = addcmul(offset, coef, x); sin
                float coef = 6.283185307179586f * (float)(1 << (n_min + f));
                float a0 = mul_rn(coef, v);                             // offset 0
                float a1 = addcmul_rn(1.5707963267948966f, coef, v);    // offset fp32(pi/2)
                dst[C + c * 2 * nfreq + 2 * f] = __float2bfloat16(sinf(a0));
                dst[C + c * 2 * nfreq + 2 * f + 1] = __float2bfloat16(sinf(a1));
            }
        }
    }
}
// Tiled variant (patch*patch divides 256): a CTA builds the operand rows of 256 / p^2 consecutive tokens in shared memory -- one
// thread per pixel -- and writes them back as one contiguous block with 16-byte stores (the rows of consecutive tokens are
// adjacent in A, pitch padding included, so there is no separate zero-fill launch).  The per-pixel version above issues 21
// scattered 2-byte stores per thread and six range-reducing sinf calls per channel value (0.39 TB/s, profiles r01).
// Fourier features: sin(2 pi 2^n v) = sinpi(2^(n+1) v), whose argument reduction is exact, evaluated once for n_min; the higher
// octaves follow by the double-angle identities, the pi/2-shifted partner is the cosine.  This is the exact-math value; the
// reference's fp32 angle arithmetic (fp32(2 pi 2^n) * v, rounded at magnitude ~2.4e3) deviates from it by <= 3e-4, a seventh of
// the bf16 rounding step the operand is stored with.
__global__ void __launch_bounds__(kThreads)
    k_patch_operand_tiled(__nv_bfloat16* __restrict__ A, const float* __restrict__ mu, bsi_rowref scale, const int32_t* __restrict__ step_ptr,
                          int total_tokens, int C, int H, int W, int p, int n_min, int nfreq, int lda) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(s_raw);
    const int step = step_ptr ? *step_ptr : 0;
    const int pp = p * p, tok_per_cta = kThreads / pp;
    const int cin = C * (1 + 2 * nfreq), cols = pp * cin;
    const int gw = W / p, T = (H / p) * gw, HW = H * W;
    const int tok0 = blockIdx.x * tok_per_cta;
    const int tl = threadIdx.x / pp, within = threadIdx.x - tl * pp;
    const int token = tok0 + tl;
    // zero the pitch padding of this CTA's rows (columns [cols, lda))
    for (int i = threadIdx.x; i < tok_per_cta * (lda - cols); i += kThreads) {
        const int r = i / (lda - cols);
        tile[r * lda + cols + (i - r * (lda - cols))] = __float2bfloat16(0.0f);
    }
    if (token < total_tokens) {
        const int b = token / T, tok = token - b * T;
        const int gy = tok / gw, gx = tok - gy * gw;
        const int py = within / p, px = within - py * p;
        const int pix = (gy * p + py) * W + gx * p + px;
        const float sc = rowref_at(scale, b, step);
        __nv_bfloat16* dst = tile + tl * lda + within * cin;
        for (int c = 0; c < C; ++c) {
            const float v = sc * mu[((int64_t)b * C + c) * HW + pix];
            dst[c] = __float2bfloat16(v);
            if (nfreq > 0) {
                float sn, cs;
                sincospif(ldexpf(v, n_min + 1), &sn, &cs);
                __nv_bfloat16* o = dst + C + c * 2 * nfreq;
                for (int f = 0; f < nfreq; ++f) {
                    o[2 * f] = __float2bfloat16(sn);
                    o[2 * f + 1] = __float2bfloat16(cs);
                    const float s2 = 2.0f * sn * cs, c2 = (cs - sn) * (cs + sn);
                    sn = s2, cs = c2;
                }
            }
        }
    }
    __syncthreads();
    // rows [tok0, tok0 + tok_per_cta) of A are one contiguous block of tok_per_cta * lda bf16 (lda % 8 == 0: 16-byte rows)
    const int rows = min(tok_per_cta, total_tokens - tok0);
    const int vecs = rows * lda / 8;
    uint4* gdst = reinterpret_cast<uint4*>(A + (int64_t)tok0 * lda);
    const uint4* ssrc = reinterpret_cast<const uint4*>(tile);
    for (int i = threadIdx.x; i < vecs; i += kThreads) gdst[i] = ssrc[i];
}
// zero the pitch padding of the operand (columns [cols, lda)) once per allocation
__global__ void __launch_bounds__(kThreads) k_zero_pad(__nv_bfloat16* __restrict__ A, int64_t rows, int cols, int lda) {
    const int pad = lda - cols;
    const int64_t total = rows * pad;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / pad;
        A[r * lda + cols + (i - r * pad)] = __float2bfloat16(0.0f);
    }
}

}  // namespace bsi

using namespace bsi;

extern "C" {

int bsi_abi_version(void) { return 2; }
long long bsi_launch_counter(void) { return g_launches.load(std::memory_order_relaxed); }
const char* bsi_last_error(void) { return g_err; }
int bsi_device_arch(void) {
    int dev = 0, major = 0, minor = 0;
    BSI_CUDA_OK(cudaGetDevice(&dev));
    BSI_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    BSI_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    return major * 10 + minor;
}

#define BSI_REQUIRE_VEC4(D) BSI_CHECK_ARG((D) > 0, "data numel per sample (%lld) must be positive", (long long)(D))
// element counts that are not a multiple of 4 take the scalar kernels
#define BSI_SCALAR_PATH(D) (((D) & 3) != 0)

int bsi_sample_init(float* mu, const float* sigma0_ptr, bsi_noise noise, int64_t n, int64_t D, void* stream) {
    BSI_CHECK_ARG(mu && sigma0_ptr && n > 0, "bsi_sample_init: null pointer or empty batch");
    BSI_REQUIRE_VEC4(D);
    if (BSI_SCALAR_PATH(D)) {
        BSI_CHECK_ROW_GRID_SCALAR(n, D);
        k_sample_init_s<<<row_grid_scalar(n, D), kThreads, 0, (cudaStream_t)stream>>>(mu, sigma0_ptr, noise, D);
        BSI_LAUNCH_OK("k_sample_init_s");
        return BSI_OK;
    }
    BSI_CHECK_ROW_GRID(n, D);
    k_sample_init<<<row_grid(n, D).grid, kThreads, 0, (cudaStream_t)stream>>>(mu, sigma0_ptr, noise, n, D);
    BSI_LAUNCH_OK("k_sample_init");
    return BSI_OK;
}

int bsi_step_fused(float* mu, const float* f, const float* coef, const int32_t* step_ptr, int32_t step, int32_t precond,
                   bsi_noise noise, float* x_hat_out, float* y_out, int64_t n, int64_t D, void* stream) {
    BSI_CHECK_ARG(mu && f && coef && n > 0, "bsi_step_fused: null pointer or empty batch");
    BSI_REQUIRE_VEC4(D);
    if (BSI_SCALAR_PATH(D)) {
        BSI_CHECK_ROW_GRID_SCALAR(n, D);
        if (precond)
            k_step_fused_s<true><<<row_grid_scalar(n, D), kThreads, 0, (cudaStream_t)stream>>>(mu, f, coef, step_ptr, step, noise, x_hat_out, y_out, D);
        else
            k_step_fused_s<false><<<row_grid_scalar(n, D), kThreads, 0, (cudaStream_t)stream>>>(mu, f, coef, step_ptr, step, noise, x_hat_out, y_out, D);
        BSI_LAUNCH_OK("k_step_fused_s");
        return BSI_OK;
    }
    BSI_CHECK_ROW_GRID(n, D);
    const dim3 grid = row_grid(n, D).grid;
    if (precond)
        k_step_fused<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(mu, f, coef, step_ptr, step, noise, x_hat_out, y_out, n, D);
    else
        k_step_fused<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(mu, f, coef, step_ptr, step, noise, x_hat_out, y_out, n, D);
    BSI_LAUNCH_OK("k_step_fused");
    return BSI_OK;
}

int bsi_step_advance(int32_t* step_ptr, void* stream) {
    BSI_CHECK_ARG(step_ptr, "bsi_step_advance: null step pointer");
    k_step_advance<<<1, 1, 0, (cudaStream_t)stream>>>(step_ptr);
    BSI_LAUNCH_OK("k_step_advance");
    return BSI_OK;
}

int bsi_edm_combine(float* x_hat, const float* mu, const float* f, bsi_rowref c_skip, bsi_rowref c_out, const int32_t* step_ptr,
                    int64_t n, int64_t D, void* stream) {
    BSI_CHECK_ARG(x_hat && mu && f && c_skip.base && c_out.base && n > 0, "bsi_edm_combine: null pointer or empty batch");
    BSI_REQUIRE_VEC4(D);
    if (BSI_SCALAR_PATH(D)) {
        BSI_CHECK_ROW_GRID_SCALAR(n, D);
        k_edm_combine_s<<<row_grid_scalar(n, D), kThreads, 0, (cudaStream_t)stream>>>(x_hat, mu, f, c_skip, c_out, step_ptr, D);
        BSI_LAUNCH_OK("k_edm_combine_s");
        return BSI_OK;
    }
    BSI_CHECK_ROW_GRID(n, D);
    k_edm_combine<<<row_grid(n, D).grid, kThreads, 0, (cudaStream_t)stream>>>(x_hat, mu, f, c_skip, c_out, step_ptr, n, D);
    BSI_LAUNCH_OK("k_edm_combine");
    return BSI_OK;
}

int bsi_scale_rows(float* out, const float* in, bsi_rowref scale, const int32_t* step_ptr, int64_t n, int64_t D, void* stream) {
    BSI_CHECK_ARG(out && in && scale.base && n > 0, "bsi_scale_rows: null pointer or empty batch");
    BSI_REQUIRE_VEC4(D);
    if (BSI_SCALAR_PATH(D)) {
        BSI_CHECK_ROW_GRID_SCALAR(n, D);
        k_scale_rows_s<<<row_grid_scalar(n, D), kThreads, 0, (cudaStream_t)stream>>>(out, in, scale, step_ptr, D);
        BSI_LAUNCH_OK("k_scale_rows_s");
        return BSI_OK;
    }
    BSI_CHECK_ROW_GRID(n, D);
    k_scale_rows<<<row_grid(n, D).grid, kThreads, 0, (cudaStream_t)stream>>>(out, in, scale, step_ptr, n, D);
    BSI_LAUNCH_OK("k_scale_rows");
    return BSI_OK;
}

int bsi_q_sample(float* mu, float* model_in, const float* x, const float* gamma, const float* sigma, const float* c_in,
                 bsi_noise noise, int64_t R, int64_t B, int64_t D, void* stream) {
    BSI_CHECK_ARG(mu && x && gamma && sigma && R > 0 && B > 0, "bsi_q_sample: null pointer or empty batch");
    BSI_CHECK_ARG(!model_in || c_in, "bsi_q_sample: model_in requested without c_in");
    BSI_REQUIRE_VEC4(D);
    BSI_CHECK_ARG(B <= 0x7fffffffLL, "bsi_q_sample: batch too large");
    if (BSI_SCALAR_PATH(D)) {
        BSI_CHECK_ROW_GRID_SCALAR(R, D);
        k_q_sample_s<<<row_grid_scalar(R, D), kThreads, 0, (cudaStream_t)stream>>>(mu, model_in, x, gamma, sigma, c_in, noise, B, D);
        BSI_LAUNCH_OK("k_q_sample_s");
        return BSI_OK;
    }
    BSI_CHECK_ROW_GRID(R, D);
    k_q_sample<<<row_grid(R, D).grid, kThreads, 0, (cudaStream_t)stream>>>(mu, model_in, x, gamma, sigma, c_in, noise, R, B, D);
    BSI_LAUNCH_OK("k_q_sample");
    return BSI_OK;
}

int bsi_bucketize(const float* x, int64_t* out_i64, uint8_t* out_u8, float lo_edge, float dx, int32_t k, int64_t numel,
                  void* stream) {
    if (numel == 0) return BSI_OK;
    BSI_CHECK_ARG(x && (out_i64 || out_u8) && numel > 0 && k > 0, "bsi_bucketize: bad arguments");
    BSI_CHECK_ARG(!out_u8 || k <= 256, "bsi_bucketize: uint8 output needs k <= 256");
    k_bucketize<<<grid_for(numel), kThreads, 0, (cudaStream_t)stream>>>(x, out_i64, out_u8, lo_edge, dx, k, numel);
    BSI_LAUNCH_OK("k_bucketize");
    return BSI_OK;
}

int bsi_to_uint8(uint8_t* out, const float* x, float lo, float hi, int64_t numel, void* stream) {
    if (numel == 0) return BSI_OK;
    BSI_CHECK_ARG(out && x && numel > 0 && hi > lo, "bsi_to_uint8: bad arguments");
    k_to_u8<<<grid_for(numel), kThreads, 0, (cudaStream_t)stream>>>(out, x, lo, hi - lo, numel);
    BSI_LAUNCH_OK("k_to_u8");
    return BSI_OK;
}

int bsi_cast_bf16(void* out_bf16, const float* in, int64_t rows, int64_t cols, int64_t ld_out, void* stream) {
    BSI_CHECK_ARG(out_bf16 && in && rows > 0 && cols > 0 && ld_out >= cols, "bsi_cast_bf16: bad arguments");
    k_cast_bf16<<<grid_for(rows * ld_out), kThreads, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)out_bf16, in, rows, cols, ld_out);
    BSI_LAUNCH_OK("k_cast_bf16");
    return BSI_OK;
}

int bsi_time_embed(void* out_bf16, float* out_f32, const float* t, const float* scale, const float* bias, int64_t rows,
                   int32_t size, void* stream) {
    BSI_CHECK_ARG((out_bf16 || out_f32) && t && scale && bias && rows > 0 && size > 0, "bsi_time_embed: bad arguments");
    k_time_embed<<<grid_for(rows * size), kThreads, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)out_bf16, out_f32, t, scale, bias,
                                                                               rows, size);
    BSI_LAUNCH_OK("k_time_embed");
    return BSI_OK;
}

int bsi_dit_patch_operand(void* A_bf16, const float* mu, bsi_rowref scale, const int32_t* step_ptr, int32_t B, int32_t C,
                          int32_t H, int32_t Wd, int32_t patch, int32_t n_min, int32_t n_max, int32_t lda, void* stream) {
    BSI_CHECK_ARG(A_bf16 && mu && scale.base && B > 0 && C > 0 && patch > 0, "bsi_dit_patch_operand: bad arguments");
    BSI_CHECK_ARG(H % patch == 0 && Wd % patch == 0, "image %dx%d not divisible by patch %d", H, Wd, patch);
    BSI_CHECK_ARG(n_max < n_min || (n_min >= 0 && n_max < 24), "Fourier exponents out of range");
    int nfreq = n_max >= n_min ? n_max - n_min + 1 : 0;
    int cols = patch * patch * C * (1 + 2 * nfreq);
    BSI_CHECK_ARG(lda >= cols, "operand pitch %d < %d", lda, cols);
    const int pp = patch * patch;
    if (256 % pp == 0 && lda % 8 == 0 && (reinterpret_cast<uintptr_t>(A_bf16) & 15) == 0 && (int64_t)B * H * Wd < 0x7fffffffLL &&
        (kThreads / pp) * lda * 2 <= 48 * 1024) {
        const int total_tokens = B * (H / patch) * (Wd / patch), tok_per_cta = kThreads / pp;
        const int smem = tok_per_cta * lda * 2;
        k_patch_operand_tiled<<<(total_tokens + tok_per_cta - 1) / tok_per_cta, kThreads, smem, (cudaStream_t)stream>>>(
            (__nv_bfloat16*)A_bf16, mu, scale, step_ptr, total_tokens, C, H, Wd, patch, n_min, nfreq, lda);
        BSI_LAUNCH_OK("k_patch_operand_tiled");
        return BSI_OK;
    }
    if (lda > cols) {
        int64_t rows = (int64_t)B * (H / patch) * (Wd / patch);
        k_zero_pad<<<grid_for(rows * (lda - cols)), kThreads, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)A_bf16, rows, cols, lda);
        BSI_LAUNCH_OK("k_zero_pad");
    }
    k_patch_operand<<<grid_for((int64_t)B * H * Wd), kThreads, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)A_bf16, mu, scale,
                                                                                           step_ptr, B, C, H, Wd, patch, n_min,
                                                                                           n_max, lda);
    BSI_LAUNCH_OK("k_patch_operand");
    return BSI_OK;
}

}  // extern "C"
