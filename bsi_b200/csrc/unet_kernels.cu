// Kernels of the VDM U-Net denoiser path (bsi/models/vdm_unet.py, bsi/nn/{residual_block,attention,simplified_unet}.py)
// that are not GEMMs: GroupNorm(+SiLU) -> bf16 NHWC operand, input builder (c_in scaling + Fourier features + NCHW->NHWC),
// 1x1 decode convolution back to NCHW, conv-weight packing, and the single-head S=1024, d=128 attention of the centre block.
// Activations are NHWC so that a 3x3 convolution is an implicit GEMM over shifted TMA boxes (gemm_sm100.cu, CONV mode).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace cg = cooperative_groups;

namespace bsi {

// ------------------------------------------------------------------ GroupNorm (+ SiLU) -> bf16   (vdm_unet.py:52, residual_block.py:42-43)
// x fp32 [B][HW][C] -> act bf16 [B][HW][C] = f(GN(x) * gamma + beta), optionally a raw bf16 copy of x (operand of the 1x1 skip conv).
// One CTA per image: pass 1 accumulates per-channel sums (deterministic order), pass 2 re-reads the image from L2.
// HBM-bound: 4 B/elem read + 2 (or 4) B/elem written.
constexpr int kGnThreads = 512;
__global__ void __launch_bounds__(kGnThreads)
    k_groupnorm_act(__nv_bfloat16* __restrict__ act, __nv_bfloat16* __restrict__ raw, const float* __restrict__ x, const float* __restrict__ gamma,
                    const float* __restrict__ beta, int HW, int C, int cpg, float eps, int apply_silu) {
    extern __shared__ float gn_smem[];  // [stripes][C][2] partials, then [C] mean | [C] rstd
    const int quads = C >> 2, stripes = kGnThreads / quads;
    const int cq = threadIdx.x % quads, ps = threadIdx.x / quads;
    const float* xb = x + (size_t)blockIdx.x * HW * C;
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
    if (ps < stripes) {
        for (int p = ps; p < HW; p += stripes) {
            const float4 v = *reinterpret_cast<const float4*>(xb + (size_t)p * C + cq * 4);
            s[0] += v.x, s[1] += v.y, s[2] += v.z, s[3] += v.w;
            ss[0] = fmaf(v.x, v.x, ss[0]), ss[1] = fmaf(v.y, v.y, ss[1]), ss[2] = fmaf(v.z, v.z, ss[2]), ss[3] = fmaf(v.w, v.w, ss[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            gn_smem[(ps * C + cq * 4 + j) * 2] = s[j];
            gn_smem[(ps * C + cq * 4 + j) * 2 + 1] = ss[j];
        }
    }
    __syncthreads();
    float* s_mean = gn_smem + stripes * C * 2;
    float* s_rstd = s_mean + C;
    const int groups = C / cpg;
    if ((int)threadIdx.x < groups) {
        double a = 0.0, b = 0.0;
        for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c)
            for (int st = 0; st < stripes; ++st) a += gn_smem[(st * C + c) * 2], b += gn_smem[(st * C + c) * 2 + 1];
        const double n = (double)HW * cpg, mean = a / n;
        const double var = fmax(b / n - mean * mean, 0.0);
        const float rstd = rsqrtf((float)var + eps);
        for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) s_mean[c] = (float)mean, s_rstd[c] = rstd;
    }
    __syncthreads();
    if (ps < stripes) {
        const float4 g4 = *reinterpret_cast<const float4*>(gamma + cq * 4), b4 = *reinterpret_cast<const float4*>(beta + cq * 4);
        const float m[4] = {s_mean[cq * 4], s_mean[cq * 4 + 1], s_mean[cq * 4 + 2], s_mean[cq * 4 + 3]};
        const float r[4] = {s_rstd[cq * 4], s_rstd[cq * 4 + 1], s_rstd[cq * 4 + 2], s_rstd[cq * 4 + 3]};
        const float g[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
        __nv_bfloat16* ab = act + (size_t)blockIdx.x * HW * C;
        __nv_bfloat16* rb = raw ? raw + (size_t)blockIdx.x * HW * C : nullptr;
        for (int p = ps; p < HW; p += stripes) {
            const float4 v = *reinterpret_cast<const float4*>(xb + (size_t)p * C + cq * 4);
            const float in[4] = {v.x, v.y, v.z, v.w};
            float y[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float t = fmaf((in[j] - m[j]) * r[j], g[j], be[j]);
                y[j] = apply_silu ? t / (1.0f + __expf(-t)) : t;
            }
            *reinterpret_cast<uint2*>(ab + (size_t)p * C + cq * 4) = make_uint2(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]));
            if (rb) *reinterpret_cast<uint2*>(rb + (size_t)p * C + cq * 4) = make_uint2(pack_bf16(in[0], in[1]), pack_bf16(in[2], in[3]));
        }
    }
}

// Cluster variant: an image is split over the 8 CTAs of a thread-block cluster.  Each CTA pulls its contiguous slice
// (HW/8 pixels x C channels, <= 64 KB) into shared memory with one bulk copy, so HBM is read exactly once with the whole
// slice in flight; the per-group sums are combined across the cluster through distributed shared memory (fixed order:
// deterministic) and the normalised activation is produced from the smem copy.  Three CTAs fit an SM, which overlaps one
// CTA's copy with another's reduction / cluster barrier / stores.  4 B read + 2 (or 4) B written per element, nothing else.
constexpr int kGnClusterThreads = 256;
constexpr int kGnClusterSize = 8;
constexpr int kGnSliceBytes = 64 * 1024;
__global__ void __launch_bounds__(kGnClusterThreads, 3)
    k_groupnorm_act_cluster(__nv_bfloat16* __restrict__ act, __nv_bfloat16* __restrict__ raw, const float* __restrict__ x,
                            const float* __restrict__ gamma, const float* __restrict__ beta, int HW, int C, int cpg, float eps, int apply_silu) {
    extern __shared__ __align__(128) uint8_t gn_raw[];
    // [slice fp32] | [stripes][C][2] partials | [C] mean | [C] rstd | [groups][2] doubles (this CTA's group sums) | mbarrier
    cg::cluster_group cluster = cg::this_cluster();
    const int quads = C >> 2, stripes = kGnClusterThreads / quads;
    const int cq = threadIdx.x % quads, ps = threadIdx.x / quads;
    const int rank = (int)cluster.block_rank(), img = blockIdx.x / kGnClusterSize;
    const int pix = HW / kGnClusterSize;  // pixels of this CTA
    const size_t base = ((size_t)img * HW + (size_t)rank * pix) * C;
    float* s_x = reinterpret_cast<float*>(gn_raw);
    float* s_partial = s_x + (size_t)pix * C;
    float* s_mean = s_partial + stripes * C * 2;
    float* s_rstd = s_mean + C;
    double* s_part = reinterpret_cast<double*>(s_rstd + C);
    const int groups = C / cpg;
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_part + 2 * groups);
    if (threadIdx.x == 0) {
        ptx::mbar_init(bar, 1);
        ptx::fence_mbar_init();
        const uint32_t bytes = (uint32_t)pix * C * 4;
        ptx::mbar_arrive_expect_tx(bar, bytes);
        ptx::bulk_load_1d(s_x, x + base, bytes, bar);
    }
    __syncthreads();
    ptx::mbar_wait(bar, 0);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
    for (int p = ps; p < pix; p += stripes) {
        const float4 v = *reinterpret_cast<const float4*>(s_x + (size_t)p * C + cq * 4);
        s[0] += v.x, s[1] += v.y, s[2] += v.z, s[3] += v.w;
        ss[0] = fmaf(v.x, v.x, ss[0]), ss[1] = fmaf(v.y, v.y, ss[1]), ss[2] = fmaf(v.z, v.z, ss[2]), ss[3] = fmaf(v.w, v.w, ss[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s_partial[(ps * C + cq * 4 + j) * 2] = s[j];
        s_partial[(ps * C + cq * 4 + j) * 2 + 1] = ss[j];
    }
    __syncthreads();
    if ((int)threadIdx.x < groups) {
        double a = 0.0, b = 0.0;
        for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c)
            for (int st = 0; st < stripes; ++st) a += s_partial[(st * C + c) * 2], b += s_partial[(st * C + c) * 2 + 1];
        s_part[threadIdx.x * 2] = a, s_part[threadIdx.x * 2 + 1] = b;
    }
    cluster.sync();
    if ((int)threadIdx.x < groups) {
        double a = 0.0, b = 0.0;
        for (int r = 0; r < kGnClusterSize; ++r) {
            const double* remote = cluster.map_shared_rank(s_part, r);
            a += remote[threadIdx.x * 2], b += remote[threadIdx.x * 2 + 1];
        }
        const double n = (double)HW * cpg, mean = a / n;
        const double var = fmax(b / n - mean * mean, 0.0);
        const float rstd = rsqrtf((float)var + eps);
        for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) s_mean[c] = (float)mean, s_rstd[c] = rstd;
    }
    cluster.sync();  // also keeps every CTA's partials alive until all peers have read them
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + cq * 4), b4 = *reinterpret_cast<const float4*>(beta + cq * 4);
    const float m[4] = {s_mean[cq * 4], s_mean[cq * 4 + 1], s_mean[cq * 4 + 2], s_mean[cq * 4 + 3]};
    const float r[4] = {s_rstd[cq * 4], s_rstd[cq * 4 + 1], s_rstd[cq * 4 + 2], s_rstd[cq * 4 + 3]};
    const float g[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
    __nv_bfloat16* ab = act + base;
    __nv_bfloat16* rb = raw ? raw + base : nullptr;
#pragma unroll 4
    for (int p = ps; p < pix; p += stripes) {
        const float4 v = *reinterpret_cast<const float4*>(s_x + (size_t)p * C + cq * 4);
        const float in[4] = {v.x, v.y, v.z, v.w};
        float y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float t = fmaf((in[j] - m[j]) * r[j], g[j], be[j]);
            y[j] = apply_silu ? __fdividef(t, 1.0f + __expf(-t)) : t;
        }
        const size_t off = (size_t)p * C + cq * 4;
        *reinterpret_cast<uint2*>(ab + off) = make_uint2(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]));
        if (rb) *reinterpret_cast<uint2*>(rb + off) = make_uint2(pack_bf16(in[0], in[1]), pack_bf16(in[2], in[3]));
    }
}

// GroupNorm + SiLU from PRE-COMPUTED statistics: the convolution that produced x has left, per 128-pixel tile, the sums and sums of
// squares of every group of 4 channels (bsi_conv_args.gn_partial).  What remains is a pure streaming pass -- 4 B read, 2 (or 4) B
// written per element, no cluster, no shared-memory staging, no second look at x: every CTA first folds its image's
// HW/128 x 64 partial sums (fp64, fixed order: deterministic) into per-channel mean / rstd, then normalises a 64-pixel strip.
// cpg = 4 (GroupNorm(32, 128)) or 8 (one source of GroupNorm(32, 256) over cat(x, skip): adjacent 4-channel groups are merged).
constexpr int kGnApplyThreads = 256, kGnApplyPixels = 64;
__global__ void __launch_bounds__(kGnApplyThreads)
    k_groupnorm_apply(__nv_bfloat16* __restrict__ act, __nv_bfloat16* __restrict__ raw, const float* __restrict__ x, const float* __restrict__ partial,
                      const float* __restrict__ gamma, const float* __restrict__ beta, int HW, int cpg, float eps, int apply_silu) {
    constexpr int C = 128;
    __shared__ float s_mean[C], s_rstd[C];
    const int img = blockIdx.y, strip = blockIdx.x, tiles = HW / 128;
    if (threadIdx.x < 32) {
        const int g4 = threadIdx.x;  // 4-channel group
        double a = 0.0, b = 0.0;
        const float* pp = partial + (size_t)img * tiles * 64;
        if (cpg == 4) {
            for (int t = 0; t < tiles; ++t) a += pp[t * 64 + 2 * g4], b += pp[t * 64 + 2 * g4 + 1];
        } else {  // cpg == 8: groups (g4 & ~1) and (g4 | 1) form one group
            const int g0 = g4 & ~1;
            for (int t = 0; t < tiles; ++t)
                a += (double)pp[t * 64 + 2 * g0] + (double)pp[t * 64 + 2 * g0 + 2], b += (double)pp[t * 64 + 2 * g0 + 1] + (double)pp[t * 64 + 2 * g0 + 3];
        }
        const double n = (double)HW * cpg, mean = a / n;
        const double var = fmax(b / n - mean * mean, 0.0);
        const float rstd = rsqrtf((float)var + eps);
#pragma unroll
        for (int j = 0; j < 4; ++j) s_mean[g4 * 4 + j] = (float)mean, s_rstd[g4 * 4 + j] = rstd;
    }
    __syncthreads();
    const int cq = threadIdx.x & 31, ps = threadIdx.x >> 5;  // 32 channel quads x 8 pixel stripes
    const float4 g4v = *reinterpret_cast<const float4*>(gamma + cq * 4), b4v = *reinterpret_cast<const float4*>(beta + cq * 4);
    const float m[4] = {s_mean[cq * 4], s_mean[cq * 4 + 1], s_mean[cq * 4 + 2], s_mean[cq * 4 + 3]};
    const float r[4] = {s_rstd[cq * 4], s_rstd[cq * 4 + 1], s_rstd[cq * 4 + 2], s_rstd[cq * 4 + 3]};
    // fold the affine into one FMA per element: y = x * (rstd*gamma) + (beta - mean*rstd*gamma) ... kept as the reference's op order
    const float g[4] = {g4v.x, g4v.y, g4v.z, g4v.w}, be[4] = {b4v.x, b4v.y, b4v.z, b4v.w};
    const size_t base = ((size_t)img * HW + (size_t)strip * kGnApplyPixels) * C;
    float4 v[kGnApplyPixels / 8];
#pragma unroll
    for (int i = 0; i < kGnApplyPixels / 8; ++i) v[i] = __ldcs(reinterpret_cast<const float4*>(x + base + (size_t)(ps + 8 * i) * C + cq * 4));
#pragma unroll
    for (int i = 0; i < kGnApplyPixels / 8; ++i) {
        const float in[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
        float y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float t = fmaf((in[j] - m[j]) * r[j], g[j], be[j]);
            y[j] = apply_silu ? __fdividef(t, 1.0f + __expf(-t)) : t;
        }
        const size_t off = base + (size_t)(ps + 8 * i) * C + cq * 4;
        *reinterpret_cast<uint2*>(act + off) = make_uint2(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]));
        if (raw) *reinterpret_cast<uint2*>(raw + off) = make_uint2(pack_bf16(in[0], in[1]), pack_bf16(in[2], in[3]));
    }
}

// ------------------------------------------------------------------ U-Net input operand   (vdm_unet.py:95-98, fourier_features.py:24-36)
// bf16 NHWC [B][H][W][Cpad]: channels [0,C) = scale*mu, then for each channel (n, {sin,cos}) Fourier features, zero padding up to Cpad.
__global__ void __launch_bounds__(256) k_unet_input(__nv_bfloat16* __restrict__ out, const float* __restrict__ mu, bsi_rowref scale,
                                                    const int32_t* __restrict__ step_ptr, int B, int C, int HW, int n_min, int n_max, int cpad) {
    const int step = step_ptr ? *step_ptr : 0;
    const int nfreq = n_max >= n_min ? n_max - n_min + 1 : 0;
    const int64_t total = (int64_t)B * HW;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / HW;
        const int pix = (int)(i - b * HW);
        __nv_bfloat16* dst = out + i * cpad;
        const float sc = rowref_at(scale, b, step);
        int used = C * (1 + 2 * nfreq);
        for (int c = 0; c < C; ++c) {
            const float v = sc * mu[(b * C + c) * HW + pix];
            dst[c] = __float2bfloat16(v);
            for (int f = 0; f < nfreq; ++f) {
                const float coef = 6.283185307179586f * (float)(1 << (n_min + f));
                dst[C + c * 2 * nfreq + 2 * f] = __float2bfloat16(sinf(__fmul_rn(coef, v)));
                dst[C + c * 2 * nfreq + 2 * f + 1] = __float2bfloat16(sinf(__fmaf_rn(coef, v, 1.5707963267948966f)));
            }
        }
        for (int c = used; c < cpad; ++c) dst[c] = __float2bfloat16(0.0f);
    }
}

// ------------------------------------------------------------------ decode: 1x1 conv C -> Cout (tiny) + NHWC -> NCHW   (vdm_unet.py:72,100)
// one warp per pixel: coalesced 4*C-byte read, Cout dot products reduced by shuffles.  Reads the fp32 stream once.
__global__ void __launch_bounds__(256) k_unet_decode(float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ bias, int64_t pixels, int HW, int C, int cout) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < pixels; p += warps) {
        const float* xp = x + p * C;
        const int64_t b = p / HW;
        const int pix = (int)(p - b * HW);
        for (int o = 0; o < cout; ++o) {
            float acc = 0.0f;
            for (int c = lane; c < C; c += 32) acc = fmaf(xp[c], w[o * C + c], acc);
            acc = warp_sum(acc);
            if (lane == 0) out[(b * cout + o) * HW + pix] = acc + bias[o];
        }
    }
}

// ------------------------------------------------------------------ conv weight packing: fp32 [N][Cin][kh][kw] -> bf16 [N][taps][Cpad]
__global__ void __launch_bounds__(256) k_pack_conv_weight(__nv_bfloat16* __restrict__ out, const float* __restrict__ w, int N, int cin, int taps,
                                                          int cpad, int c_offset, int ctotal) {
    // writes channels [c_offset, c_offset + cpad) of a [N][taps][ctotal] layout (ctotal = sum of the sources' padded channels)
    const int64_t total = (int64_t)N * taps * cpad;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cpad), tap = (int)((i / cpad) % taps), n = (int)(i / ((int64_t)cpad * taps));
        const float v = c < cin ? w[((int64_t)n * cin + c) * taps + tap] : 0.0f;
        out[((int64_t)n * taps + tap) * ctotal + c_offset + c] = __float2bfloat16(v);
    }
}

// ------------------------------------------------------------------ attention, 1 head, d = 128   (bsi/nn/attention.py:32-41)
// qkv bf16 [B*T][3*128] columns (q | k | v) -> out bf16 [B*T][128].  Flash-style: 128 query rows per CTA (8 warps x 16 rows),
// K/V streamed in 64-key blocks through a cp.async double buffer, online softmax in fp32, warp-level mma.sync m16n8k16.
// 3 % of the U-Net's flops; tensor-bound 4*T*T*128 flop per image.
constexpr int kA2Threads = 256, kA2D = 128, kA2Q = 128, kA2KB = 64;
__device__ __forceinline__ uint32_t sw256(int row, int chunk) { return (uint32_t)(row * 256 + ((chunk ^ (row & 7)) << 4)); }
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kA2Threads, 1) k_attention_d128(__nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ qkv, int T, float scale_log2) {
    extern __shared__ __align__(128) uint8_t a2_smem[];
    const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(a2_smem);
    const uint32_t sK = sQ + kA2Q * 256, sV = sK + 2 * kA2KB * 256;  // K and V: two 64-key buffers each
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * kA2Q, b = blockIdx.y;
    const size_t ld = 3 * kA2D;
    const __nv_bfloat16* base = qkv + (size_t)b * T * ld;

    auto load_kv = [&](int blk, int bufi) {
        for (int i = tid; i < kA2KB * 16; i += kA2Threads) {
            const int r = i >> 4, c = i & 15;
            const __nv_bfloat16* src = base + (size_t)(blk * kA2KB + r) * ld + c * 8;
            cp16(sK + bufi * kA2KB * 256 + sw256(r, c), src + kA2D);
            cp16(sV + bufi * kA2KB * 256 + sw256(r, c), src + 2 * kA2D);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int i = tid; i < kA2Q * 16; i += kA2Threads) {
        const int r = i >> 4, c = i & 15;
        cp16(sQ + sw256(r, c), base + (size_t)(q0 + r) * ld + c * 8);
    }
    load_kv(0, 0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int r0 = warp * 16;
    uint32_t qf[8][4];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8, chunk = ks * 2 + (lane >> 4);
        ldsm4(sQ + sw256(row, chunk), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
    float o[16][4];
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.0f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.0f, 0.0f};

    const int nblk = T / kA2KB;
    for (int blk = 0; blk < nblk; ++blk) {
        const int cur = blk & 1;
        if (blk + 1 < nblk) load_kv(blk + 1, cur ^ 1);  // prefetch behind this block's math
        const uint32_t kb = sK + cur * kA2KB * 256, vb = sV + cur * kA2KB * 256;
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                const int key = np * 16 + (lane & 7) + (lane >> 4) * 8, chunk = ks * 2 + ((lane >> 3) & 1);
                uint32_t b0, b1, b2, b3;
                ldsm4(kb + sw256(key, chunk), b0, b1, b2, b3);
                mma16816(s[2 * np], qf[ks], b0, b1);
                mma16816(s[2 * np + 1], qf[ks], b2, b3);
            }
        }
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
        }
        float corr[2], msc[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
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
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            o[j][0] *= corr[0], o[j][1] *= corr[0];
            o[j][2] *= corr[1], o[j][3] *= corr[1];
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const uint32_t pa[4] = {pack_bf16(s[2 * kk][0], s[2 * kk][1]), pack_bf16(s[2 * kk][2], s[2 * kk][3]),
                                    pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
            for (int dn = 0; dn < 8; ++dn) {
                const int key = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, chunk = dn * 2 + (lane >> 4);
                uint32_t b0, b1, b2, b3;
                ldsm4t(vb + sw256(key, chunk), b0, b1, b2, b3);
                mma16816(o[2 * dn], pa, b0, b1);
                mma16816(o[2 * dn + 1], pa, b2, b3);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // next block landed; everyone is done with the current buffers
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
    const int g = lane >> 2, t4 = lane & 3;
    __nv_bfloat16* ob = out + ((size_t)b * T + q0 + r0) * kA2D;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        *reinterpret_cast<uint32_t*>(ob + (size_t)g * kA2D + j * 8 + 2 * t4) = pack_bf16(o[j][0] * inv0, o[j][1] * inv0);
        *reinterpret_cast<uint32_t*>(ob + (size_t)(g + 8) * kA2D + j * 8 + 2 * t4) = pack_bf16(o[j][2] * inv1, o[j][3] * inv1);
    }
}

}  // namespace bsi

using namespace bsi;

extern "C" {

int bsi_groupnorm_act_bf16(void* act_bf16, void* raw_bf16, const float* x, const float* gamma, const float* beta, int32_t B, int32_t HW, int32_t C,
                           int32_t channels_per_group, float eps, int32_t apply_silu, void* stream) {
    BSI_CHECK_ARG(act_bf16 && x && gamma && beta && B > 0 && HW > 0, "bsi_groupnorm_act_bf16: bad arguments");
    BSI_CHECK_ARG(C % 4 == 0 && C >= 32 && C <= 512 && kGnThreads % (C / 4) == 0 && channels_per_group > 0 && C % channels_per_group == 0 &&
                      C / channels_per_group <= kGnThreads,
                  "bsi_groupnorm_act_bf16: unsupported channel count %d / group size %d", C, channels_per_group);
    const int stripes = kGnThreads / (C / 4);
    const int smem = (stripes * C * 2 + 2 * C) * (int)sizeof(float);
    // cluster path: the image split over 2, 4 or 8 CTAs fits the threads' registers (<= 16 float4 each)
    // cluster path: the image splits into 8 contiguous slices of <= 64 KB that each CTA stages in shared memory
    static const bool use_cluster = [] { const char* e = getenv("BSI_GN_CLUSTER"); return !(e && e[0] == '0'); }();
    const int cstripes = kGnClusterThreads / (C / 4);
    if (use_cluster && kGnClusterThreads % (C / 4) == 0 && C / channels_per_group <= kGnClusterThreads && HW % kGnClusterSize == 0 &&
        (int64_t)(HW / kGnClusterSize) * C * 4 <= kGnSliceBytes && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const int slice = (HW / kGnClusterSize) * C * 4;
        const int smem_c = slice + (cstripes * C * 2 + 2 * C) * (int)sizeof(float) + 2 * (C / channels_per_group) * (int)sizeof(double) + 16;
        BSI_ENSURE_SMEM(k_groupnorm_act_cluster, smem_c);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)B * kGnClusterSize), cfg.blockDim = dim3(kGnClusterThreads), cfg.dynamicSmemBytes = smem_c, cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kGnClusterSize, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr, cfg.numAttrs = 1;
        BSI_CUDA_OK(cudaLaunchKernelEx(&cfg, k_groupnorm_act_cluster, (__nv_bfloat16*)act_bf16, (__nv_bfloat16*)raw_bf16, x, gamma, beta, (int)HW, (int)C,
                                       (int)channels_per_group, eps, (int)apply_silu));
        BSI_LAUNCH_OK("k_groupnorm_act_cluster");
        return BSI_OK;
    }
    if (smem > 48 * 1024) BSI_ENSURE_SMEM(k_groupnorm_act, smem);
    k_groupnorm_act<<<B, kGnThreads, smem, (cudaStream_t)stream>>>((__nv_bfloat16*)act_bf16, (__nv_bfloat16*)raw_bf16, x, gamma, beta, HW, C,
                                                                  channels_per_group, eps, apply_silu);
    BSI_LAUNCH_OK("k_groupnorm_act");
    return BSI_OK;
}

int bsi_groupnorm_apply_bf16(void* act_bf16, void* raw_bf16, const float* x, const float* partial, const float* gamma, const float* beta, int32_t B,
                             int32_t HW, int32_t C, int32_t channels_per_group, float eps, int32_t apply_silu, void* stream) {
    BSI_CHECK_ARG(act_bf16 && x && partial && gamma && beta && B > 0 && HW > 0, "bsi_groupnorm_apply_bf16: bad arguments");
    BSI_CHECK_ARG(C == 128 && (channels_per_group == 4 || channels_per_group == 8) && HW % 128 == 0 && B <= 65535,
                  "bsi_groupnorm_apply_bf16: implemented for C = 128, groups of 4 or 8 channels, H*W %% 128 == 0 (got C=%d cpg=%d HW=%d)", C,
                  channels_per_group, HW);
    k_groupnorm_apply<<<dim3(HW / kGnApplyPixels, B), kGnApplyThreads, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)act_bf16, (__nv_bfloat16*)raw_bf16, x, partial,
                                                                                                 gamma, beta, HW, channels_per_group, eps, apply_silu);
    BSI_LAUNCH_OK("k_groupnorm_apply");
    return BSI_OK;
}

int bsi_unet_input_bf16(void* out_bf16, const float* mu, bsi_rowref scale, const int32_t* step_ptr, int32_t B, int32_t C, int32_t HW, int32_t n_min,
                        int32_t n_max, int32_t cpad, void* stream) {
    const int nfreq = n_max >= n_min ? n_max - n_min + 1 : 0;
    BSI_CHECK_ARG(out_bf16 && mu && scale.base && B > 0 && C > 0 && HW > 0 && cpad >= C * (1 + 2 * nfreq) && cpad % 8 == 0,
                  "bsi_unet_input_bf16: bad arguments");
    const int64_t total = (int64_t)B * HW;
    const int64_t need = (total + 255) / 256, cap = (int64_t)sm_count() * 8;
    k_unet_input<<<(unsigned)(need < cap ? need : cap), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)out_bf16, mu, scale, step_ptr, B, C, HW, n_min,
                                                                                         n_max, cpad);
    BSI_LAUNCH_OK("k_unet_input");
    return BSI_OK;
}

int bsi_unet_decode(float* out, const float* x, const float* w, const float* bias, int32_t B, int32_t HW, int32_t C, int32_t cout, void* stream) {
    BSI_CHECK_ARG(out && x && w && bias && B > 0 && HW > 0 && C > 0 && cout > 0, "bsi_unet_decode: bad arguments");
    const int64_t pixels = (int64_t)B * HW;
    const int64_t need = (pixels + 7) / 8, cap = (int64_t)sm_count() * 8;
    k_unet_decode<<<(unsigned)(need < cap ? need : cap), 256, 0, (cudaStream_t)stream>>>(out, x, w, bias, pixels, HW, C, cout);
    BSI_LAUNCH_OK("k_unet_decode");
    return BSI_OK;
}

int bsi_pack_conv_weight(void* out_bf16, const float* w, int32_t N, int32_t cin, int32_t taps, int32_t cpad, int32_t c_offset, int32_t ctotal,
                         void* stream) {
    BSI_CHECK_ARG(out_bf16 && w && N > 0 && cin > 0 && (taps == 1 || taps == 9) && cpad >= cin && c_offset >= 0 && c_offset + cpad <= ctotal,
                  "bsi_pack_conv_weight: bad arguments");
    const int64_t total = (int64_t)N * taps * cpad;
    const int64_t need = (total + 255) / 256, cap = (int64_t)sm_count() * 8;
    k_pack_conv_weight<<<(unsigned)(need < cap ? need : cap), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)out_bf16, w, N, cin, taps, cpad, c_offset,
                                                                                               ctotal);
    BSI_LAUNCH_OK("k_pack_conv_weight");
    return BSI_OK;
}

int bsi_attention_d128_bf16(void* out_bf16, const void* qkv_bf16, int32_t B, int32_t T, void* stream) {
    BSI_CHECK_ARG(out_bf16 && qkv_bf16 && B > 0, "bsi_attention_d128_bf16: bad arguments");
    if (T % kA2Q != 0) {
        set_error("bsi_attention_d128_bf16: sequence length %d must be a multiple of 128", T);
        return BSI_ERR_UNSUPPORTED;
    }
    const int smem = (kA2Q + 4 * kA2KB) * 256;
    BSI_ENSURE_SMEM(k_attention_d128, smem);
    const float scale_log2 = 1.4426950408889634f / sqrtf((float)kA2D);
    k_attention_d128<<<dim3(T / kA2Q, B), kA2Threads, smem, (cudaStream_t)stream>>>((__nv_bfloat16*)out_bf16, (const __nv_bfloat16*)qkv_bf16, T, scale_log2);
    BSI_LAUNCH_OK("k_attention_d128");
    return BSI_OK;
}

}  // extern "C"
