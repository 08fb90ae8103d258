// Elementwise / reduction kernels of the DiT training path (backward of bsi/models/dit.py:50-55,87-103).
// All HBM-bound; the GEMMs around them are bsi_gemm_bf16 (forward, data gradient with transposed weights) and
// bsi_gemm_wgrad_bf16 (weight gradient).
//   gate_residual            x_out = x + gate[b] * branch                 (torch.addcmul(x, gate, branch), dit.py:93-102)
//   gate_residual_backward   dbranch = gate[b] * dx ; dgate[b] = sum_t dx * branch ; per-sample column sums of dbranch (bias gradient)
//   colsum_bf16              partial column sums of a bf16 matrix (bias gradients of the other linears)
//   gelu / gelu_backward     nn.GELU(approximate="tanh") on the bf16 pre-activation (dit.py:75)
//   layernorm_mod_backward   backward of modulate(LayerNorm(x), shift, scale) (dit.py:50-55) and of the affine decoder LayerNorm (:164):
//                            dx += rstd * (g - mean(g) - xhat * mean(g * xhat)), g = da * (1 + scale)   [or da * gamma]
//                            dscale[b] = sum_t da * xhat, dshift[b] = sum_t da   (per-CTA partial sums, fixed order)
#include <cstdlib>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace bsi {

constexpr int kTrThreads = 256;

__device__ __forceinline__ float2 bf16x2_to_float2(uint32_t v) { return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)); }

// ------------------------------------------------------------------ x += gate * branch
__global__ void __launch_bounds__(kTrThreads) k_gate_residual(float* x_out, const float* x, const __nv_bfloat16* __restrict__ br, bsi_rowref gate, int T, int64_t M, int D) {
    const int64_t quads = M * (D / 4);
    for (int64_t i = (int64_t)blockIdx.x * kTrThreads + threadIdx.x; i < quads; i += (int64_t)gridDim.x * kTrThreads) {
        const int64_t row = i / (D / 4);
        const int c = (int)(i - row * (D / 4)) * 4;
        const float4 g = gate.base ? *reinterpret_cast<const float4*>(rowref_ptr(gate, row / T, 0) + c) : make_float4(1.f, 1.f, 1.f, 1.f);
        const uint2 b = *reinterpret_cast<const uint2*>(br + row * D + c);
        const float2 b0 = bf16x2_to_float2(b.x), b1 = bf16x2_to_float2(b.y);
        float4 v = *reinterpret_cast<const float4*>(x + row * D + c);
        v.x = fmaf(g.x, b0.x, v.x), v.y = fmaf(g.y, b0.y, v.y), v.z = fmaf(g.z, b1.x, v.z), v.w = fmaf(g.w, b1.y, v.w);
        *reinterpret_cast<float4*>(x_out + row * D + c) = v;
    }
}

// ------------------------------------------------------------------ x_out = x + gate * branch, then LayerNorm + modulation of x_out -> bf16
// The residual update at the end of one branch and the LayerNorm that opens the next one (dit.py:93-102) in one pass: the row
// is read once, kept in registers for the statistics, and written once as fp32 (saved for the backward) and once as bf16.
template <int NV>
__global__ void __launch_bounds__(kTrThreads, 2)
    k_gate_residual_layernorm(__nv_bfloat16* __restrict__ out, float* __restrict__ x_out, const float* __restrict__ x, const __nv_bfloat16* __restrict__ br,
                              bsi_rowref gate, bsi_rowref shift, bsi_rowref scale, const float* __restrict__ gamma, const float* __restrict__ beta,
                              int rows_per_sample, int64_t M, float eps, uint32_t drop_thresh, uint32_t drop_seed, float drop_inv) {
    constexpr int dim = 128 * NV;
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * (kTrThreads / 32);
    for (int64_t row = (int64_t)blockIdx.x * (kTrThreads / 32) + (threadIdx.x >> 5); row < M; row += warps_total) {
        const int64_t sample = row / rows_per_sample;
        const float4* xr = reinterpret_cast<const float4*>(x + row * dim);
        const uint2* brr = reinterpret_cast<const uint2*>(br + row * dim);
        const float4* gr = gate.base ? reinterpret_cast<const float4*>(rowref_ptr(gate, sample, 0)) : nullptr;
        float4* xo = reinterpret_cast<float4*>(x_out + row * dim);
        float4 v[NV];
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[i] = xr[lane + 32 * i];
            const uint2 b = brr[lane + 32 * i];
            const float2 b0 = bf16x2_to_float2(b.x), b1 = bf16x2_to_float2(b.y);
            const float4 g = gr ? gr[lane + 32 * i] : make_float4(1.f, 1.f, 1.f, 1.f);
            v[i].x = fmaf(g.x, b0.x, v[i].x), v[i].y = fmaf(g.y, b0.y, v[i].y), v[i].z = fmaf(g.z, b1.x, v[i].z), v[i].w = fmaf(g.w, b1.y, v[i].w);
            xo[lane + 32 * i] = v[i];
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        const float mean = warp_sum(s) * (1.0f / dim);
        float ss = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            ss += (a * a + b * b) + (c * c + d * d);
        }
        const float rstd = rsqrtf(warp_sum(ss) * (1.0f / dim) + eps);
        const float4 *p_mul, *p_add;
        if (gamma) {
            p_mul = reinterpret_cast<const float4*>(gamma), p_add = reinterpret_cast<const float4*>(beta);
        } else {
            p_mul = reinterpret_cast<const float4*>(rowref_ptr(scale, sample, 0)), p_add = reinterpret_cast<const float4*>(rowref_ptr(shift, sample, 0));
        }
        const float one = gamma ? 0.0f : 1.0f;
        uint2* o = reinterpret_cast<uint2*>(out + row * dim);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float4 m = p_mul[lane + 32 * i], a = p_add[lane + 32 * i];
            float y0 = fmaf((v[i].x - mean) * rstd, m.x + one, a.x), y1 = fmaf((v[i].y - mean) * rstd, m.y + one, a.y);
            float y2 = fmaf((v[i].z - mean) * rstd, m.z + one, a.z), y3 = fmaf((v[i].w - mean) * rstd, m.w + one, a.w);
            if (drop_thresh) {
                const uint32_t e = (uint32_t)row * dim + (lane + 32 * i) * 4;
                bool k0, k1, k2, k3;  // e is a multiple of 4: two hashes for four elements
                dropout_keep_pair(drop_seed, e, drop_thresh, k0, k1);
                dropout_keep_pair(drop_seed, e + 2, drop_thresh, k2, k3);
                y0 = k0 ? y0 * drop_inv : 0.0f, y1 = k1 ? y1 * drop_inv : 0.0f;
                y2 = k2 ? y2 * drop_inv : 0.0f, y3 = k3 ? y3 * drop_inv : 0.0f;
            }
            o[lane + 32 * i] = make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3));
        }
    }
}

// The same kernel with the rows staged through shared memory by the bulk-copy engine (round 2).  The register version above keeps
// only 16 warps per SM resident and each of them alternates between waiting for its row and computing on it: ncu shows nothing
// saturated (DRAM 46 %, L2 25 %, 24 % of the warp slots) -- 4.0 TB/s.  Here every warp owns two row buffers (x row + branch row,
// 6 KB each); lane 0 issues `cp.async.bulk` for the NEXT row before the warp starts on the current one, completion arrives on a
// per-warp mbarrier, so 16 warps keep 96 KB of loads in flight per SM at all times, independent of how the compiler schedules the
// math.  A CTA walks a CONTIGUOUS range of rows (balanced to one row over the persistent grid), so the per-sample gate / shift /
// scale vectors it reads through L1 belong to one or two samples.  Arithmetic and its order are those of the register version:
// the results are bit-identical.
constexpr int kPipeWarps = kTrThreads / 32;
constexpr int kLnbWarps = 4;  // warps per CTA of the backward row kernels (their rows carry 10 / 6 bytes per element and register partial sums)
template <int NV>
__global__ void __launch_bounds__(kTrThreads, 2)
    k_gate_residual_layernorm_pipe(__nv_bfloat16* __restrict__ out, float* __restrict__ x_out, const float* __restrict__ x, const __nv_bfloat16* __restrict__ br,
                                   bsi_rowref gate, bsi_rowref shift, bsi_rowref scale, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   int rows_per_sample, int64_t M, float eps, uint32_t drop_thresh, uint32_t drop_seed, float drop_inv) {
    constexpr int dim = 128 * NV;
    constexpr uint32_t kXBytes = dim * 4, kBrBytes = dim * 2, kRowBytes = kXBytes + kBrBytes;
    extern __shared__ __align__(128) uint8_t pipe_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* my = pipe_smem + (size_t)warp * 2 * kRowBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(pipe_smem + (size_t)kPipeWarps * 2 * kRowBytes) + warp * 2;
    if (lane == 0) {
        ptx::mbar_init(bars, 1);
        ptx::mbar_init(bars + 1, 1);
        ptx::fence_mbar_init();
    }
    __syncwarp();
    const int64_t per = M / gridDim.x, rem = M % gridDim.x;
    const int64_t r_begin = blockIdx.x * per + (blockIdx.x < rem ? blockIdx.x : rem), r_end = r_begin + per + (blockIdx.x < rem ? 1 : 0);
    auto issue = [&](int64_t row, int buf) {
        ptx::mbar_arrive_expect_tx(bars + buf, kRowBytes);
        ptx::bulk_load_1d(my + buf * kRowBytes, x + row * dim, kXBytes, bars + buf);
        ptx::bulk_load_1d(my + buf * kRowBytes + kXBytes, br + row * dim, kBrBytes, bars + buf);
    };
    int64_t row = r_begin + warp;
    if (row < r_end && lane == 0) issue(row, 0);
    for (int k = 0; row < r_end; row += kPipeWarps, ++k) {
        const int buf = k & 1;
        if (row + kPipeWarps < r_end && lane == 0) issue(row + kPipeWarps, buf ^ 1);  // that buffer was consumed in iteration k - 1 (warp-synchronised below)
        const int64_t sample = row / rows_per_sample;
        const float4* gr = gate.base ? reinterpret_cast<const float4*>(rowref_ptr(gate, sample, 0)) : nullptr;
        float4* xo = reinterpret_cast<float4*>(x_out + row * dim);
        ptx::mbar_wait(bars + buf, (k >> 1) & 1);
        const float4* xr = reinterpret_cast<const float4*>(my + buf * kRowBytes);
        const uint2* brr = reinterpret_cast<const uint2*>(my + buf * kRowBytes + kXBytes);
        float4 v[NV];
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[i] = xr[lane + 32 * i];
            const uint2 b = brr[lane + 32 * i];
            const float2 b0 = bf16x2_to_float2(b.x), b1 = bf16x2_to_float2(b.y);
            const float4 g = gr ? __ldg(gr + lane + 32 * i) : make_float4(1.f, 1.f, 1.f, 1.f);
            v[i].x = fmaf(g.x, b0.x, v[i].x), v[i].y = fmaf(g.y, b0.y, v[i].y), v[i].z = fmaf(g.z, b1.x, v[i].z), v[i].w = fmaf(g.w, b1.y, v[i].w);
            xo[lane + 32 * i] = v[i];
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        const float mean = warp_sum(s) * (1.0f / dim);
        float ss = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            ss += (a * a + b * b) + (c * c + d * d);
        }
        const float rstd = rsqrtf(warp_sum(ss) * (1.0f / dim) + eps);
        const float4 *p_mul, *p_add;
        if (gamma) {
            p_mul = reinterpret_cast<const float4*>(gamma), p_add = reinterpret_cast<const float4*>(beta);
        } else {
            p_mul = reinterpret_cast<const float4*>(rowref_ptr(scale, sample, 0)), p_add = reinterpret_cast<const float4*>(rowref_ptr(shift, sample, 0));
        }
        const float one = gamma ? 0.0f : 1.0f;
        uint2* o = reinterpret_cast<uint2*>(out + row * dim);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float4 m = __ldg(p_mul + lane + 32 * i), a = __ldg(p_add + lane + 32 * i);
            float y0 = fmaf((v[i].x - mean) * rstd, m.x + one, a.x), y1 = fmaf((v[i].y - mean) * rstd, m.y + one, a.y);
            float y2 = fmaf((v[i].z - mean) * rstd, m.z + one, a.z), y3 = fmaf((v[i].w - mean) * rstd, m.w + one, a.w);
            if (drop_thresh) {
                const uint32_t e = (uint32_t)row * dim + (lane + 32 * i) * 4;
                bool k0, k1, k2, k3;  // e is a multiple of 4: two hashes for four elements
                dropout_keep_pair(drop_seed, e, drop_thresh, k0, k1);
                dropout_keep_pair(drop_seed, e + 2, drop_thresh, k2, k3);
                y0 = k0 ? y0 * drop_inv : 0.0f, y1 = k1 ? y1 * drop_inv : 0.0f;
                y2 = k2 ? y2 * drop_inv : 0.0f, y3 = k3 ? y3 * drop_inv : 0.0f;
            }
            o[lane + 32 * i] = make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3));
        }
        __syncwarp();  // every lane has read this buffer: lane 0 may refill it in the next iteration
    }
}

// ------------------------------------------------------------------ dbranch = gate * dx, dgate[b] = sum_t dx * branch
// grid (D / 512, B): a thread owns two adjacent columns of one sample and walks its T token rows (coalesced across the CTA)
__global__ void __launch_bounds__(kTrThreads) k_gate_residual_backward(__nv_bfloat16* __restrict__ dbr, float* __restrict__ dgate, float* __restrict__ dbias_part,
                                                                       const float* __restrict__ dx, const __nv_bfloat16* __restrict__ br, bsi_rowref gate, int T,
                                                                       int D) {
    const int c = (blockIdx.x * kTrThreads + threadIdx.x) * 2;
    if (c >= D) return;
    const int64_t b = blockIdx.y;
    const float2 g = gate.base ? *reinterpret_cast<const float2*>(rowref_ptr(gate, b, 0) + c) : make_float2(1.f, 1.f);
    float a0 = 0.f, a1 = 0.f, s0 = 0.f, s1 = 0.f;
#pragma unroll 4
    for (int t = 0; t < T; ++t) {
        const int64_t off = (b * T + t) * D + c;
        const float2 d = *reinterpret_cast<const float2*>(dx + off);
        const float2 v = bf16x2_to_float2(*reinterpret_cast<const uint32_t*>(br + off));
        a0 = fmaf(d.x, v.x, a0), a1 = fmaf(d.y, v.y, a1);
        const uint32_t o = pack_bf16(g.x * d.x, g.y * d.y);
        const float2 of = bf16x2_to_float2(o);  // the bias gradient sums what the weight-gradient GEMM sees (bf16-rounded)
        s0 += of.x, s1 += of.y;
        *reinterpret_cast<uint32_t*>(dbr + off) = o;
    }
    if (dgate) *reinterpret_cast<float2*>(dgate + b * D + c) = make_float2(a0, a1);
    if (dbias_part) *reinterpret_cast<float2*>(dbias_part + b * D + c) = make_float2(s0, s1);
}

// Row-pipelined version (round 2): the column-owner kernel above has B * D / 512 = 256 CTAs at batch 128 and a dependent load per token
// in every thread -- 4.0 TB/s.  Here a CTA of 4 warps owns `rows_per_cta` consecutive rows of one sample, each warp double-buffers its
// rows (dx 4 B + branch 2 B per element) in shared memory through the bulk-copy engine (see k_gate_residual_layernorm_pipe) and keeps
// the two column sums of its rows in registers; the CTA writes ONE partial row of each sum, [M / rows_per_cta][D], which the host adds
// per sample (dgate) or over everything (bias gradient) -- fixed order, deterministic.
template <int NV>
__global__ void __launch_bounds__(kLnbWarps * 32, 4)
    k_gate_residual_backward_pipe(__nv_bfloat16* __restrict__ dbr, float* __restrict__ dgate_part, float* __restrict__ dbias_part, const float* __restrict__ dx,
                                  const __nv_bfloat16* __restrict__ br, bsi_rowref gate, int rows_per_sample, int rows_per_cta, int64_t M) {
    constexpr int dim = 128 * NV;
    constexpr uint32_t kXBytes = dim * 4, kBrBytes = dim * 2, kRowBytes = kXBytes + kBrBytes;
    extern __shared__ __align__(128) uint8_t pipe_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* my = pipe_smem + (size_t)warp * 2 * kRowBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(pipe_smem + (size_t)kLnbWarps * 2 * kRowBytes) + warp * 2;
    const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r_end = row0 + rows_per_cta < M ? row0 + rows_per_cta : M;
    if (lane == 0) {
        ptx::mbar_init(bars, 1);
        ptx::mbar_init(bars + 1, 1);
        ptx::fence_mbar_init();
    }
    __syncwarp();
    auto issue = [&](int64_t row, int buf) {
        ptx::mbar_arrive_expect_tx(bars + buf, kRowBytes);
        ptx::bulk_load_1d(my + buf * kRowBytes, dx + row * dim, kXBytes, bars + buf);
        ptx::bulk_load_1d(my + buf * kRowBytes + kXBytes, br + row * dim, kBrBytes, bars + buf);
    };
    int64_t row = row0 + warp;
    if (row < r_end && lane == 0) issue(row, 0);
    float4 g[NV], pg[NV], pb[NV];
    {
        const float4* gr = gate.base ? reinterpret_cast<const float4*>(rowref_ptr(gate, row0 / rows_per_sample, 0)) : nullptr;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            g[i] = gr ? __ldg(gr + lane + 32 * i) : make_float4(1.f, 1.f, 1.f, 1.f);
            pg[i] = pb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    for (int k = 0; row < r_end; row += kLnbWarps, ++k) {
        const int buf = k & 1;
        if (row + kLnbWarps < r_end && lane == 0) issue(row + kLnbWarps, buf ^ 1);
        ptx::mbar_wait(bars + buf, (k >> 1) & 1);
        const float4* dr = reinterpret_cast<const float4*>(my + buf * kRowBytes);
        const uint2* brr = reinterpret_cast<const uint2*>(my + buf * kRowBytes + kXBytes);
        uint2* o = reinterpret_cast<uint2*>(dbr + row * dim);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float4 d = dr[lane + 32 * i];
            const uint2 b = brr[lane + 32 * i];
            const float2 b0 = bf16x2_to_float2(b.x), b1 = bf16x2_to_float2(b.y);
            pg[i].x = fmaf(d.x, b0.x, pg[i].x), pg[i].y = fmaf(d.y, b0.y, pg[i].y), pg[i].z = fmaf(d.z, b1.x, pg[i].z), pg[i].w = fmaf(d.w, b1.y, pg[i].w);
            const uint2 ov = make_uint2(pack_bf16(g[i].x * d.x, g[i].y * d.y), pack_bf16(g[i].z * d.z, g[i].w * d.w));
            const float2 o0 = bf16x2_to_float2(ov.x), o1 = bf16x2_to_float2(ov.y);  // the bias gradient sums what the weight-gradient GEMM sees (bf16-rounded)
            pb[i].x += o0.x, pb[i].y += o0.y, pb[i].z += o1.x, pb[i].w += o1.y;
            o[lane + 32 * i] = ov;
        }
        __syncwarp();
    }
    __syncthreads();
    float4* sm = reinterpret_cast<float4*>(pipe_smem);  // [warp][2][dim] over the idle row buffers
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        sm[(warp * 2 + 0) * (dim / 4) + lane + 32 * i] = pg[i];
        sm[(warp * 2 + 1) * (dim / 4) + lane + 32 * i] = pb[i];
    }
    __syncthreads();
    for (int q = threadIdx.x; q < 2 * (dim / 4); q += kLnbWarps * 32) {
        const int which = q / (dim / 4), col = q - which * (dim / 4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int w = 0; w < kLnbWarps; ++w) {
            const float4 t = sm[(w * 2 + which) * (dim / 4) + col];
            acc.x += t.x, acc.y += t.y, acc.z += t.z, acc.w += t.w;
        }
        float* dst = which == 0 ? dgate_part : dbias_part;
        reinterpret_cast<float4*>(dst + (int64_t)blockIdx.x * dim)[col] = acc;
    }
}

// ------------------------------------------------------------------ GELU (tanh) forward / backward on bf16
__device__ __forceinline__ float gelu_fwd(float x) {
    const float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
    return 0.5f * x * (1.0f + tanhf(u));
}
__device__ __forceinline__ float gelu_grad(float x) {
    const float x2 = x * x;
    const float t = tanhf(0.7978845608028654f * fmaf(0.044715f * x2, x, x));
    const float du = 0.7978845608028654f * fmaf(3.0f * 0.044715f, x2, 1.0f);
    return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
}
template <bool BWD>
// (out may alias up: the backward runs in place on the upstream gradient)
__global__ void __launch_bounds__(kTrThreads) k_gelu(__nv_bfloat16* out, const __nv_bfloat16* up, const __nv_bfloat16* __restrict__ pre, int64_t n8) {
    for (int64_t i = (int64_t)blockIdx.x * kTrThreads + threadIdx.x; i < n8; i += (int64_t)gridDim.x * kTrThreads) {
        const uint4 p = reinterpret_cast<const uint4*>(pre)[i];
        uint4 u = make_uint4(0, 0, 0, 0);
        if (BWD) u = reinterpret_cast<const uint4*>(up)[i];
        const uint32_t pw[4] = {p.x, p.y, p.z, p.w}, uw[4] = {u.x, u.y, u.z, u.w};
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 x = bf16x2_to_float2(pw[j]);
            if (BWD) {
                const float2 g = bf16x2_to_float2(uw[j]);
                ow[j] = pack_bf16(g.x * gelu_grad(x.x), g.y * gelu_grad(x.y));
            } else {
                ow[j] = pack_bf16(gelu_fwd(x.x), gelu_fwd(x.y));
            }
        }
        reinterpret_cast<uint4*>(out)[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

// ------------------------------------------------------------------ LayerNorm (+ modulation / affine) backward
// One warp per row (row in registers), a CTA of 8 warps owns `rows_per_cta` consecutive rows (all of one sample) and writes one
// partial row of dscale / dshift sums; the host adds the partials of a sample (fixed order -> deterministic).
template <int NV>
__global__ void __launch_bounds__(kTrThreads, 2)  // the per-warp partial sums live in shared memory: the row and its gradient fit 128 registers -> 16 warps per SM
    k_layernorm_mod_backward(float* __restrict__ dx_io, float* __restrict__ dscale_part, float* __restrict__ dshift_part, const __nv_bfloat16* __restrict__ da,
                             const float* __restrict__ x, bsi_rowref scale, const float* __restrict__ gamma, int rows_per_sample, int rows_per_cta,
                             int64_t M, float eps, uint32_t drop_thresh, uint32_t drop_seed, float drop_inv) {
    constexpr int dim = 128 * NV;
    extern __shared__ float ln_smem[];  // [8 warps][2][dim]: sums of da * xhat and of da over the warp's rows (each lane owns its columns)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
    float4* sm = reinterpret_cast<float4*>(ln_smem);
    float4* my_s = sm + (warp * 2 + 0) * (dim / 4) + lane;
    float4* my_h = sm + (warp * 2 + 1) * (dim / 4) + lane;
#pragma unroll
    for (int i = 0; i < NV; ++i) my_s[32 * i] = my_h[32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* s_mul = sm + (kTrThreads / 32) * 2 * (dim / 4);  // the row multiplier (1 + scale of this CTA's sample, or gamma), staged once
    {
        const float4* p_mul = gamma ? reinterpret_cast<const float4*>(gamma) : reinterpret_cast<const float4*>(rowref_ptr(scale, row0 / rows_per_sample, 0));
        const float one = gamma ? 0.0f : 1.0f;
        for (int q = threadIdx.x; q < dim / 4; q += kTrThreads) {
            const float4 m = p_mul[q];
            s_mul[q] = make_float4(m.x + one, m.y + one, m.z + one, m.w + one);
        }
    }
    __syncthreads();
    for (int r = warp; r < rows_per_cta; r += kTrThreads / 32) {
        const int64_t row = row0 + r;
        if (row >= M) break;
        const float4* xr = reinterpret_cast<const float4*>(x + row * dim);
        const uint2* ar = reinterpret_cast<const uint2*>(da + row * dim);
        float4 v[NV];
        uint2 apk[NV];          // the incoming gradient stays packed; the dropout decisions are kept as one bit each
        uint32_t keep = ~0u;
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[i] = xr[lane + 32 * i];
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) apk[i] = ar[lane + 32 * i];
        const float mean = warp_sum(s) * (1.0f / dim);
        float ss = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[i].x -= mean, v[i].y -= mean, v[i].z -= mean, v[i].w -= mean;
            ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
        const float rstd = rsqrtf(warp_sum(ss) * (1.0f / dim) + eps);
        if (drop_thresh) {  // the forward dropped / rescaled these outputs: the same mask applies to their gradient
            keep = 0;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const uint32_t e = (uint32_t)row * dim + (lane + 32 * i) * 4;
#pragma unroll
                for (int j = 0; j < 4; j += 2) {  // e is a multiple of 4: one hash per two elements
                    bool k0, k1;
                    dropout_keep_pair(drop_seed, e + j, drop_thresh, k0, k1);
                    keep |= ((k0 ? 1u : 0u) | (k1 ? 2u : 0u)) << (4 * i + j);
                }
            }
        }
        auto grad_in = [&](int i, float& g0, float& g1, float& g2, float& g3) {
            const float2 a0 = bf16x2_to_float2(apk[i].x), a1 = bf16x2_to_float2(apk[i].y);
            g0 = a0.x, g1 = a0.y, g2 = a1.x, g3 = a1.y;
            if (drop_thresh) {
                g0 = (keep >> (4 * i)) & 1 ? g0 * drop_inv : 0.0f, g1 = (keep >> (4 * i + 1)) & 1 ? g1 * drop_inv : 0.0f;
                g2 = (keep >> (4 * i + 2)) & 1 ? g2 * drop_inv : 0.0f, g3 = (keep >> (4 * i + 3)) & 1 ? g3 * drop_inv : 0.0f;
            }
        };
        float sg = 0.0f, sgx = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[i].x *= rstd, v[i].y *= rstd, v[i].z *= rstd, v[i].w *= rstd;  // xhat
            float a0, a1, a2, a3;
            grad_in(i, a0, a1, a2, a3);
            const float4 m = s_mul[lane + 32 * i];
            float4 ps = my_s[32 * i], ph = my_h[32 * i];
            ps.x = fmaf(a0, v[i].x, ps.x), ps.y = fmaf(a1, v[i].y, ps.y), ps.z = fmaf(a2, v[i].z, ps.z), ps.w = fmaf(a3, v[i].w, ps.w);
            ph.x += a0, ph.y += a1, ph.z += a2, ph.w += a3;
            my_s[32 * i] = ps, my_h[32 * i] = ph;
            const float4 g = make_float4(a0 * m.x, a1 * m.y, a2 * m.z, a3 * m.w);
            sg += (g.x + g.y) + (g.z + g.w);
            sgx += (g.x * v[i].x + g.y * v[i].y) + (g.z * v[i].z + g.w * v[i].w);
        }
        const float mg = warp_sum(sg) * (1.0f / dim), mgx = warp_sum(sgx) * (1.0f / dim);
        float4* dr = reinterpret_cast<float4*>(dx_io + row * dim);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float a0, a1, a2, a3;
            grad_in(i, a0, a1, a2, a3);
            const float4 m = s_mul[lane + 32 * i];
            const float4 g = make_float4(a0 * m.x, a1 * m.y, a2 * m.z, a3 * m.w);
            float4 d = dr[lane + 32 * i];
            d.x += rstd * (g.x - mg - v[i].x * mgx), d.y += rstd * (g.y - mg - v[i].y * mgx);
            d.z += rstd * (g.z - mg - v[i].z * mgx), d.w += rstd * (g.w - mg - v[i].w * mgx);
            dr[lane + 32 * i] = d;
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < 2 * (dim / 4); q += kTrThreads) {
        const int which = q / (dim / 4), col = q - which * (dim / 4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int w = 0; w < kTrThreads / 32; ++w) {
            const float4 t = sm[(w * 2 + which) * (dim / 4) + col];
            acc.x += t.x, acc.y += t.y, acc.z += t.z, acc.w += t.w;
        }
        float* dst = which == 0 ? dscale_part : dshift_part;
        reinterpret_cast<float4*>(dst + (int64_t)blockIdx.x * dim)[col] = acc;
    }
}

// Bulk-copy pipelined version (round 2; see k_gate_residual_layernorm_pipe): a CTA of 4 warps owns `rows_per_cta` consecutive rows, every warp
// double-buffers its rows (x 4 B + da 2 B + dx 4 B per element) in shared memory and keeps the partial sums in registers; two CTAs per SM.
template <int NV>
__global__ void __launch_bounds__(kLnbWarps * 32, 2)
    k_layernorm_mod_backward_pipe(float* __restrict__ dx_io, float* __restrict__ dscale_part, float* __restrict__ dshift_part, const __nv_bfloat16* __restrict__ da,
                                  const float* __restrict__ x, bsi_rowref scale, const float* __restrict__ gamma, int rows_per_sample, int rows_per_cta,
                                  int64_t M, float eps, uint32_t drop_thresh, uint32_t drop_seed, float drop_inv) {
    constexpr int dim = 128 * NV;
    constexpr uint32_t kXBytes = dim * 4, kDaBytes = dim * 2, kRowBytes = 2 * kXBytes + kDaBytes;  // x | dx | da
    extern __shared__ __align__(128) uint8_t pipe_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* my = pipe_smem + (size_t)warp * 2 * kRowBytes;
    float4* s_mul = reinterpret_cast<float4*>(pipe_smem + (size_t)kLnbWarps * 2 * kRowBytes);  // 1 + scale of this CTA's sample, or gamma
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_mul + dim / 4) + warp * 2;
    const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r_end = row0 + rows_per_cta < M ? row0 + rows_per_cta : M;
    if (lane == 0) {
        ptx::mbar_init(bars, 1);
        ptx::mbar_init(bars + 1, 1);
        ptx::fence_mbar_init();
    }
    __syncwarp();
    auto issue = [&](int64_t row, int buf) {
        uint8_t* dst = my + buf * kRowBytes;
        ptx::mbar_arrive_expect_tx(bars + buf, kRowBytes);
        ptx::bulk_load_1d(dst, x + row * dim, kXBytes, bars + buf);
        ptx::bulk_load_1d(dst + kXBytes, dx_io + row * dim, kXBytes, bars + buf);
        ptx::bulk_load_1d(dst + 2 * kXBytes, da + row * dim, kDaBytes, bars + buf);
    };
    int64_t row = row0 + warp;
    if (row < r_end && lane == 0) issue(row, 0);
    {
        const float4* p_mul = gamma ? reinterpret_cast<const float4*>(gamma) : reinterpret_cast<const float4*>(rowref_ptr(scale, row0 / rows_per_sample, 0));
        const float one = gamma ? 0.0f : 1.0f;
        for (int q = threadIdx.x; q < dim / 4; q += kLnbWarps * 32) {
            const float4 m = p_mul[q];
            s_mul[q] = make_float4(m.x + one, m.y + one, m.z + one, m.w + one);
        }
    }
    __syncthreads();
    float4 ps[NV], ph[NV];  // per-lane partial sums of da * xhat and da over this warp's rows
#pragma unroll
    for (int i = 0; i < NV; ++i) ps[i] = ph[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; row < r_end; row += kLnbWarps, ++k) {
        const int buf = k & 1;
        if (row + kLnbWarps < r_end && lane == 0) issue(row + kLnbWarps, buf ^ 1);
        ptx::mbar_wait(bars + buf, (k >> 1) & 1);
        const float4* xr = reinterpret_cast<const float4*>(my + buf * kRowBytes);
        const float4* dr = reinterpret_cast<const float4*>(my + buf * kRowBytes + kXBytes);
        const uint2* ar = reinterpret_cast<const uint2*>(my + buf * kRowBytes + 2 * kXBytes);
        float4 v[NV];
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[i] = xr[lane + 32 * i];
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        const float mean = warp_sum(s) * (1.0f / dim);
        float ss = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[i].x -= mean, v[i].y -= mean, v[i].z -= mean, v[i].w -= mean;
            ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
        const float rstd = rsqrtf(warp_sum(ss) * (1.0f / dim) + eps);
        uint32_t keep = ~0u;
        if (drop_thresh) {  // the forward dropped / rescaled these outputs: the same mask applies to their gradient
            keep = 0;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const uint32_t e = (uint32_t)row * dim + (lane + 32 * i) * 4;
#pragma unroll
                for (int j = 0; j < 4; j += 2) {  // e is a multiple of 4: one hash per two elements
                    bool k0, k1;
                    dropout_keep_pair(drop_seed, e + j, drop_thresh, k0, k1);
                    keep |= ((k0 ? 1u : 0u) | (k1 ? 2u : 0u)) << (4 * i + j);
                }
            }
        }
        auto grad_in = [&](int i, float& g0, float& g1, float& g2, float& g3) {
            const uint2 a = ar[lane + 32 * i];
            const float2 a0 = bf16x2_to_float2(a.x), a1 = bf16x2_to_float2(a.y);
            g0 = a0.x, g1 = a0.y, g2 = a1.x, g3 = a1.y;
            if (drop_thresh) {
                g0 = (keep >> (4 * i)) & 1 ? g0 * drop_inv : 0.0f, g1 = (keep >> (4 * i + 1)) & 1 ? g1 * drop_inv : 0.0f;
                g2 = (keep >> (4 * i + 2)) & 1 ? g2 * drop_inv : 0.0f, g3 = (keep >> (4 * i + 3)) & 1 ? g3 * drop_inv : 0.0f;
            }
        };
        float sg = 0.0f, sgx = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[i].x *= rstd, v[i].y *= rstd, v[i].z *= rstd, v[i].w *= rstd;  // xhat
            float a0, a1, a2, a3;
            grad_in(i, a0, a1, a2, a3);
            const float4 m = s_mul[lane + 32 * i];
            ps[i].x = fmaf(a0, v[i].x, ps[i].x), ps[i].y = fmaf(a1, v[i].y, ps[i].y), ps[i].z = fmaf(a2, v[i].z, ps[i].z), ps[i].w = fmaf(a3, v[i].w, ps[i].w);
            ph[i].x += a0, ph[i].y += a1, ph[i].z += a2, ph[i].w += a3;
            const float4 g = make_float4(a0 * m.x, a1 * m.y, a2 * m.z, a3 * m.w);
            sg += (g.x + g.y) + (g.z + g.w);
            sgx += (g.x * v[i].x + g.y * v[i].y) + (g.z * v[i].z + g.w * v[i].w);
        }
        const float mg = warp_sum(sg) * (1.0f / dim), mgx = warp_sum(sgx) * (1.0f / dim);
        float4* dout = reinterpret_cast<float4*>(dx_io + row * dim);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float a0, a1, a2, a3;
            grad_in(i, a0, a1, a2, a3);
            const float4 m = s_mul[lane + 32 * i];
            const float4 g = make_float4(a0 * m.x, a1 * m.y, a2 * m.z, a3 * m.w);
            float4 d = dr[lane + 32 * i];
            d.x += rstd * (g.x - mg - v[i].x * mgx), d.y += rstd * (g.y - mg - v[i].y * mgx);
            d.z += rstd * (g.z - mg - v[i].z * mgx), d.w += rstd * (g.w - mg - v[i].w * mgx);
            dout[lane + 32 * i] = d;
        }
        __syncwarp();  // every lane has read this buffer: lane 0 may refill it in the next iteration
    }
    // combine the warps' partial sums through the (now idle) row buffers: [warp][2][dim]
    __syncthreads();
    float4* sm = reinterpret_cast<float4*>(pipe_smem);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        sm[(warp * 2 + 0) * (dim / 4) + lane + 32 * i] = ps[i];
        sm[(warp * 2 + 1) * (dim / 4) + lane + 32 * i] = ph[i];
    }
    __syncthreads();
    for (int q = threadIdx.x; q < 2 * (dim / 4); q += kLnbWarps * 32) {
        const int which = q / (dim / 4), col = q - which * (dim / 4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int w = 0; w < kLnbWarps; ++w) {
            const float4 t = sm[(w * 2 + which) * (dim / 4) + col];
            acc.x += t.x, acc.y += t.y, acc.z += t.z, acc.w += t.w;
        }
        float* dst = which == 0 ? dscale_part : dshift_part;
        reinterpret_cast<float4*>(dst + (int64_t)blockIdx.x * dim)[col] = acc;
    }
}

// ------------------------------------------------------------------ several partial-sum reductions in one launch
// The row kernels above leave per-CTA partial rows behind ([groups * rows][D]: `rows` partial rows per group).  One launch of this kernel
// finishes up to kMaxReduceJobs of them: dst[g][:] (+)= sum_r src[g * rows + r][:], rows added in a fixed order (8 interleaved slices, then
// slice 0..7).  A CTA owns 128 columns of one group of one job; a transformer block's backward needs one launch instead of a dozen
// torch.sum / add_ calls (each of them a kernel, some with a memset, and a launch gap on the stream).
constexpr int kMaxReduceJobs = 12;
struct ReduceJobs {
    bsi_reduce_job job[kMaxReduceJobs];
    int cta_begin[kMaxReduceJobs + 1];
    int n;
};
__global__ void __launch_bounds__(kTrThreads) k_reduce_rows(const __grid_constant__ ReduceJobs J) {
    constexpr int kSlices = 8, kColq = kTrThreads / kSlices;  // 32 float4 columns (128 floats) x 8 interleaved row slices per CTA
    __shared__ float4 part[kSlices][kColq];
    int j = 0;
    while (j + 1 < J.n && (int)blockIdx.x >= J.cta_begin[j + 1]) ++j;
    const bsi_reduce_job& jb = J.job[j];
    const int chunks = (jb.D + 127) / 128;
    const int local = (int)blockIdx.x - J.cta_begin[j];
    const int g = local / chunks, chunk = local - g * chunks;
    const int colq = threadIdx.x & (kColq - 1), slice = threadIdx.x / kColq;
    const int col = chunk * 128 + colq * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < jb.D) {
        const float* src = jb.src + ((int64_t)g * jb.rows) * jb.D + col;
#pragma unroll 8
        for (int r = slice; r < jb.rows; r += kSlices) {
            const float4 v = *reinterpret_cast<const float4*>(src + (int64_t)r * jb.D);
            acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
        }
    }
    part[slice][colq] = acc;
    __syncthreads();
    if (slice == 0 && col < jb.D) {
        float4 t = part[0][colq];
#pragma unroll
        for (int sl = 1; sl < kSlices; ++sl) {
            const float4 v = part[sl][colq];
            t.x += v.x, t.y += v.y, t.z += v.z, t.w += v.w;
        }
        float4* d = reinterpret_cast<float4*>(jb.dst + (int64_t)g * jb.dst_ld + col);
        if (jb.accumulate) {
            const float4 o = *d;
            t.x += o.x, t.y += o.y, t.z += o.z, t.w += o.w;
        }
        *d = t;
    }
}

// ------------------------------------------------------------------ column sums of a bf16 [M][N] matrix (bias gradients)
// grid (N / 512, ceil(M / rows_per_cta)): a thread owns two adjacent columns and walks rows_per_cta rows; partial[chunk][N]
__global__ void __launch_bounds__(kTrThreads) k_colsum_bf16(float* __restrict__ partial, const __nv_bfloat16* __restrict__ a, int64_t M, int N, int64_t ld,
                                                            int rows_per_cta) {
    const int c = (blockIdx.x * kTrThreads + threadIdx.x) * 2;
    if (c >= N) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta, r1 = min(r0 + rows_per_cta, M);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
    for (int64_t r = r0; r < r1; ++r) {
        const float2 v = bf16x2_to_float2(*reinterpret_cast<const uint32_t*>(a + r * ld + c));
        s0 += v.x, s1 += v.y;
    }
    *reinterpret_cast<float2*>(partial + (int64_t)blockIdx.y * N + c) = make_float2(s0, s1);
}

// ------------------------------------------------------------------ fp32 weight -> bf16 copy and bf16 transposed copy in one pass
// in [R][C] fp32 -> out [R][ld_out] bf16 (forward operand, optional) and out_t [C][ld_t] bf16 (operand of the data-gradient GEMM);
// 32 x 32 tiles through shared memory so that both global accesses are coalesced.  block (32, 8).
__global__ void __launch_bounds__(256) k_cast_transpose_bf16(__nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out_t, const float* __restrict__ in,
                                                             int R, int Cc, int ld_out, int ld_t) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        const float v = (r < R && c < Cc) ? in[(int64_t)r * Cc + c] : 0.0f;
        tile[j][threadIdx.x] = v;
        if (out && r < R && c < Cc) out[(int64_t)r * ld_out + c] = __float2bfloat16(v);
    }
    __syncthreads();
    if (out_t) {
        for (int j = threadIdx.y; j < 32; j += 8) {
            const int c = c0 + j, r = r0 + threadIdx.x;
            if (c < Cc && r < R) out_t[(int64_t)c * ld_t + r] = __float2bfloat16(tile[threadIdx.x][j]);
        }
    }
}

static unsigned tr_grid(int64_t work) {
    const int64_t need = (work + kTrThreads - 1) / kTrThreads, cap = (int64_t)sm_count() * 8;
    return (unsigned)(need < cap ? need : cap);
}

}  // namespace bsi

using namespace bsi;

extern "C" {

int bsi_gate_residual(float* x_out, const float* x, const void* branch_bf16, bsi_rowref gate, int32_t rows_per_sample, int64_t M, int32_t D, void* stream) {
    BSI_CHECK_ARG(x_out && x && branch_bf16 && M > 0 && D > 0 && D % 4 == 0 && rows_per_sample > 0, "bsi_gate_residual: bad arguments (D=%d must be a multiple of 4)", D);
    k_gate_residual<<<tr_grid(M * (D / 4)), kTrThreads, 0, (cudaStream_t)stream>>>(x_out, x, (const __nv_bfloat16*)branch_bf16, gate, rows_per_sample, M, D);
    BSI_LAUNCH_OK("k_gate_residual");
    return BSI_OK;
}

int bsi_gate_residual_layernorm_bf16(void* out_bf16, float* x_out, const float* x, const void* branch_bf16, bsi_rowref gate, bsi_rowref shift, bsi_rowref scale,
                                     const float* gamma, const float* beta, int32_t rows_per_sample, int64_t M, int32_t dim, float eps, float drop_p,
                                     uint32_t drop_seed, void* stream) {
    BSI_CHECK_ARG(out_bf16 && x_out && x && branch_bf16 && M > 0 && rows_per_sample > 0, "bsi_gate_residual_layernorm_bf16: null pointer or empty input");
    BSI_CHECK_ARG((gamma && beta) || (shift.base && scale.base), "bsi_gate_residual_layernorm_bf16: need either gamma/beta or shift/scale");
    BSI_CHECK_ARG(dim % 128 == 0 && dim >= 128 && dim <= 1024, "bsi_gate_residual_layernorm_bf16: dim=%d must be a multiple of 128 in [128,1024]", dim);
    BSI_CHECK_ARG(drop_p >= 0.0f && drop_p < 1.0f && (drop_p == 0.0f || M * dim < (int64_t)1 << 32), "bsi_gate_residual_layernorm_bf16: bad dropout arguments");
    const uint32_t drop_thresh = dropout_thresh(drop_p);
    const float drop_inv = drop_p > 0.0f ? 1.0f / (1.0f - drop_p) : 1.0f;
    // BSI_TRAIN_PIPE=0 keeps the register-resident kernels (A/B measurements); default: rows staged by the bulk-copy engine
    static const bool pipe = [] { const char* e = getenv("BSI_TRAIN_PIPE"); return !(e && e[0] == '0'); }();
    if (pipe && M >= 2 * (int64_t)sm_count()) {
        const int grid = 2 * sm_count();
        const int smem = kPipeWarps * 2 * dim * 6 + kPipeWarps * 2 * 8;
#define BSI_GRLP_CASE(NV)                                                                                                                                   \
    case NV:                                                                                                                                                \
        BSI_ENSURE_SMEM(k_gate_residual_layernorm_pipe<NV>, smem);                                                                                           \
        k_gate_residual_layernorm_pipe<NV><<<grid, kTrThreads, smem, (cudaStream_t)stream>>>((__nv_bfloat16*)out_bf16, x_out, x, (const __nv_bfloat16*)branch_bf16, \
                                                                                             gate, shift, scale, gamma, beta, rows_per_sample, M, eps,      \
                                                                                             drop_thresh, drop_seed, drop_inv);                             \
        break;
        switch (dim / 128) {
            BSI_GRLP_CASE(1) BSI_GRLP_CASE(2) BSI_GRLP_CASE(3) BSI_GRLP_CASE(4) BSI_GRLP_CASE(5) BSI_GRLP_CASE(6) BSI_GRLP_CASE(7) BSI_GRLP_CASE(8)
            default: set_error("unsupported dim %d", dim); return BSI_ERR_UNSUPPORTED;
        }
#undef BSI_GRLP_CASE
        BSI_LAUNCH_OK("k_gate_residual_layernorm_pipe");
        return BSI_OK;
    }
    const int64_t blocks = (M + (kTrThreads / 32) - 1) / (kTrThreads / 32), cap = (int64_t)sm_count() * 8;
    const int grid = (int)(blocks < cap ? blocks : cap);
#define BSI_GRL_CASE(NV)                                                                                                                          \
    case NV:                                                                                                                                      \
        k_gate_residual_layernorm<NV><<<grid, kTrThreads, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)out_bf16, x_out, x, (const __nv_bfloat16*)branch_bf16, \
                                                                                     gate, shift, scale, gamma, beta, rows_per_sample, M, eps, drop_thresh, \
                                                                                     drop_seed, drop_inv);                                       \
        break;
    switch (dim / 128) {
        BSI_GRL_CASE(1) BSI_GRL_CASE(2) BSI_GRL_CASE(3) BSI_GRL_CASE(4) BSI_GRL_CASE(5) BSI_GRL_CASE(6) BSI_GRL_CASE(7) BSI_GRL_CASE(8)
        default: set_error("unsupported dim %d", dim); return BSI_ERR_UNSUPPORTED;
    }
#undef BSI_GRL_CASE
    BSI_LAUNCH_OK("k_gate_residual_layernorm");
    return BSI_OK;
}

int bsi_gate_residual_backward(void* dbranch_bf16, float* dgate, float* dbias_part, const float* dx, const void* branch_bf16, bsi_rowref gate,
                               int32_t rows_per_sample, int32_t B, int32_t D, void* stream) {
    BSI_CHECK_ARG(dbranch_bf16 && dx && branch_bf16 && B > 0 && D > 0 && D % 2 == 0 && rows_per_sample > 0, "bsi_gate_residual_backward: bad arguments");
    dim3 grid((D / 2 + kTrThreads - 1) / kTrThreads, B);
    k_gate_residual_backward<<<grid, kTrThreads, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)dbranch_bf16, dgate, dbias_part, dx,
                                                                          (const __nv_bfloat16*)branch_bf16, gate, rows_per_sample, D);
    BSI_LAUNCH_OK("k_gate_residual_backward");
    return BSI_OK;
}

int bsi_gate_residual_backward_rows(void* dbranch_bf16, float* dgate_part, float* dbias_part, const float* dx, const void* branch_bf16, bsi_rowref gate,
                                    int32_t rows_per_sample, int32_t rows_per_cta, int64_t M, int32_t D, void* stream) {
    BSI_CHECK_ARG(dbranch_bf16 && dgate_part && dbias_part && dx && branch_bf16 && M > 0 && rows_per_sample > 0, "bsi_gate_residual_backward_rows: bad arguments");
    BSI_CHECK_ARG(D % 128 == 0 && D >= 128 && D <= 1024, "bsi_gate_residual_backward_rows: D=%d must be a multiple of 128 in [128,1024]", D);
    BSI_CHECK_ARG(rows_per_cta > 0 && rows_per_sample % rows_per_cta == 0, "bsi_gate_residual_backward_rows: rows_per_cta must divide rows_per_sample");
    const int grid = (int)((M + rows_per_cta - 1) / rows_per_cta);
    const int smem = kLnbWarps * 2 * D * 6 + kLnbWarps * 2 * 8;
#define BSI_GRB_CASE(NV)                                                                                                                          \
    case NV:                                                                                                                                      \
        BSI_ENSURE_SMEM(k_gate_residual_backward_pipe<NV>, smem);                                                                                  \
        k_gate_residual_backward_pipe<NV><<<grid, kLnbWarps * 32, smem, (cudaStream_t)stream>>>((__nv_bfloat16*)dbranch_bf16, dgate_part, dbias_part, dx, \
                                                                                                (const __nv_bfloat16*)branch_bf16, gate, rows_per_sample, \
                                                                                                rows_per_cta, M);                                 \
        break;
    switch (D / 128) {
        BSI_GRB_CASE(1) BSI_GRB_CASE(2) BSI_GRB_CASE(3) BSI_GRB_CASE(4) BSI_GRB_CASE(5) BSI_GRB_CASE(6) BSI_GRB_CASE(7) BSI_GRB_CASE(8)
        default: set_error("unsupported D %d", D); return BSI_ERR_UNSUPPORTED;
    }
#undef BSI_GRB_CASE
    BSI_LAUNCH_OK("k_gate_residual_backward_pipe");
    return BSI_OK;
}

int bsi_reduce_rows(const bsi_reduce_job* jobs, int32_t n_jobs, void* stream) {
    BSI_CHECK_ARG(jobs && n_jobs > 0 && n_jobs <= kMaxReduceJobs, "bsi_reduce_rows: between 1 and %d jobs per call (got %d)", kMaxReduceJobs, n_jobs);
    ReduceJobs J;
    J.n = n_jobs;
    int total = 0;
    for (int i = 0; i < n_jobs; ++i) {
        const bsi_reduce_job& jb = jobs[i];
        BSI_CHECK_ARG(jb.src && jb.dst && jb.groups > 0 && jb.rows > 0 && jb.D > 0 && jb.D % 4 == 0 && jb.dst_ld >= jb.D && jb.dst_ld % 4 == 0,
                      "bsi_reduce_rows: job %d: bad arguments (groups=%d rows=%d D=%d dst_ld=%d; D and dst_ld must be multiples of 4)", i, jb.groups, jb.rows,
                      jb.D, jb.dst_ld);
        BSI_CHECK_ARG((reinterpret_cast<uintptr_t>(jb.src) | reinterpret_cast<uintptr_t>(jb.dst)) % 16 == 0, "bsi_reduce_rows: job %d: pointers must be 16-byte aligned", i);
        J.job[i] = jb;
        J.cta_begin[i] = total;
        total += jb.groups * ((jb.D + 127) / 128);
    }
    for (int i = n_jobs; i <= kMaxReduceJobs; ++i) J.cta_begin[i] = total;
    k_reduce_rows<<<total, kTrThreads, 0, (cudaStream_t)stream>>>(J);
    BSI_LAUNCH_OK("k_reduce_rows");
    return BSI_OK;
}

int bsi_colsum_bf16(float* partial, const void* a_bf16, int64_t M, int32_t N, int64_t ld, int32_t rows_per_cta, void* stream) {
    BSI_CHECK_ARG(partial && a_bf16 && M > 0 && N > 0 && N % 2 == 0 && ld >= N && ld % 2 == 0 && rows_per_cta > 0, "bsi_colsum_bf16: bad arguments");
    dim3 grid((N / 2 + kTrThreads - 1) / kTrThreads, (unsigned)((M + rows_per_cta - 1) / rows_per_cta));
    k_colsum_bf16<<<grid, kTrThreads, 0, (cudaStream_t)stream>>>(partial, (const __nv_bfloat16*)a_bf16, M, N, ld, rows_per_cta);
    BSI_LAUNCH_OK("k_colsum_bf16");
    return BSI_OK;
}

int bsi_cast_transpose_bf16(void* out_bf16, void* out_t_bf16, const float* in, int32_t rows, int32_t cols, int32_t ld_out, int32_t ld_t, void* stream) {
    BSI_CHECK_ARG((out_bf16 || out_t_bf16) && in && rows > 0 && cols > 0 && (!out_bf16 || ld_out >= cols) && (!out_t_bf16 || ld_t >= rows),
                  "bsi_cast_transpose_bf16: bad arguments");
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    k_cast_transpose_bf16<<<grid, block, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)out_bf16, (__nv_bfloat16*)out_t_bf16, in, rows, cols, ld_out, ld_t);
    BSI_LAUNCH_OK("k_cast_transpose_bf16");
    return BSI_OK;
}

int bsi_gelu_bf16(void* out_bf16, const void* pre_bf16, int64_t numel, void* stream) {
    BSI_CHECK_ARG(out_bf16 && pre_bf16 && numel > 0 && numel % 8 == 0, "bsi_gelu_bf16: numel (%lld) must be a positive multiple of 8", (long long)numel);
    k_gelu<false><<<tr_grid(numel / 8), kTrThreads, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)out_bf16, nullptr, (const __nv_bfloat16*)pre_bf16, numel / 8);
    BSI_LAUNCH_OK("k_gelu");
    return BSI_OK;
}

int bsi_gelu_backward_bf16(void* dpre_bf16, const void* dout_bf16, const void* pre_bf16, int64_t numel, void* stream) {
    BSI_CHECK_ARG(dpre_bf16 && dout_bf16 && pre_bf16 && numel > 0 && numel % 8 == 0, "bsi_gelu_backward_bf16: numel (%lld) must be a positive multiple of 8",
                  (long long)numel);
    k_gelu<true><<<tr_grid(numel / 8), kTrThreads, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)dpre_bf16, (const __nv_bfloat16*)dout_bf16,
                                                                             (const __nv_bfloat16*)pre_bf16, numel / 8);
    BSI_LAUNCH_OK("k_gelu_backward");
    return BSI_OK;
}

int bsi_layernorm_mod_backward(float* dx_io, float* dscale_part, float* dshift_part, const void* da_bf16, const float* x, bsi_rowref scale, const float* gamma,
                               int32_t rows_per_sample, int32_t rows_per_cta, int64_t M, int32_t dim, float eps, float drop_p, uint32_t drop_seed,
                               void* stream) {
    BSI_CHECK_ARG(drop_p >= 0.0f && drop_p < 1.0f && (drop_p == 0.0f || M * dim < (int64_t)1 << 32), "bsi_layernorm_mod_backward: bad dropout arguments");
    const uint32_t drop_thresh = dropout_thresh(drop_p);
    const float drop_inv = drop_p > 0.0f ? 1.0f / (1.0f - drop_p) : 1.0f;
    BSI_CHECK_ARG(dx_io && dscale_part && dshift_part && da_bf16 && x && M > 0, "bsi_layernorm_mod_backward: null pointer or empty input");
    BSI_CHECK_ARG(gamma || (scale.base && rows_per_sample > 0), "bsi_layernorm_mod_backward: need either gamma or scale");
    BSI_CHECK_ARG(dim % 128 == 0 && dim >= 128 && dim <= 1024, "bsi_layernorm_mod_backward: dim=%d must be a multiple of 128 in [128,1024]", dim);
    BSI_CHECK_ARG(rows_per_cta > 0 && (gamma || rows_per_sample % rows_per_cta == 0), "bsi_layernorm_mod_backward: rows_per_cta must divide rows_per_sample");
    const int grid = (int)((M + rows_per_cta - 1) / rows_per_cta);
    static const bool pipe = [] { const char* e = getenv("BSI_TRAIN_PIPE"); return !(e && e[0] == '0'); }();
    if (pipe) {
        const int smem_p = kLnbWarps * 2 * dim * 10 + dim * 4 + kLnbWarps * 2 * 8;
#define BSI_LNBP_CASE(NV)                                                                                                                              \
    case NV:                                                                                                                                           \
        BSI_ENSURE_SMEM(k_layernorm_mod_backward_pipe<NV>, smem_p);                                                                                     \
        k_layernorm_mod_backward_pipe<NV><<<grid, kLnbWarps * 32, smem_p, (cudaStream_t)stream>>>(dx_io, dscale_part, dshift_part, (const __nv_bfloat16*)da_bf16, x, \
                                                                                                  scale, gamma, rows_per_sample, rows_per_cta, M, eps, \
                                                                                                  drop_thresh, drop_seed, drop_inv);                   \
        break;
        switch (dim / 128) {
            BSI_LNBP_CASE(1) BSI_LNBP_CASE(2) BSI_LNBP_CASE(3) BSI_LNBP_CASE(4) BSI_LNBP_CASE(5) BSI_LNBP_CASE(6) BSI_LNBP_CASE(7) BSI_LNBP_CASE(8)
            default: set_error("unsupported dim %d", dim); return BSI_ERR_UNSUPPORTED;
        }
#undef BSI_LNBP_CASE
        BSI_LAUNCH_OK("k_layernorm_mod_backward_pipe");
        return BSI_OK;
    }
    const int smem = ((kTrThreads / 32) * 2 + 1) * dim * (int)sizeof(float);
#define BSI_LNB_CASE(NV)                                                                                                                      \
    case NV:                                                                                                                                  \
        BSI_ENSURE_SMEM(k_layernorm_mod_backward<NV>, smem);                                                                                   \
        k_layernorm_mod_backward<NV><<<grid, kTrThreads, smem, (cudaStream_t)stream>>>(dx_io, dscale_part, dshift_part, (const __nv_bfloat16*)da_bf16, x, \
                                                                                       scale, gamma, rows_per_sample, rows_per_cta, M, eps, drop_thresh, drop_seed, drop_inv); \
        break;
    switch (dim / 128) {
        BSI_LNB_CASE(1) BSI_LNB_CASE(2) BSI_LNB_CASE(3) BSI_LNB_CASE(4) BSI_LNB_CASE(5) BSI_LNB_CASE(6) BSI_LNB_CASE(7) BSI_LNB_CASE(8)
        default: set_error("unsupported dim %d", dim); return BSI_ERR_UNSUPPORTED;
    }
#undef BSI_LNB_CASE
    BSI_LAUNCH_OK("k_layernorm_mod_backward");
    return BSI_OK;
}

}  // extern "C"
