// LayerNorm (no affine) fused with the adaLN modulation, emitting the bf16 A-operand of the next
// GEMM (bsi/models/dit.py:55,66,96,101), plus the affine LayerNorm of the patch decoder
// (bsi/models/dit.py:164).  HBM-bound: reads 4 B/elem (fp32 residual stream), writes 2 B/elem.
// One warp per token row; the row lives in registers (two-pass mean / variance).
#include <cstdlib>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace bsi {

constexpr int kLnThreads = 256;

template <int NV, bool F32OUT = false>  // NV float4 per lane: dim = 128 * NV; F32OUT: fp32 output (the fp32-accurate mode, exact_kernels.cu)
__global__ void __launch_bounds__(kLnThreads, 2)
    k_layernorm_mod(void* __restrict__ out_raw, const float* __restrict__ x, bsi_rowref shift, bsi_rowref scale,
                    const int32_t* __restrict__ step_ptr, const float* __restrict__ gamma, const float* __restrict__ beta,
                    int rows_per_sample, int64_t M, float eps, uint32_t drop_thresh, uint32_t drop_seed, float drop_inv) {
    constexpr int dim = 128 * NV;
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * (kLnThreads / 32);
    pdl_prologue_done();
    const int step = step_ptr ? *step_ptr : 0;
    for (int64_t row = (int64_t)blockIdx.x * (kLnThreads / 32) + (threadIdx.x >> 5); row < M; row += warps_total) {
        const float4* xr = reinterpret_cast<const float4*>(x + row * dim);
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
            float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            ss += (a * a + b * b) + (c * c + d * d);
        }
        const float rstd = rsqrtf(warp_sum(ss) * (1.0f / dim) + eps);
        const float4 *p_mul, *p_add;
        if (gamma) {
            p_mul = reinterpret_cast<const float4*>(gamma);
            p_add = reinterpret_cast<const float4*>(beta);
        } else {
            const int64_t sample = row / rows_per_sample;
            p_mul = reinterpret_cast<const float4*>(rowref_ptr(scale, sample, step));
            p_add = reinterpret_cast<const float4*>(rowref_ptr(shift, sample, step));
        }
        const float one = gamma ? 0.0f : 1.0f;  // modulate uses (1 + scale)
        uint2* o = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out_raw) + row * dim);
        float4* o32 = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_raw) + row * dim);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float4 m = p_mul[lane + 32 * i], a = p_add[lane + 32 * i];
            float y0 = fmaf((v[i].x - mean) * rstd, m.x + one, a.x);
            float y1 = fmaf((v[i].y - mean) * rstd, m.y + one, a.y);
            float y2 = fmaf((v[i].z - mean) * rstd, m.z + one, a.z);
            float y3 = fmaf((v[i].w - mean) * rstd, m.w + one, a.w);
            if (drop_thresh) {  // nn.Dropout on the modulated activations (training, dit.py:101): element index = row * dim + column
                const uint32_t e = (uint32_t)row * dim + (lane + 32 * i) * 4;
                bool k0, k1, k2, k3;  // e is a multiple of 4: two hashes for four elements
                dropout_keep_pair(drop_seed, e, drop_thresh, k0, k1);
                dropout_keep_pair(drop_seed, e + 2, drop_thresh, k2, k3);
                y0 = k0 ? y0 * drop_inv : 0.0f, y1 = k1 ? y1 * drop_inv : 0.0f;
                y2 = k2 ? y2 * drop_inv : 0.0f, y3 = k3 ? y3 * drop_inv : 0.0f;
            }
            if constexpr (F32OUT) o32[lane + 32 * i] = make_float4(y0, y1, y2, y3);
            else o[lane + 32 * i] = make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3));
        }
    }
}

// Bulk-copy pipelined version (round 2).  The register kernel above is latency-bound whenever the modulation vectors differ per sample
// (ELBO / training: t is per data point): each row costs a DRAM round trip for x and then an L2 round trip for shift / scale, with 16
// warps per SM -- 2.9 TB/s at [32768, 1024] against 5.7 TB/s when one conditioning row serves every sample (the sampler).  Here each
// warp owns kLnBufs row buffers in shared memory that lane 0 keeps filled two rows ahead with `cp.async.bulk` (completion on per-warp
// mbarriers), a CTA walks a contiguous range of rows (one or two samples: their vectors stay in L1) and the vector loads are issued
// before the warp waits for its row.  Same arithmetic in the same order: bit-identical output.
constexpr int kLnBufs = 3;
template <int NV, bool F32OUT>
__global__ void __launch_bounds__(kLnThreads, 2)
    k_layernorm_mod_pipe(void* __restrict__ out_raw, const float* __restrict__ x, bsi_rowref shift, bsi_rowref scale, const int32_t* __restrict__ step_ptr,
                         const float* __restrict__ gamma, const float* __restrict__ beta, int rows_per_sample, int64_t M, float eps, uint32_t drop_thresh,
                         uint32_t drop_seed, float drop_inv) {
    constexpr int dim = 128 * NV;
    constexpr uint32_t kRowBytes = dim * 4;
    constexpr int kWarps = kLnThreads / 32;
    extern __shared__ __align__(128) uint8_t ln_pipe_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* my = ln_pipe_smem + (size_t)warp * kLnBufs * kRowBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ln_pipe_smem + (size_t)kWarps * kLnBufs * kRowBytes) + warp * kLnBufs;
    if (lane == 0) {
        for (int i = 0; i < kLnBufs; ++i) ptx::mbar_init(bars + i, 1);
        ptx::fence_mbar_init();
    }
    __syncwarp();
    pdl_prologue_done();  // x is written by the kernel before this one: nothing of it may be read above this line
    const int step = step_ptr ? *step_ptr : 0;
    const int64_t per = M / gridDim.x, rem = M % gridDim.x;
    const int64_t r_begin = blockIdx.x * per + (blockIdx.x < rem ? blockIdx.x : rem), r_end = r_begin + per + (blockIdx.x < rem ? 1 : 0);
    auto issue = [&](int64_t row, int buf) {
        ptx::mbar_arrive_expect_tx(bars + buf, kRowBytes);
        ptx::bulk_load_1d(my + buf * kRowBytes, x + row * dim, kRowBytes, bars + buf);
    };
    int64_t row = r_begin + warp;
    if (lane == 0) {
        for (int i = 0; i < kLnBufs - 1; ++i)
            if (row + (int64_t)i * kWarps < r_end) issue(row + (int64_t)i * kWarps, i);
    }
    int buf = 0;
    uint32_t phase = 0;
    for (; row < r_end; row += kWarps) {
        const int nbuf = buf == 0 ? kLnBufs - 1 : buf - 1;  // the buffer consumed in the previous iteration
        if (row + (int64_t)(kLnBufs - 1) * kWarps < r_end && lane == 0) issue(row + (int64_t)(kLnBufs - 1) * kWarps, nbuf);
        const float4 *p_mul, *p_add;
        if (gamma) {
            p_mul = reinterpret_cast<const float4*>(gamma);
            p_add = reinterpret_cast<const float4*>(beta);
        } else {
            const int64_t sample = row / rows_per_sample;
            p_mul = reinterpret_cast<const float4*>(rowref_ptr(scale, sample, step));
            p_add = reinterpret_cast<const float4*>(rowref_ptr(shift, sample, step));
        }
        float4 m[NV], a[NV];  // issued before the wait: their latency hides behind the row's
#pragma unroll
        for (int i = 0; i < NV; ++i) m[i] = __ldg(p_mul + lane + 32 * i), a[i] = __ldg(p_add + lane + 32 * i);
        ptx::mbar_wait(bars + buf, phase);
        const float4* xr = reinterpret_cast<const float4*>(my + buf * kRowBytes);
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
            float a_ = v[i].x - mean, b_ = v[i].y - mean, c_ = v[i].z - mean, d_ = v[i].w - mean;
            ss += (a_ * a_ + b_ * b_) + (c_ * c_ + d_ * d_);
        }
        const float rstd = rsqrtf(warp_sum(ss) * (1.0f / dim) + eps);
        const float one = gamma ? 0.0f : 1.0f;  // modulate uses (1 + scale)
        uint2* o = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out_raw) + row * dim);
        float4* o32 = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_raw) + row * dim);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float y0 = fmaf((v[i].x - mean) * rstd, m[i].x + one, a[i].x);
            float y1 = fmaf((v[i].y - mean) * rstd, m[i].y + one, a[i].y);
            float y2 = fmaf((v[i].z - mean) * rstd, m[i].z + one, a[i].z);
            float y3 = fmaf((v[i].w - mean) * rstd, m[i].w + one, a[i].w);
            if (drop_thresh) {
                const uint32_t e = (uint32_t)row * dim + (lane + 32 * i) * 4;
                bool k0, k1, k2, k3;  // e is a multiple of 4: two hashes for four elements
                dropout_keep_pair(drop_seed, e, drop_thresh, k0, k1);
                dropout_keep_pair(drop_seed, e + 2, drop_thresh, k2, k3);
                y0 = k0 ? y0 * drop_inv : 0.0f, y1 = k1 ? y1 * drop_inv : 0.0f;
                y2 = k2 ? y2 * drop_inv : 0.0f, y3 = k3 ? y3 * drop_inv : 0.0f;
            }
            if constexpr (F32OUT) o32[lane + 32 * i] = make_float4(y0, y1, y2, y3);
            else o[lane + 32 * i] = make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3));
        }
        __syncwarp();  // every lane has read this buffer: lane 0 may refill it in the next iteration
        if (++buf == kLnBufs) buf = 0, phase ^= 1;
    }
}

}  // namespace bsi

using namespace bsi;

static int layernorm_mod_launch(void* out_bf16, const float* x, bsi_rowref shift, bsi_rowref scale, const int32_t* step_ptr, const float* gamma,
                                const float* beta, int32_t rows_per_sample, int64_t M, int32_t dim, float eps, float drop_p, uint32_t drop_seed,
                                void* stream, bool f32_out = false) {
    const uint32_t drop_thresh = dropout_thresh(drop_p);
    const float drop_inv = drop_p > 0.0f ? 1.0f / (1.0f - drop_p) : 1.0f;
    BSI_CHECK_ARG(out_bf16 && x && M > 0, "bsi_layernorm_mod_bf16: null pointer or empty input");
    BSI_CHECK_ARG((gamma && beta) || (shift.base && scale.base && rows_per_sample > 0),
                  "bsi_layernorm_mod_bf16: need either gamma/beta or shift/scale");
    BSI_CHECK_ARG(dim % 128 == 0 && dim >= 128 && dim <= 2048, "bsi_layernorm_mod_bf16: dim=%d must be a multiple of 128 in [128,2048]", dim);
    int64_t blocks = (M + (kLnThreads / 32) - 1) / (kLnThreads / 32);
    int64_t cap = (int64_t)sm_count() * 8;
    int grid = (int)(blocks < cap ? blocks : cap);
    void* o = out_bf16;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(kLnThreads), cfg.dynamicSmemBytes = 0, cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    fill_pdl_attr(&attr[0]);
    cfg.attrs = attr, cfg.numAttrs = use_pdl() ? 1 : 0;
    // The pipelined kernel serves per-sample conditioning (ELBO, training: +4.7 % on elbo(x[256], 1, 10)); with one conditioning row for the
    // whole batch (the sampler) the vectors stay in L1 and the register kernel already runs at 5.7 TB/s -- measured equal within 0.3 %,
    // so that path is left as it was.  BSI_LN_PIPE=0 / 2: never / always pipelined (A/B measurements).
    static const int pipe = [] { const char* e = getenv("BSI_LN_PIPE"); return e ? atoi(e) : 1; }();
    const bool per_sample = !gamma && (shift.sample_stride != 0 || scale.sample_stride != 0);
    if ((pipe == 2 || (pipe == 1 && per_sample)) && dim <= 1024 && M >= 2 * (int64_t)sm_count()) {
        cfg.gridDim = dim3(2 * sm_count());
        cfg.dynamicSmemBytes = (kLnThreads / 32) * kLnBufs * (dim * 4 + 8);
#define BSI_LNP_CASE(NV)                                                                                                                                   \
    case NV:                                                                                                                                               \
        if (f32_out) {                                                                                                                                     \
            BSI_ENSURE_SMEM((k_layernorm_mod_pipe<NV, true>), (int)cfg.dynamicSmemBytes);                                                                   \
            BSI_CUDA_OK(cudaLaunchKernelEx(&cfg, k_layernorm_mod_pipe<NV, true>, o, x, shift, scale, step_ptr, gamma, beta, rows_per_sample, M, eps, drop_thresh, drop_seed, drop_inv)); \
        } else {                                                                                                                                           \
            BSI_ENSURE_SMEM((k_layernorm_mod_pipe<NV, false>), (int)cfg.dynamicSmemBytes);                                                                  \
            BSI_CUDA_OK(cudaLaunchKernelEx(&cfg, k_layernorm_mod_pipe<NV, false>, o, x, shift, scale, step_ptr, gamma, beta, rows_per_sample, M, eps, drop_thresh, drop_seed, drop_inv)); \
        }                                                                                                                                                  \
        break;
        switch (dim / 128) {
            BSI_LNP_CASE(1) BSI_LNP_CASE(2) BSI_LNP_CASE(3) BSI_LNP_CASE(4) BSI_LNP_CASE(5) BSI_LNP_CASE(6) BSI_LNP_CASE(7) BSI_LNP_CASE(8)
            default: set_error("unsupported dim %d", dim); return BSI_ERR_UNSUPPORTED;
        }
#undef BSI_LNP_CASE
        BSI_LAUNCH_OK("k_layernorm_mod_pipe");
        return BSI_OK;
    }
#define BSI_LN_CASE(NV)                                                                                                         \
    case NV:                                                                                                                    \
        if (f32_out)                                                                                                            \
            BSI_CUDA_OK(cudaLaunchKernelEx(&cfg, k_layernorm_mod<NV, true>, o, x, shift, scale, step_ptr, gamma, beta, rows_per_sample, M, eps, drop_thresh, drop_seed, drop_inv)); \
        else                                                                                                                    \
            BSI_CUDA_OK(cudaLaunchKernelEx(&cfg, k_layernorm_mod<NV, false>, o, x, shift, scale, step_ptr, gamma, beta, rows_per_sample, M, eps, drop_thresh, drop_seed, drop_inv)); \
        break;
    switch (dim / 128) {
        BSI_LN_CASE(1) BSI_LN_CASE(2) BSI_LN_CASE(3) BSI_LN_CASE(4) BSI_LN_CASE(5) BSI_LN_CASE(6) BSI_LN_CASE(7) BSI_LN_CASE(8)
        BSI_LN_CASE(9) BSI_LN_CASE(10) BSI_LN_CASE(11) BSI_LN_CASE(12) BSI_LN_CASE(13) BSI_LN_CASE(14) BSI_LN_CASE(15) BSI_LN_CASE(16)
        default: set_error("unsupported dim %d", dim); return BSI_ERR_UNSUPPORTED;
    }
#undef BSI_LN_CASE
    BSI_LAUNCH_OK("k_layernorm_mod");
    return BSI_OK;
}

extern "C" int bsi_layernorm_mod_bf16(void* out_bf16, const float* x, bsi_rowref shift, bsi_rowref scale, const int32_t* step_ptr,
                                      const float* gamma, const float* beta, int32_t rows_per_sample, int64_t M, int32_t dim,
                                      float eps, void* stream) {
    return layernorm_mod_launch(out_bf16, x, shift, scale, step_ptr, gamma, beta, rows_per_sample, M, dim, eps, 0.0f, 0u, stream);
}

extern "C" int bsi_layernorm_mod_dropout_bf16(void* out_bf16, const float* x, bsi_rowref shift, bsi_rowref scale, int32_t rows_per_sample, int64_t M,
                                              int32_t dim, float eps, float drop_p, uint32_t drop_seed, void* stream) {
    BSI_CHECK_ARG(drop_p >= 0.0f && drop_p < 1.0f && M * dim < (int64_t)1 << 32, "bsi_layernorm_mod_dropout_bf16: p in [0,1) and M*dim < 2^32 required");
    return layernorm_mod_launch(out_bf16, x, shift, scale, nullptr, nullptr, nullptr, rows_per_sample, M, dim, eps, drop_p, drop_seed, stream);
}

// fp32 output: the operand of the fp32-accurate mode before it is split into bf16 terms (exact_kernels.cu)
extern "C" int bsi_layernorm_mod_f32(float* out_f32, const float* x, bsi_rowref shift, bsi_rowref scale, const int32_t* step_ptr, const float* gamma,
                                     const float* beta, int32_t rows_per_sample, int64_t M, int32_t dim, float eps, void* stream) {
    return layernorm_mod_launch(out_f32, x, shift, scale, step_ptr, gamma, beta, rows_per_sample, M, dim, eps, 0.0f, 0u, stream, true);
}
