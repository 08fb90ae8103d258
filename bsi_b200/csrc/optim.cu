// Optimizer side of the training step: gradient clipping by global norm, AdamW, EMA of the
// weights and the bf16 copy the GEMM engines read, as two launches over one flat fp32 arena instead of the reference's
// clip_grad_norm_ + torch.optim.AdamW(fused) + torch._foreach_lerp_ passes
// (config/train.yaml:40, config/task/optimizer/adamw.yaml, bsi/tasks/ema_pytorch.py:343-434, bsi/tasks/bsi.py:196-198).
// HBM-bound.  Algorithmic bytes per parameter: sum of squares 4 (read g); step 20 read (p, g, m, v, ema) + 16 write
// (p, m, v, ema) + 2 (bf16 copy) + 4 (zeroed gradient) = 42.
// The arithmetic follows torch's single-tensor AdamW op by op, each result rounded to fp32 like the eager ops do:
//   p.mul_(1 - lr*wd); m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, value=1-b2);
//   denom = (v.sqrt() / sqrt(1-b2^t)).add_(eps); p.addcdiv_(m, denom, value=-lr/(1-b1^t))
// with g = grad * min(max_norm / (||grad|| + 1e-6), 1) as torch.nn.utils.clip_grad_norm_ scales it.  Under data parallelism
// the gradient arena is summed over ranks by ONE all-reduce (bsi_b200/optim.py) and the 1/world_size of
// DistributedDataParallel (bsi/tasks/bsi.py:163-166) is applied here as grad_scale instead of in a pass of its own.
#include "common.cuh"

namespace bsi {

constexpr int kOThreads = 256;
constexpr int kSumsqBlocks = 1184;  // 8 x 148: fixed grid -> fixed summation order -> deterministic norm

// torch.lerp (ATen/native/Lerp.h): weight < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w), the multiply-add contracted to one FMA
__device__ __forceinline__ float torch_lerp(float a, float b, float w) {
    const float diff = __fsub_rn(b, a);
    return w < 0.5f ? __fmaf_rn(w, diff, a) : __fmaf_rn(-diff, __fsub_rn(1.0f, w), b);
}

__global__ void __launch_bounds__(kOThreads) k_grad_sumsq_partial(float* __restrict__ partial, const float* __restrict__ g, int64_t n4) {
    __shared__ float red[kOThreads / 32];
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float acc = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * kOThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kOThreads) {
        const float4 v = __ldg(g4 + i);
        acc = fmaf(v.x, v.x, acc), acc = fmaf(v.y, v.y, acc), acc = fmaf(v.z, v.z, acc), acc = fmaf(v.w, v.w, acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < kOThreads / 32 ? red[threadIdx.x] : 0.0f;
        t = warp_sum(t);
        if (threadIdx.x == 0) partial[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kOThreads) k_grad_sumsq_final(float* __restrict__ out, const float* __restrict__ partial, int nblocks) {
    __shared__ double red[kOThreads / 32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += kOThreads) acc += (double)partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < kOThreads / 32; ++i) t += red[i];
        out[0] = (float)t;
    }
}

struct StepConsts {
    float decay_mul;    // 1 - lr*wd
    float w1;           // 1 - beta1
    float beta2, w2;    // beta2, 1 - beta2
    float bc2_sqrt;     // sqrt(1 - beta2^t)
    float eps;
    float neg_step;     // -lr / (1 - beta1^t)
    float max_norm;     // <= 0: no clipping
    float grad_scale;   // 1/world_size after a sum all-reduce of the gradient arena (DDP's average), else 1
    float ema_weight;   // 1 - decay
    int ema_mode;       // 0 none, 1 copy, 2 lerp
    int zero_grad;
};

__device__ __forceinline__ float adamw_one(float& p, float g, float& m, float& v, const StepConsts& c, float coef) {
    g = __fmul_rn(__fmul_rn(g, c.grad_scale), coef);
    p = __fmul_rn(p, c.decay_mul);
    m = torch_lerp(m, g, c.w1);
    v = __fmaf_rn(__fmul_rn(c.w2, g), g, __fmul_rn(v, c.beta2));  // addcmul: (value*g)*g + v*beta2 in one FMA
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), c.bc2_sqrt), c.eps);
    p = __fmaf_rn(c.neg_step, __fdiv_rn(m, denom), p);  // addcdiv
    return p;
}

__global__ void __launch_bounds__(kOThreads)
    k_adamw_ema(float* __restrict__ param, float* __restrict__ grad, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                float* __restrict__ ema, __nv_bfloat16* __restrict__ param_bf16, const float* __restrict__ grad_sumsq, int64_t n4,
                const StepConsts c) {
    float coef = 1.0f;
    if (c.max_norm > 0.0f) {
        const float total_norm = __fmul_rn(__fsqrt_rn(grad_sumsq[0]), c.grad_scale);  // norm of the averaged gradient
        coef = fminf(__fdiv_rn(c.max_norm, __fadd_rn(total_norm, 1e-6f)), 1.0f);
    }
    float4* p4 = reinterpret_cast<float4*>(param);
    float4* g4 = reinterpret_cast<float4*>(grad);
    float4* m4 = reinterpret_cast<float4*>(exp_avg);
    float4* v4 = reinterpret_cast<float4*>(exp_avg_sq);
    float4* e4 = reinterpret_cast<float4*>(ema);
    uint2* b4 = reinterpret_cast<uint2*>(param_bf16);
    for (int64_t i = (int64_t)blockIdx.x * kOThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kOThreads) {
        float4 p = p4[i], m = m4[i], v = v4[i];
        const float4 g = g4[i];
        adamw_one(p.x, g.x, m.x, v.x, c, coef);
        adamw_one(p.y, g.y, m.y, v.y, c, coef);
        adamw_one(p.z, g.z, m.z, v.z, c, coef);
        adamw_one(p.w, g.w, m.w, v.w, c, coef);
        p4[i] = p, m4[i] = m, v4[i] = v;
        if (c.zero_grad) g4[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (c.ema_mode == 1) {
            e4[i] = p;
        } else if (c.ema_mode == 2) {
            float4 e = e4[i];
            e.x = torch_lerp(e.x, p.x, c.ema_weight), e.y = torch_lerp(e.y, p.y, c.ema_weight);
            e.z = torch_lerp(e.z, p.z, c.ema_weight), e.w = torch_lerp(e.w, p.w, c.ema_weight);
            e4[i] = e;
        }
        if (param_bf16) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
            b4[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        }
    }
}

// EMA on its own (bsi/tasks/ema_pytorch.py:316-341 when the optimizer is not ours): mode 1 copy, mode 2 lerp
__global__ void __launch_bounds__(kOThreads)
    k_ema_update(float* __restrict__ ema, const float* __restrict__ param, int64_t n4, float weight, int mode) {
    float4* e4 = reinterpret_cast<float4*>(ema);
    const float4* p4 = reinterpret_cast<const float4*>(param);
    for (int64_t i = (int64_t)blockIdx.x * kOThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kOThreads) {
        const float4 p = p4[i];
        if (mode == 1) {
            e4[i] = p;
        } else {
            float4 e = e4[i];
            e.x = torch_lerp(e.x, p.x, weight), e.y = torch_lerp(e.y, p.y, weight);
            e.z = torch_lerp(e.z, p.z, weight), e.w = torch_lerp(e.w, p.w, weight);
            e4[i] = e;
        }
    }
}

static unsigned stream_grid(int64_t n4) {
    const int64_t need = (n4 + kOThreads - 1) / kOThreads, cap = (int64_t)sm_count() * 8;
    return (unsigned)(need < cap ? need : cap);
}

}  // namespace bsi

using namespace bsi;

extern "C" {

int bsi_grad_sumsq(float* out, float* workspace, const float* grad, int64_t numel, void* stream) {
    BSI_CHECK_ARG(out && workspace && grad && numel > 0, "bsi_grad_sumsq: null pointer or empty gradient");
    BSI_CHECK_ARG(numel % 4 == 0, "bsi_grad_sumsq: numel (%lld) must be a multiple of 4 (pad the arena)", (long long)numel);
    k_grad_sumsq_partial<<<kSumsqBlocks, kOThreads, 0, (cudaStream_t)stream>>>(workspace, grad, numel / 4);
    BSI_LAUNCH_OK("k_grad_sumsq_partial");
    k_grad_sumsq_final<<<1, kOThreads, 0, (cudaStream_t)stream>>>(out, workspace, kSumsqBlocks);
    BSI_LAUNCH_OK("k_grad_sumsq_final");
    return BSI_OK;
}

int bsi_adamw_ema_step(const bsi_adamw_args* a, void* stream) {
    BSI_CHECK_ARG(a && a->param && a->grad && a->exp_avg && a->exp_avg_sq && a->numel > 0, "bsi_adamw_ema_step: null pointer or empty arena");
    BSI_CHECK_ARG(a->numel % 4 == 0, "bsi_adamw_ema_step: numel (%lld) must be a multiple of 4 (pad the arena)", (long long)a->numel);
    BSI_CHECK_ARG(a->step >= 1, "bsi_adamw_ema_step: step counts from 1 (got %lld)", (long long)a->step);
    BSI_CHECK_ARG(a->max_norm <= 0.0f || a->grad_sumsq, "bsi_adamw_ema_step: clipping needs the device sum of squares (bsi_grad_sumsq)");
    BSI_CHECK_ARG(a->ema_mode >= 0 && a->ema_mode <= 2 && (a->ema_mode == 0 || a->ema), "bsi_adamw_ema_step: bad EMA mode / missing EMA arena");
    // scalar prologue in double like the Python scalars of torch.optim.adamw._single_tensor_adamw, then rounded to fp32 once
    StepConsts c;
    c.decay_mul = (float)(1.0 - a->lr * a->weight_decay);
    c.w1 = (float)(1.0 - a->beta1);
    c.beta2 = (float)a->beta2, c.w2 = (float)(1.0 - a->beta2);
    const double bc1 = 1.0 - pow(a->beta1, (double)a->step), bc2 = 1.0 - pow(a->beta2, (double)a->step);
    c.bc2_sqrt = (float)sqrt(bc2);
    c.eps = (float)a->eps;
    c.neg_step = (float)(-(a->lr / bc1));
    c.max_norm = a->max_norm;
    c.grad_scale = a->grad_scale > 0.0f ? a->grad_scale : 1.0f;
    c.ema_weight = a->ema_weight, c.ema_mode = a->ema_mode, c.zero_grad = a->zero_grad;
    k_adamw_ema<<<stream_grid(a->numel / 4), kOThreads, 0, (cudaStream_t)stream>>>(a->param, a->grad, a->exp_avg, a->exp_avg_sq, a->ema,
                                                                                   (__nv_bfloat16*)a->param_bf16, a->grad_sumsq, a->numel / 4, c);
    BSI_LAUNCH_OK("k_adamw_ema");
    return BSI_OK;
}

int bsi_ema_update(float* ema, const float* param, int64_t numel, float weight, int32_t mode, void* stream) {
    BSI_CHECK_ARG(ema && param && numel > 0 && numel % 4 == 0, "bsi_ema_update: null pointer, or numel (%lld) not a positive multiple of 4", (long long)numel);
    BSI_CHECK_ARG(mode == 1 || mode == 2, "bsi_ema_update: mode must be 1 (copy) or 2 (lerp)");
    k_ema_update<<<stream_grid(numel / 4), kOThreads, 0, (cudaStream_t)stream>>>(ema, param, numel / 4, weight, mode);
    BSI_LAUNCH_OK("k_ema_update");
    return BSI_OK;
}

int32_t bsi_grad_sumsq_workspace_floats(void) { return kSumsqBlocks; }

}  // extern "C"
