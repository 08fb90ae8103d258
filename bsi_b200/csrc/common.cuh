// Shared host/device helpers for the bsi_b200 CUDA library.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/bsi_b200.h"

namespace bsi {

void set_error(const char* fmt, ...);
void count_launch();  // host-side counter of kernels launched by this library (bsi_launch_counter)

#define BSI_CHECK_ARG(cond, ...)             \
    do {                                     \
        if (!(cond)) {                       \
            ::bsi::set_error(__VA_ARGS__);   \
            return BSI_ERR_INVALID_ARGUMENT; \
        }                                    \
    } while (0)

#define BSI_CUDA_OK(expr)                                                                        \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            ::bsi::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return BSI_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

#define BSI_LAUNCH_OK(name)                                                                \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            ::bsi::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));     \
            return BSI_ERR_CUDA;                                                           \
        }                                                                                  \
        ::bsi::count_launch();                                                             \
    } while (0)

int sm_count();
// Programmatic dependent launch (PDL): kernels that start with pdl_prologue_done() may be launched with the
// programmatic-stream-serialization attribute, so their prologue (barrier init, TMEM allocation, descriptor prefetch, launch
// latency) overlaps the tail of the previous kernel in the stream / graph.  Off by default (measured slower inside the
// captured sampler graph); BSI_PDL=1 in the environment enables it.
bool use_pdl();
void fill_pdl_attr(cudaLaunchAttribute* attr);

// Opt a kernel into `bytes` of dynamic shared memory, once per (call site, device).  The limit is only ever raised: another
// call site of the same kernel may already have asked for more (the cache below is per call site, the attribute per kernel).
#define BSI_ENSURE_SMEM(kernel, bytes)                                                                            \
    do {                                                                                                          \
        static int _done[64] = {0};                                                                               \
        int _dev = 0;                                                                                             \
        BSI_CUDA_OK(cudaGetDevice(&_dev));                                                                        \
        if (_dev >= 0 && _dev < 64 && _done[_dev] < (bytes)) {                                                    \
            cudaFuncAttributes _fa;                                                                               \
            BSI_CUDA_OK(cudaFuncGetAttributes(&_fa, kernel));                                                     \
            const int _want = (int)(bytes) > _fa.maxDynamicSharedSizeBytes ? (int)(bytes) : _fa.maxDynamicSharedSizeBytes; \
            BSI_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, _want));        \
            /* every kernel that opts in to large shared memory asks for the same (maximal) L1/shared split: a kernel with a moderate   \
             * request after a 227 KB one otherwise makes the SMs reconfigure their carve-out (58 us idle per occurrence, r02 timeline) */ \
            cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);   \
            _done[_dev] = (bytes);                                                                                \
        }                                                                                                         \
    } while (0)

// Resolve a bsi_rowref for (sample, step).
__device__ __forceinline__ float rowref_at(const bsi_rowref& r, int64_t sample, int step) {
    return r.base[(int64_t)step * r.step_stride + sample * r.sample_stride];
}
__device__ __forceinline__ const float* rowref_ptr(const bsi_rowref& r, int64_t sample, int step) {
    return r.base + (int64_t)step * r.step_stride + sample * r.sample_stride;
}

// Let the next kernel of the stream begin its prologue, then wait until every kernel this one depends on has completed and
// its memory is visible.  Must precede the first access to global memory written by earlier kernels.  No-op without PDL.
__device__ __forceinline__ void pdl_prologue_done() {
    asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Counter-based dropout mask (training path): keep element `idx` of the stream keyed by `seed` with probability 1 - p, where
// murmur3's 32-bit finaliser over a Weyl-multiplied index: stateless, so the forward and both backward
// kernels regenerate the identical mask from (seed, index) instead of storing it (tests/helpers.py restates it in Python).
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x85ebca6bu;
    x ^= x >> 13;
    x *= 0xc2b2ae35u;
    x ^= x >> 16;
    return x;
}
// 16 random bits per element: elements 2i and 2i + 1 share the hash of pair i (round 2: the mask generation was 40 % of the tcgen05
// attention forward's instructions; a thread that walks consecutive elements now hashes once per two).  thresh = floor(p * 2^16), i.e. the
// drop probability is quantised to 1/65536 (p = 0.05 -> 0.04999); the 1 / (1 - p) rescale uses the caller's p.
__host__ __device__ __forceinline__ uint32_t dropout_bits(uint32_t seed, uint32_t pair) { return mix32(pair * 0x9E3779B9u + seed); }
__host__ __device__ __forceinline__ bool dropout_keep(uint32_t seed, uint32_t idx, uint32_t thresh) {
    const uint32_t h = dropout_bits(seed, idx >> 1);
    return ((idx & 1u) ? (h >> 16) : (h & 0xffffu)) >= thresh;
}
// keep decisions of elements idx_even and idx_even + 1 from one hash
__host__ __device__ __forceinline__ void dropout_keep_pair(uint32_t seed, uint32_t idx_even, uint32_t thresh, bool& k0, bool& k1) {
    const uint32_t h = dropout_bits(seed, idx_even >> 1);
    k0 = (h & 0xffffu) >= thresh, k1 = (h >> 16) >= thresh;
}
__host__ __device__ __forceinline__ uint32_t dropout_thresh(float p) { return p <= 0.0f ? 0u : (uint32_t)((double)p * 65536.0); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// Philox4x32-10 (Salmon et al., SC'11), same constants as Random123.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
// Four N(0,1) values for (seed, global sample, draw, element-quad): Box-Muller on (r0,r1) and (r2,r3).
// The transcendental part uses the SFU approximations (lg2 / sqrt / sin / cos .approx): with the library versions the generator
// costs ~230 instructions per quad and every kernel that draws noise is instruction-bound (ncu r02: k_q_sample 77 % issue-active at
// 0.34 of the HBM roof).  Absolute error of a sample <= 4e-5 (only where |z| < 0.01, from lg2.approx near 1), typically 1e-6:
// far below what any statistic of the sampler resolves; the oracle's fp64 restatement agrees to 1e-4 (tests/test_gpu_elementwise.py).
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t sample, uint32_t draw, uint32_t quad) {
    uint4 r = philox4x32_10(make_uint4(quad, draw, (uint32_t)sample, (uint32_t)(sample >> 32)),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float k2m32 = 2.3283064365386963e-10f;  // 2^-32
    float u0 = ((float)r.x + 0.5f) * k2m32, u1 = ((float)r.y + 0.5f) * k2m32;
    float u2 = ((float)r.z + 0.5f) * k2m32, u3 = ((float)r.w + 0.5f) * k2m32;
    // (float)r can round up to 2^32 -> u = 1 -> log = 0: fine (radius 0).  u > 0 always.
    float rad0, rad1;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad0) : "f"(-2.0f * __logf(u0)));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad1) : "f"(-2.0f * __logf(u2)));
    // angle 2 pi u in [0, 2 pi): evaluate at x = 2 pi u - pi in [-pi, pi), where sin/cos.approx are accurate to 2^-21; sin(x + pi) = -sin x
    const float x0 = fmaf(6.283185307179586f, u1, -3.141592653589793f), x1 = fmaf(6.283185307179586f, u3, -3.141592653589793f);
    const float s0 = -__sinf(x0), c0 = -__cosf(x0), s1 = -__sinf(x1), c1 = -__cosf(x1);
    return make_float4(rad0 * c0, rad0 * s0, rad1 * c1, rad1 * s1);
}

}  // namespace bsi
