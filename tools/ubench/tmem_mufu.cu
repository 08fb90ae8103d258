// Microbenchmark: TMEM read port (tcgen05.ld) vs SFU (ex2) throughput and how well they overlap, one 512-thread CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I bsi_b200/csrc tools/ubench/tmem_mufu.cu -o tools/ubench/tmem_mufu.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"


__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, int active_warps, long long* cycles, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { ptx::tmem_alloc<1>(&slot, 512); ptx::tmem_relinquish<1>(); }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t t = slot + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    float acc = 0.f;
    uint32_t a[16], b[16], c[16], d[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = b[i] = c[i] = d[i] = threadIdx.x + i;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < active_warps) {
        for (int it = 0; it < iters; ++it) {
            if (MODE == 0) {  // loads only: 64 columns per thread per iteration
                ptx::tmem_ld_32x32b_x16(t, a); ptx::tmem_ld_32x32b_x16(t + 16, b); ptx::tmem_ld_32x32b_x16(t + 32, c); ptx::tmem_ld_32x32b_x16(t + 48, d);
                ptx::tmem_ld_wait();
                acc += __uint_as_float(a[0] ^ b[1] ^ c[2] ^ d[3]);
            } else if (MODE == 1) {  // 64 ex2 only
#pragma unroll
                for (int i = 0; i < 16; ++i) acc += ex2(__uint_as_float(a[i]) * 1e-30f) + ex2(__uint_as_float(b[i]) * 1e-30f) + ex2(__uint_as_float(c[i]) * 1e-30f) + ex2(__uint_as_float(d[i]) * 1e-30f);
            } else if (MODE == 2) {  // interleaved: ex2 over a chunk, then reload that chunk
#pragma unroll
                for (int i = 0; i < 16; ++i) acc += ex2(__uint_as_float(a[i]) * 1e-30f);
                ptx::tmem_ld_32x32b_x16(t, a);
#pragma unroll
                for (int i = 0; i < 16; ++i) acc += ex2(__uint_as_float(b[i]) * 1e-30f);
                ptx::tmem_ld_32x32b_x16(t + 16, b);
#pragma unroll
                for (int i = 0; i < 16; ++i) acc += ex2(__uint_as_float(c[i]) * 1e-30f);
                ptx::tmem_ld_32x32b_x16(t + 32, c);
#pragma unroll
                for (int i = 0; i < 16; ++i) acc += ex2(__uint_as_float(d[i]) * 1e-30f);
                ptx::tmem_ld_32x32b_x16(t + 48, d);
                ptx::tmem_ld_wait();
            } else if (MODE == 3) {  // serial phases: all loads, wait, then all ex2
                ptx::tmem_ld_32x32b_x16(t, a); ptx::tmem_ld_32x32b_x16(t + 16, b); ptx::tmem_ld_32x32b_x16(t + 32, c); ptx::tmem_ld_32x32b_x16(t + 48, d);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) acc += ex2(__uint_as_float(a[i]) * 1e-30f) + ex2(__uint_as_float(b[i]) * 1e-30f) + ex2(__uint_as_float(c[i]) * 1e-30f) + ex2(__uint_as_float(d[i]) * 1e-30f);
            } else if (MODE == 4) {  // half the warps load while the other half computes (roles swap every iteration)
                if (((warp >> 1) + it) & 1) {
                    ptx::tmem_ld_32x32b_x16(t, a); ptx::tmem_ld_32x32b_x16(t + 16, b); ptx::tmem_ld_32x32b_x16(t + 32, c); ptx::tmem_ld_32x32b_x16(t + 48, d);
                    ptx::tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc += ex2(__uint_as_float(a[i]) * 1e-30f) + ex2(__uint_as_float(b[i]) * 1e-30f) + ex2(__uint_as_float(c[i]) * 1e-30f) + ex2(__uint_as_float(d[i]) * 1e-30f);
                }
            } else if (MODE >= 6) {  // the attention exp2 phase: 4 chunks of (16 x (ffma, ex2, add), 8 packs, st x8 [, wait::st] [, ld x16])
                const float sc = 0.18f, mo = 3.0f;
                float s0 = 0.f, s1 = 0.f;
                uint32_t* arr[4] = {a, b, c, d};
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t p[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float p0 = ex2(fmaf(__uint_as_float(arr[ch][2 * i]), sc, -mo)), p1 = ex2(fmaf(__uint_as_float(arr[ch][2 * i + 1]), sc, -mo));
                        s0 += p0, s1 += p1;
                        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p[i]) : "f"(p1), "f"(p0));
                    }
                    ptx::tmem_st_32x32b_x8(t + 256 + ch * 8, p);
                    if (MODE >= 8) { ptx::tmem_st_wait(); ptx::tc_fence_before(); __syncwarp(); }
                    if (MODE >= 7) ptx::tmem_ld_32x32b_x16(t + ch * 16, *reinterpret_cast<uint32_t(*)[16]>(arr[ch]));
                }
                acc += s0 + s1;
                if (MODE == 6) ptx::tmem_st_wait();
                if (MODE >= 7) ptx::tmem_ld_wait();
            } else if (MODE == 5) {  // loads only, x32 shape
                uint32_t e[32], f[32];
                ptx::tmem_ld_32x32b_x32(t, e); ptx::tmem_ld_32x32b_x32(t + 32, f);
                ptx::tmem_ld_wait();
                acc += __uint_as_float(e[0] ^ f[1]);
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<1>(slot, 512); }
}

template <int MODE>
void run(const char* name, int warps) {
    long long* cyc; float* sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
    const int iters = 2000;
    k<MODE><<<148, 512>>>(10, warps, cyc, sink);
    k<MODE><<<148, 512>>>(iters, warps, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double per_iter = avg / iters;
    printf("%-28s warps=%2d  %8.1f cyc/iter  -> %6.1f B/clk/SM tmem, %5.2f ex2/clk/SM  (%s)\n", name, warps, per_iter,
           warps * 32 * 64 * 4.0 / per_iter, warps * 32 * 64.0 / per_iter, cudaGetErrorString(e));
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    for (int w : {4, 8, 16}) run<0>("ld x16 only", w);
    for (int w : {4, 16}) run<5>("ld x32 only", w);
    for (int w : {4, 8, 16}) run<1>("ex2 only", w);
    run<3>("serial ld then ex2", 16);
    run<2>("interleaved ld/ex2", 16);
    run<4>("half load / half ex2 (numbers x0.5)", 16);
    run<6>("exp phase: ex2+pack+st", 16);
    run<7>("exp phase + ld x16", 16);
    run<8>("exp phase + ld + wait::st", 16);
    return 0;
}
