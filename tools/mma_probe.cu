// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, bf16 operands) for the tile shapes / operand sources the attention and GEMM
// kernels use -- how many clocks one MMA takes when a long train of them is issued back to back by one thread, with the A operand in
// shared memory (ss) or in tensor memory (ts).  The question it answers: is a cta_group::1 MMA limited by the 4096 MAC/clk tensor
// pipe, or by the shared-memory operand fetch?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I bsi_b200/csrc -o tools/mma_probe.bin tools/mma_probe.cu -lcuda
//   tools/mma_probe.bin [ctas]
#include <cstdio>
#include <cstdlib>

#include "ptx_sm100.cuh"



constexpr int kReps = 512;
constexpr int kSmemBytes = 200 * 1024;

struct Case {
    int M, N, a_tmem, a_mn, b_mn, ndst;  // ndst: consecutive MMAs rotate over this many independent accumulators
};

// Everything about a case is a template parameter and the 16-MMA body is fully unrolled with precomputed descriptors, so that the
// issuing thread spends ~3 instructions per MMA (the first version of this probe rebuilt the descriptors at run time and measured
// 135 clk per MMA for EVERY shape: the single issuing thread, not the tensor pipe, was the limit -- which is also what the attention
// kernels' control warps suffered from).
template <int M, int N, int ATMEM, int AMN, int BMN, int NDST>
__global__ void __launch_bounds__(128, 1) k_probe(long long* out, int ncase_index) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (kSmemBytes - 2048) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) {
            ptx::mbar_init(&bar, 1);
            ptx::fence_mbar_init();
        }
        __syncwarp();
        ptx::tmem_alloc<1>(&slot, 512);
        ptx::tmem_relinquish<1>();
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = ptx::umma_idesc_bf16(M, N, AMN, BMN);
        const uint32_t base = ptx::smem_u32(smem);
        // A tiles in the first 64 KB, B tiles in the next 128 KB; consecutive MMAs read different 16-element k-slices, as k-steps of a tile do
        const uint64_t da0 = AMN ? ptx::umma_desc_mn_sw128(base, 8192, 1024) : ptx::umma_desc_k_sw128(base);
        const uint64_t db0 = BMN ? ptx::umma_desc_mn_sw128(base + 65536, 8192, 1024) : ptx::umma_desc_k_sw128(base + 65536);
        long long t0 = 0, t1 = 0;
        for (int pass = 0; pass < 2; ++pass) {  // pass 0 warms up
            t0 = clock64();
#pragma unroll 1
            for (int r = 0; r < kReps / 16; ++r) {
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    constexpr int kslice_k = 2, kslice_mn = 2048 >> 4;  // descriptor address units (16 B)
                    const uint64_t da = da0 + (uint64_t)((u & 3) * (AMN ? kslice_mn : kslice_k) + (u >> 2) * (16384 >> 4));
                    const uint64_t db = db0 + (uint64_t)((u & 3) * (BMN ? kslice_mn : kslice_k) + (u >> 2) * (32768 >> 4));
                    if constexpr (ATMEM)
                        ptx::umma_bf16_ts(tmem + 256 + (u % NDST) * 64, tmem + (u & 7) * 8, db, idesc, 1u);
                    else
                        ptx::umma_bf16_ss<1>(tmem + 256 + (u % NDST) * 64, da, db, idesc, 1u);
                }
            }
            ptx::umma_commit<1>(&bar);
            ptx::mbar_wait(&bar, pass & 1);
            t1 = clock64();
        }
        if (blockIdx.x == 0) out[ncase_index] = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem, 512);
    }
}

// The MMA stream of one 64-query sub-step of the attention backward (attention_bwd_sm100.cu, k_attention_bwd_tc2), issued back to back:
// 8 score MMAs (128 x 64, K-major x K-major, two accumulators) + 4 x {dV (A from TMEM, B MN-major), dK (A K-major, B MN-major),
// 2 x dQ (M = 64, A and B MN-major)}.  GROUPED = 1 issues the 4 dV, then the 4 dK, then the 8 dQ instead of interleaving them.
template <int GROUPED>
__global__ void __launch_bounds__(128, 1) k_probe_mix(long long* out, int idx) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (kSmemBytes - 2048) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) {
            ptx::mbar_init(&bar, 1);
            ptx::fence_mbar_init();
        }
        __syncwarp();
        ptx::tmem_alloc<1>(&slot, 512);
        ptx::tmem_relinquish<1>();
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, 64), idesc_kv = ptx::umma_idesc_bf16(128, 64, 0, 1), idesc_q = ptx::umma_idesc_bf16(64, 64, 1, 1);
        const uint32_t base = ptx::smem_u32(smem);
        const uint64_t kA = ptx::umma_desc_k_sw128(base), kB = ptx::umma_desc_k_sw128(base + 65536);
        const uint64_t mA = ptx::umma_desc_mn_sw128(base + 32768, 16384, 1024), mB = ptx::umma_desc_mn_sw128(base + 98304, 8192, 1024);
        long long t0 = 0, t1 = 0;
        for (int pass = 0; pass < 2; ++pass) {
            t0 = clock64();
#pragma unroll 1
            for (int r = 0; r < 64; ++r) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    ptx::umma_bf16_ss<1>(tmem + 0, kA + 2 * k, kB + 2 * k, idesc_s, 1u);
                    ptx::umma_bf16_ss<1>(tmem + 64, kA + 1024 + 2 * k, kB + 1024 + 2 * k, idesc_s, 1u);
                }
                if constexpr (GROUPED) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_bf16_ts(tmem + 256, tmem + 128 + 8 * k, mB + k * 128, idesc_kv, 1u);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_bf16_ss<1>(tmem + 320, kA + 2 * k, mB + 512 + k * 128, idesc_kv, 1u);
#pragma unroll
                    for (int k = 0; k < 8; ++k) ptx::umma_bf16_ss<1>(tmem + 384, mA + k * 128, mB + k * 128, idesc_q, 1u);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        ptx::umma_bf16_ts(tmem + 256, tmem + 128 + 8 * k, mB + k * 128, idesc_kv, 1u);
                        ptx::umma_bf16_ss<1>(tmem + 320, kA + 2 * k, mB + 512 + k * 128, idesc_kv, 1u);
                        ptx::umma_bf16_ss<1>(tmem + 384, mA + (2 * k) * 128, mB + (2 * k) * 128, idesc_q, 1u);
                        ptx::umma_bf16_ss<1>(tmem + 384, mA + (2 * k + 1) * 128, mB + (2 * k + 1) * 128, idesc_q, 1u);
                    }
                }
            }
            ptx::umma_commit<1>(&bar);
            ptx::mbar_wait(&bar, pass & 1);
            t1 = clock64();
        }
        if (blockIdx.x == 0) out[idx] = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem, 512);
    }
}

template <int M, int N, int ATMEM, int AMN, int BMN, int NDST>
static void run(long long* d, int ctas, int idx, Case* cases) {
    cases[idx] = Case{M, N, ATMEM, AMN, BMN, NDST};
    cudaFuncSetAttribute(k_probe<M, N, ATMEM, AMN, BMN, NDST>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    k_probe<M, N, ATMEM, AMN, BMN, NDST><<<ctas, 128, kSmemBytes>>>(d, idx);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("{\"error\": \"%s\", \"case\": %d}\n", cudaGetErrorString(e), idx);
        exit(1);
    }
}

int main(int argc, char** argv) {
    const int ctas = argc > 1 ? atoi(argv[1]) : 1;
    Case cases[32];
    long long* d;
    cudaMalloc(&d, 32 * sizeof(long long));
    int n = 0;
    // ss, both K-major, one accumulator (a dependent chain, like the k-steps of one tile)
    run<128, 64, 0, 0, 0, 1>(d, ctas, n++, cases);
    run<128, 128, 0, 0, 0, 1>(d, ctas, n++, cases);
    run<128, 256, 0, 0, 0, 1>(d, ctas, n++, cases);
    // ts: A from TMEM
    run<128, 64, 1, 0, 0, 1>(d, ctas, n++, cases);
    run<128, 128, 1, 0, 0, 1>(d, ctas, n++, cases);
    run<128, 256, 1, 0, 0, 1>(d, ctas, n++, cases);
    // independent accumulators, round robin
    run<128, 64, 0, 0, 0, 2>(d, ctas, n++, cases);
    run<128, 64, 0, 0, 0, 4>(d, ctas, n++, cases);
    run<128, 64, 1, 0, 0, 2>(d, ctas, n++, cases);
    run<128, 64, 1, 0, 0, 4>(d, ctas, n++, cases);
    run<128, 128, 0, 0, 0, 2>(d, ctas, n++, cases);
    // B MN-major (dK / dV of the attention backward), A MN-major (dQ) with M = 128 and M = 64
    run<128, 64, 0, 0, 1, 1>(d, ctas, n++, cases);
    run<128, 64, 1, 0, 1, 1>(d, ctas, n++, cases);
    run<128, 64, 0, 1, 1, 1>(d, ctas, n++, cases);
    run<64, 64, 0, 1, 1, 1>(d, ctas, n++, cases);
    run<64, 64, 0, 1, 1, 2>(d, ctas, n++, cases);
    run<64, 128, 0, 0, 0, 1>(d, ctas, n++, cases);
    long long h[32];
    cudaMemcpy(h, d, n * sizeof(long long), cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; ++i) {
        const Case& c = cases[i];
        const double clk = (double)h[i] / kReps;
        const double floor_clk = (c.M > 128 ? c.M : 128) * c.N / 256.0;
        const double a_bytes = c.a_tmem ? 0 : c.M * 32.0, b_bytes = c.N * 32.0;
        printf("{\"ctas\": %d, \"accumulators\": %d, \"M\": %d, \"N\": %d, \"A\": \"%s\", \"B\": \"%s\", \"clk_per_mma\": %.1f, \"floor_clk\": %.0f, \"mac_per_clk\": %.0f, \"smem_B_per_clk\": %.1f}\n",
               ctas, c.ndst, c.M, c.N, c.a_tmem ? "tmem" : (c.a_mn ? "smem-mn" : "smem-k"), c.b_mn ? "smem-mn" : "smem-k", clk, floor_clk, c.M * c.N * 16.0 / clk,
               (a_bytes + b_bytes) / clk);
    }
    for (int grouped = 0; grouped < 2; ++grouped) {
        if (grouped) {
            cudaFuncSetAttribute(k_probe_mix<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
            k_probe_mix<1><<<ctas, 128, kSmemBytes>>>(d, 0);
        } else {
            cudaFuncSetAttribute(k_probe_mix<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
            k_probe_mix<0><<<ctas, 128, kSmemBytes>>>(d, 0);
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("{\"error\": \"%s\", \"case\": \"mix\"}\n", cudaGetErrorString(e));
            return 1;
        }
        cudaMemcpy(h, d, sizeof(long long), cudaMemcpyDeviceToHost);
        printf("{\"ctas\": %d, \"case\": \"attention-backward sub-step stream (24 MMAs), %s\", \"clk_per_substep\": %.0f, \"clk_per_mma\": %.1f}\n", ctas,
               grouped ? "grouped by accumulator" : "interleaved", (double)h[0] / 64, (double)h[0] / 64 / 24);
    }
    return 0;
}
