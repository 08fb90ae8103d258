// Weight-gradient GEMM on tcgen05 (backward of every nn.Linear of bsi/models/dit.py; groundwork for SURVEY §8 a23):
//     dW[n][k] += sum_m dY[m][n] * X[m][k]          dY bf16 [M][N], X bf16 [M][K], dW fp32 [N][K]
// The contraction runs over the token dimension m, which is the slow (row) index of both operands, so both are fed to the
// tensor core MN-major: a TMA box {64 columns, 64 rows} of dY (or X) lands as eight 1024-byte swizzle atoms whose rows are
// K (= m) and whose 128-byte lines are the MN dimension -- no transposed copies of the activations are ever made.
// CTA pair (cta_group::2): UMMA 256 (n) x 256 (k) x 16 (m); each CTA loads its 128 n-columns of dY and its 128 k-columns of X.
// The M range is split over `splits` work items per output tile so that all SMs are busy even for the 1024x1024 projection;
// partial tiles are combined in fp32 by TMA reduce-add stores (cp.reduce.async.bulk.tensor .add), which is also autograd's
// "+=" into the gradient arena.  Tensor-bound: 2*M*N*K flop.
#include <cuda.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace bsi {

int make_tile_map(CUtensorMap* map, const void* base, int esize, int64_t rows, int64_t cols, int64_t ld, int64_t batch, int64_t batch_stride,
                  int box_rows);

namespace wg {
constexpr int BM = 128;               // dW rows (n) per CTA; 256 per pair
constexpr int BN = 256;               // dW columns (k) per tile; each CTA loads 128 of them
constexpr int BK = 64, UMMA_K = 16;   // contraction (m) block
constexpr int kAtomBytes = BK * 128;  // 64 m-rows x 64 MN-elements
constexpr int kABytes = 2 * kAtomBytes, kBBytes = 2 * kAtomBytes, kStageBytes = kABytes + kBBytes;
constexpr int kStages = 6;
constexpr int kEpiBufBytes = BM * 128;
constexpr int kThreads = 256;
constexpr int kSmem = kStages * kStageBytes + 2 * kEpiBufBytes + 1024 + 256;
}  // namespace wg

__global__ void __launch_bounds__(wg::kThreads, 1)
    k_wgrad_bf16(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dw,
                 const int n_tiles, const int k_tiles, const int splits, const int kb_per_split, const int total_kb) {
    using namespace wg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_buf = smem + kStages * kStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_buf + 2 * kEpiBufBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;   // [2]: one per accumulator buffer
    uint64_t* tmem_empty = tmem_full + 2;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta_rank = (int)ptx::cluster_ctarank();
    const int worker = blockIdx.x / 2, num_workers = gridDim.x / 2;
    const int total_items = n_tiles * k_tiles * splits;
    // Work distribution.  splits > 0: items (tile, split) dealt round-robin, every split kb_per_split k-blocks long (round 1: whole items
    // on 74 CTA pairs quantise badly -- 64 tiles x 3 splits = 2.6 waves, paid as 3).  splits == 0: every pair gets the same number of
    // block-steps, laid out so that pairs running at the same time sweep the SAME token range: the operands are re-read once per output
    // tile (2 GB for the MLP shapes against 0.3 GB of unique data), which only works out of L2 if the pairs move through M together
    // -- plain contiguous "stream-K" ranges start every pair at a different token offset and fall to DRAM speed (measured 267 vs 207 us).
    //   heads: pair w < tiles * f (f = pairs / tiles) takes k-blocks [seg * H, (seg + 1) * H) of tile w % tiles, seg = w / tiles
    //   tails: the k-blocks [f * H, total_kb) of all tiles, cut into equal contiguous ranges for the remaining pairs
    // with H = (tiles * total_kb) / pairs.  Every flush is a reduce-add into dW anyway (dW +=), so partial tiles cost nothing extra.
    const int tiles = n_tiles * k_tiles;
    const int sk_f = num_workers / tiles, sk_heads = tiles * sk_f, sk_left = num_workers - sk_heads;
    const int sk_H = kb_per_split;
    const int tail0 = sk_left == 0 ? total_kb : sk_f * sk_H, tail_len = total_kb - tail0;
    const long long tail_total = (long long)tiles * tail_len;
    const int tw = worker - sk_heads;
    const long long sk_lo = sk_left > 0 && tw >= 0 ? tail_total * tw / sk_left : 0, sk_hi = sk_left > 0 && tw >= 0 ? tail_total * (tw + 1) / sk_left : 0;
    struct Cursor {
        long long pos;
        int item;
    };
    auto first = [&]() { return Cursor{sk_lo, splits > 0 ? worker : 0}; };
    // next segment of this pair: output tile origin (n0, k0) and k-block range [kb0, kb1); false when the pair is done
    auto next = [&](Cursor& c, int& n0, int& k0, int& kb0, int& kb1) {
        int t;
        if (splits > 0) {
            if (c.item >= total_items) return false;
            const int sp = c.item % splits;
            t = c.item / splits;
            kb0 = sp * kb_per_split, kb1 = min(kb0 + kb_per_split, total_kb);
            c.item += num_workers;
        } else if (worker < sk_heads) {
            if (c.item != 0) return false;
            c.item = 1;
            const int seg = worker / tiles;
            t = worker - seg * tiles;
            if (sk_left == 0) kb0 = (int)((long long)total_kb * seg / sk_f), kb1 = (int)((long long)total_kb * (seg + 1) / sk_f);
            else kb0 = seg * sk_H, kb1 = kb0 + sk_H;
        } else {
            if (c.pos >= sk_hi) return false;
            t = (int)(c.pos / tail_len);
            const int off = (int)(c.pos - (long long)t * tail_len);
            const long long left = sk_hi - c.pos;
            const int len = (int)(left < (long long)(tail_len - off) ? left : tail_len - off);
            kb0 = tail0 + off, kb1 = kb0 + len;
            c.pos += len;
        }
        n0 = (t % n_tiles) * (2 * BM), k0 = (t / n_tiles) * BN;
        return kb1 > kb0;
    };

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_dy);
        ptx::prefetch_tensormap(&map_x);
        ptx::prefetch_tensormap(&map_dw);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(tmem_full + b, 1);
            ptx::mbar_init(tmem_empty + b, 4 * 2);  // one arrive per epilogue warp of both CTAs
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc<2>(tmem_slot, 512);  // two 128 x 256 fp32 accumulators: the read-out of one segment runs under the MMAs of the next
        ptx::tmem_relinquish<2>();
    }
    ptx::tc_fence_before();
    ptx::cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            Cursor cur = first();
            int n0, k0, kb0, kb1;
            while (next(cur, n0, k0, kb0, kb1)) {
                const int na = n0 + cta_rank * BM, ka = k0 + cta_rank * (BN / 2);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * kStageBytes;
                    if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);
                    ptx::tma_load_3d_2sm(sa, &map_dy, &full_bar[stage], na, kb * BK, 0);
                    ptx::tma_load_3d_2sm(sa + kAtomBytes, &map_dy, &full_bar[stage], na + 64, kb * BK, 0);
                    ptx::tma_load_3d_2sm(sa + kABytes, &map_x, &full_bar[stage], ka, kb * BK, 0);
                    ptx::tma_load_3d_2sm(sa + kABytes + kAtomBytes, &map_x, &full_bar[stage], ka + 64, kb * BK, 0);
                    if (++stage == kStages) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(2 * BM, BN, 1, 1);  // both operands MN-major
            int stage = 0, it = 0;
            uint32_t phase = 0;
            Cursor cur = first();
            int n0, k0, kb0, kb1;
            for (; next(cur, n0, k0, kb0, kb1); ++it) {
                const int ab = it & 1;  // accumulator buffer
                const uint32_t acc_addr = tmem_base + ab * BN;
                ptx::mbar_wait(tmem_empty + ab, ((it >> 1) & 1) ^ 1);  // the epilogue has read this buffer's previous segment out
                ptx::tc_fence_after();
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * kStageBytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // 16 m-rows = two 8-row groups of 1024 B; the second 64-wide MN atom sits kAtomBytes further
                        const uint64_t da = ptx::umma_desc_mn_sw128(sa + k * 2048, kAtomBytes, 1024);
                        const uint64_t db = ptx::umma_desc_mn_sw128(sa + kABytes + k * 2048, kAtomBytes, 1024);
                        ptx::umma_bf16_ss<2>(acc_addr, da, db, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                    }
                    ptx::umma_commit_mcast2(&empty_bar[stage], 0x3);
                    if (++stage == kStages) stage = 0, phase ^= 1;
                }
                ptx::umma_commit_mcast2(tmem_full + ab, 0x3);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, et = threadIdx.x - 128;
        const uint32_t buf0 = ptx::smem_u32(epi_buf);
        int it = 0;
        Cursor cur = first();
        int n0, k0, kb0, kb1;
        for (; next(cur, n0, k0, kb0, kb1); ++it) {
            const int ab = it & 1;
            const int row_base = n0 + cta_rank * BM;
            const uint32_t taddr = tmem_base + ab * BN + (static_cast<uint32_t>(q * 32) << 16);
            ptx::mbar_wait(tmem_full + ab, (it >> 1) & 1);
            ptx::tc_fence_after();
            uint32_t acc[2][32];
            ptx::tmem_ld_32x32b_x32(taddr, acc[0]);
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
                ptx::tmem_ld_wait();
                if (c + 1 < BN / 32) {
                    ptx::tmem_ld_32x32b_x32(taddr + (c + 1) * 32, acc[(c + 1) & 1]);
                } else {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (cta_rank == 0) ptx::mbar_arrive(tmem_empty + ab);
                        else ptx::mbar_arrive_cluster(tmem_empty + ab, 0);
                    }
                }
                const uint32_t sb = buf0 + (c & 1) * kEpiBufBytes;
                if (et == 0) ptx::tma_store_wait_read<1>();
                asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t addr = sb + (uint32_t)(et * 128 + ((j ^ (et & 7)) << 4));
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(acc[c & 1][j * 4]), "r"(acc[c & 1][j * 4 + 1]),
                                 "r"(acc[c & 1][j * 4 + 2]), "r"(acc[c & 1][j * 4 + 3])
                                 : "memory");
                }
                ptx::fence_proxy_async();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (et == 0) {
                    ptx::tma_reduce_add_3d(&map_dw, epi_buf + (c & 1) * kEpiBufBytes, k0 + c * 32, row_base, 0);
                    ptx::tma_store_commit();
                }
            }
        }
        if (et == 0) ptx::tma_store_wait_all<0>();
    }

    ptx::tc_fence_before();
    ptx::cluster_sync_all();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<2>(tmem_base, 512);
    }
}

}  // namespace bsi

using namespace bsi;

extern "C" int bsi_gemm_wgrad_bf16(float* dW, const void* dY_bf16, const void* X_bf16, int32_t M, int32_t N, int32_t K, int32_t ldy, int32_t ldx,
                                   int32_t ldw, int32_t splits, void* stream) {
    using namespace wg;
    BSI_CHECK_ARG(dW && dY_bf16 && X_bf16 && M > 0 && N > 0 && K > 0, "bsi_gemm_wgrad_bf16: null pointer or empty problem");
    BSI_CHECK_ARG(N % 8 == 0 && K % 8 == 0 && ldy >= N && ldx >= K && ldw >= K && ldy % 8 == 0 && ldx % 8 == 0 && ldw % 4 == 0,
                  "bsi_gemm_wgrad_bf16: N, K and the pitches must be multiples of 8 (ldw of 4); got N=%d K=%d ldy=%d ldx=%d ldw=%d", N, K, ldy, ldx, ldw);
    BSI_ENSURE_SMEM(k_wgrad_bf16, kSmem);
    CUtensorMap my, mx, mw;
    int rc = make_tile_map(&my, dY_bf16, 2, M, N, ldy, 1, 0, BK);
    if (rc != BSI_OK) return rc;
    rc = make_tile_map(&mx, X_bf16, 2, M, K, ldx, 1, 0, BK);
    if (rc != BSI_OK) return rc;
    rc = make_tile_map(&mw, dW, 4, N, K, ldw, 1, 0, BM);
    if (rc != BSI_OK) return rc;
    const int n_tiles = (N + 2 * BM - 1) / (2 * BM), k_tiles = (K + BN - 1) / BN, total_kb = (M + BK - 1) / BK;
    const int pairs = sm_count() / 2;
    int kb_per_split = total_kb, workers;
    if (splits <= 0) {  // equal block-steps per CTA pair, heads + tails (kernel comment); at least 8 per pair
        splits = 0;
        const long long all_steps = (long long)n_tiles * k_tiles * total_kb;
        const long long w = all_steps / 8 > 0 ? all_steps / 8 : 1;
        workers = (int)(w < pairs ? w : pairs);
        kb_per_split = (int)(all_steps / workers);  // H
    } else {
        if (splits > total_kb) splits = total_kb;
        kb_per_split = (total_kb + splits - 1) / splits;
        splits = (total_kb + kb_per_split - 1) / kb_per_split;  // no empty splits
        const int items = n_tiles * k_tiles * splits;
        workers = items < pairs ? items : pairs;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(workers * 2), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = kSmem, cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    BSI_CUDA_OK(cudaLaunchKernelEx(&cfg, k_wgrad_bf16, my, mx, mw, n_tiles, k_tiles, splits, kb_per_split, total_kb));
    BSI_LAUNCH_OK("k_wgrad_bf16");
    return BSI_OK;
}
