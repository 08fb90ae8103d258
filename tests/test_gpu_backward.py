"""Backward building blocks for config 5 (not wired into an engine yet): the weight-gradient GEMM on tcgen05
with MN-major operands and split-M TMA reduce-add, and the data-gradient GEMM through the forward kernel with transposed
weights -- each against torch fp32 matmuls of the same bf16-rounded operands."""

import pytest
import torch

from gpu_util import call, dev, report, sync
from bsi_b200 import _lib as L
from test_gpu_gemm import gemm, rnd

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K,splits", [(4096, 256, 256, 0), (4096, 256, 256, 1), (1000, 384, 320, 3), (8192, 1024, 1024, 0), (320, 136, 72, 0)])
def test_wgrad_vs_torch(M, N, K, splits):
    dy = rnd(f"wg.dy{M}", (M, N)).bfloat16()
    x = rnd(f"wg.x{M}", (M, K)).bfloat16()
    base = rnd(f"wg.acc{M}", (N, K))
    dw = base.clone()
    call("bsi_gemm_wgrad_bf16", L.ptr(dw), L.ptr(dy), L.ptr(x), M, N, K, N, K, K, splits, L.stream_ptr())
    sync()
    ref = base + dy.float().t() @ x.float()
    scale = float(ref.abs().max())
    report(f"wgrad {M}x{N}x{K} splits={splits}", dw, ref, 1e-3, 1e-3 * scale)
    # accumulating twice adds the same product again (autograd's +=)
    call("bsi_gemm_wgrad_bf16", L.ptr(dw), L.ptr(dy), L.ptr(x), M, N, K, N, K, K, splits, L.stream_ptr())
    sync()
    report("wgrad second accumulation", dw, base + 2 * (dy.float().t() @ x.float()), 1e-3, 2e-3 * scale)


def test_wgrad_strided_operands_and_bad_arguments():
    M, N, K = 2048, 256, 192
    big_dy = rnd("wg.sdy", (M, 3 * N)).bfloat16()  # dY is a column slice of a packed [M, 3N] gradient (the QKV layout)
    x = rnd("wg.sx", (M, K)).bfloat16()
    dw = torch.zeros((N, K + 8), device=dev())
    dy = big_dy[:, N : 2 * N]
    call("bsi_gemm_wgrad_bf16", L.ptr(dw), dy.data_ptr(), L.ptr(x), M, N, K, 3 * N, K, K + 8, 0, L.stream_ptr())
    sync()
    report("wgrad strided", dw[:, :K], dy.float().t() @ x.float(), 1e-3, 1e-2)
    assert float(dw[:, K:].abs().max()) == 0.0
    lib = L.load()
    assert lib.bsi_gemm_wgrad_bf16(L.ptr(dw), dy.data_ptr(), L.ptr(x), M, N, 100, 3 * N, K, K + 8, 0, L.stream_ptr()) != 0
    assert b"multiples of 8" in lib.bsi_last_error()


def test_dgrad_through_forward_kernel_with_transposed_weights():
    """dX = dY W: the forward kernel computes A W'^T, so feeding W' = W^T (a bf16 [K][N] copy refreshed once per optimizer
    step) gives the data gradient with no new kernel."""
    M, N, K = 4096, 384, 256
    dy = rnd("dg.dy", (M, N)).bfloat16()
    w = rnd("dg.w", (N, K), 0.05)
    wt = w.t().contiguous().bfloat16()  # [K][N]
    dx = torch.zeros((M, K), device=dev())
    zero = torch.zeros(K, device=dev())
    gemm(dy, wt, dx, zero, L.EPI_BIAS_F32)
    report("dgrad", dx, dy.float() @ w.bfloat16().float(), 1e-3, 1e-3)
