#!/usr/bin/env python
"""Micro-benchmarks of the native kernels on one B200 (development aid; bench.py is the judged benchmark).

Prints one JSON line per measurement: GEMM TFLOP/s on the DiT-L shapes next to cuBLAS (torch.matmul) on the same
shape, attention / LayerNorm / step-kernel timings, and a full DiT-L/4 forward.
"""

import ctypes
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bsi_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda", 0)


def timeit(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def gemm_args(A, W, C, bias, epi, M, N, K, **kw):
    a = L.GemmArgs()
    a.A, a.W, a.C, a.bias = A.data_ptr(), W.data_ptr(), C.data_ptr(), bias.data_ptr() if bias is not None else None
    a.M, a.N, a.K, a.lda, a.ldw, a.ldc, a.batch = M, N, K, K, K, N, 1
    a.epilogue = epi
    a.gate = kw.get("gate", L.RowRef(None, 0, 0))
    a.rows_per_sample = kw.get("rows_per_sample", 256)
    return a


def main():
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    out = []
    M = int(os.environ.get("PROBE_M", 65536))
    shapes = [("qkv", 3072, 1024, L.EPI_BIAS_BF16), ("out", 1024, 1024, L.EPI_GATE_RESID_F32), ("mlp1", 4096, 1024, L.EPI_BIAS_GELU_BF16),
              ("mlp2", 1024, 4096, L.EPI_GATE_RESID_F32)]
    for name, N, K, epi in shapes:
        A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        W = (torch.randn(N, K, device=dev) / K**0.5).bfloat16()
        bias = torch.randn(N, device=dev) * 0.1
        f32 = epi >= L.EPI_BIAS_F32
        Cc = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
        gate = torch.randn(M // 256, N, device=dev)
        a = gemm_args(A, W, Cc, bias, epi, M, N, K, gate=L.rowref(gate, N, 0, 0))
        ms = timeit(lambda: L.check(lib.bsi_gemm_bf16(ctypes.byref(a), st)))
        lib.bsi_gemm_force_cta_group(1)
        ms1 = timeit(lambda: L.check(lib.bsi_gemm_bf16(ctypes.byref(a), st)))
        lib.bsi_gemm_force_cta_group(0)
        ms_cublas = timeit(lambda: torch.matmul(A, W.t()))
        fl = 2.0 * M * N * K
        out.append(dict(kernel=f"gemm_{name}", M=M, N=N, K=K, ms=ms, tflops=fl / ms / 1e9, single_cta_tflops=fl / ms1 / 1e9, cublas_ms=ms_cublas,
                        cublas_tflops=fl / ms_cublas / 1e9))
        print(json.dumps(out[-1]), flush=True)
        del A, W, Cc
    B = M // 256
    qkv = torch.randn(M, 3072, device=dev).bfloat16()
    o = torch.empty(M, 1024, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: L.check(lib.bsi_attention_bf16(o.data_ptr(), qkv.data_ptr(), B, 256, 16, 64, st)))
    lib.bsi_attention_force_legacy(1)
    ms_legacy = timeit(lambda: L.check(lib.bsi_attention_bf16(o.data_ptr(), qkv.data_ptr(), B, 256, 16, 64, st)))
    lib.bsi_attention_force_legacy(0)
    q, k, v = qkv.reshape(B, 256, 3, 16, 64).permute(2, 0, 3, 1, 4).contiguous()
    ms_sdpa = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
    fl = 4.0 * B * 16 * 256 * 256 * 64
    print(json.dumps(dict(kernel="attention", B=B, ms=ms, tflops=fl / ms / 1e9, mma_sync_ms=ms_legacy, sdpa_ms=ms_sdpa, sdpa_tflops=fl / ms_sdpa / 1e9)), flush=True)
    x = torch.randn(M, 1024, device=dev)
    tab = torch.randn(B, 6144, device=dev)
    xm = torch.empty(M, 1024, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: L.check(lib.bsi_layernorm_mod_bf16(xm.data_ptr(), x.data_ptr(), L.rowref(tab, 6144, 0, 0), L.rowref(tab, 6144, 0, 1024), None, None,
                                                           None, 256, M, 1024, 1e-5, st)))
    print(json.dumps(dict(kernel="layernorm_mod", M=M, ms=ms, gbs=M * 1024 * 6 / ms / 1e6)), flush=True)
    n, D = B, 12288
    mu, f = torch.randn(n, D, device=dev), torch.randn(n, D, device=dev)
    coef = torch.rand(4, 8, device=dev) + 0.5
    ms = timeit(lambda: L.check(lib.bsi_step_fused(mu.data_ptr(), f.data_ptr(), coef.data_ptr(), None, 1, 1, L.noise(seed=1, draw=1), None, None, n, D, st)), iters=50)
    print(json.dumps(dict(kernel="step_fused_philox", n=n, D=D, ms=ms, gbs=n * D * 12 / ms / 1e6)), flush=True)

    # full DiT-L/4 forward (random init, adaLN gates re-randomised so that every kernel does real work)
    from bsi_b200.models import DenoisingDiT
    from bsi_b200.nn import FourierFeatures

    depth = int(os.environ.get("PROBE_DEPTH", 24))
    torch.manual_seed(0)
    m = DenoisingDiT((3, 64, 64), 4, 1024, depth, 16, dropout=0.05, fourier_features=FourierFeatures(n_min=6, n_max=8))
    for blk in m.dit.blocks:
        torch.nn.init.normal_(blk.adaLN_modulation[2].weight, std=0.02)
        torch.nn.init.normal_(blk.adaLN_modulation[2].bias, std=0.02)
    m = m.to(dev).eval().requires_grad_(False)
    mu = torch.randn(B, 3, 64, 64, device=dev)
    t = torch.rand(B, device=dev)
    with torch.inference_mode():
        t0 = time.time()
        y = m(mu, t)
        torch.cuda.synchronize()
        first = time.time() - t0
        ms = timeit(lambda: m(mu, t), warmup=2, iters=5)
    fl = B * (161.61e9 * depth / 24)
    print(json.dumps(dict(kernel="dit_forward", B=B, depth=depth, ms=ms, tflops=fl / ms / 1e9, first_call_s=first, finite=bool(torch.isfinite(y).all()))), flush=True)


if __name__ == "__main__":
    main()
