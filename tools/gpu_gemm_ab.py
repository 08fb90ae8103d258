#!/usr/bin/env python
"""GEMM A/B at the four DiT-L shapes (M = 65536) for the library selected by BSI_B200_LIB: back-to-back launches for ~1 s per shape
(sustained, power-capped clocks, like inside a sampler step).  One JSON line."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bsi_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda", 0)
lib, st = L.load(), torch.cuda.current_stream().cuda_stream
M = 65536
shapes = [("qkv", 3072, 1024, L.EPI_BIAS_BF16), ("out", 1024, 1024, L.EPI_GATE_RESID_F32), ("mlp1", 4096, 1024, L.EPI_BIAS_GELU_BF16), ("mlp2", 1024, 4096, L.EPI_GATE_RESID_F32)]
res = {"lib": os.path.basename(L.LIB_PATH)}
for name, N, K, epi in shapes:
    if os.environ.get("PROBE_ONLY") and name not in os.environ["PROBE_ONLY"].split(","):
        continue
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    W = (torch.randn(N, K, device=dev) / K**0.5).bfloat16()
    bias = torch.randn(N, device=dev) * 0.1
    f32 = epi >= L.EPI_BIAS_F32
    Cc = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    gate = torch.randn(M // 256, N, device=dev) * 0.01
    a = L.GemmArgs()
    a.A, a.W, a.C, a.bias = A.data_ptr(), W.data_ptr(), Cc.data_ptr(), bias.data_ptr()
    a.M, a.N, a.K, a.lda, a.ldw, a.ldc, a.batch = M, N, K, K, K, N, 1
    a.epilogue, a.gate, a.rows_per_sample = epi, L.rowref(gate, N, 0, 0), 256
    fn = lambda: L.check(lib.bsi_gemm_bf16(ctypes.byref(a), st))
    for _ in range(200):  # ~80 ms: settle the clocks under load
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 1500
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    res[name] = {"us": round(ms * 1e3, 1), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}
    del A, W, Cc
print(json.dumps(res), flush=True)
