#!/usr/bin/env python
"""Attention backward A/B at the training shape (B = 128, T = 256, 16 heads of 64): tcgen05 kernel vs the mma.sync row-owner kernels
(BSI_ATT_BWD_VARIANT=0), with and without dropout.  One JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bsi_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda", 0)
lib, st = L.load(), torch.cuda.current_stream().cuda_stream
B = int(os.environ.get("PROBE_B", 128))
M = B * 256
torch.manual_seed(0)
qkv = torch.randn(M, 3072, device=dev).bfloat16()
dout = torch.randn(M, 1024, device=dev).bfloat16()
out = torch.empty(M, 1024, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B * 16 * 256, device=dev)
dsum = torch.empty_like(lse)
dqkv = torch.empty_like(qkv)
L.check(lib.bsi_attention_lse_bf16(out.data_ptr(), lse.data_ptr(), qkv.data_ptr(), B, 256, 16, 64, st))


def timeit(fn, warmup=2, iters=10):
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


res = {"variant": os.environ.get("BSI_ATT_BWD_VARIANT", "1"), "B": B}
for p in (0.0, 0.05):
    ms = timeit(lambda: L.check(lib.bsi_attention_backward_bf16(dqkv.data_ptr(), lse.data_ptr(), dsum.data_ptr(), qkv.data_ptr(), out.data_ptr(), dout.data_ptr(),
                                                                 B, 256, 16, 64, p, 1234, 1, st)))
    fl = 10.0 * B * 16 * 256 * 256 * 64  # five 2*T*T*64 products per (sample, head)
    res[f"p{p}"] = {"us": round(ms * 1e3, 1), "tflops": round(fl / ms / 1e9, 1), "finite": bool(torch.isfinite(dqkv.float()).all())}
if os.environ.get("BSI_ATT_BWD_VARIANT") == "9":
    import ctypes

    buf = (ctypes.c_ulonglong * 16)()
    lib.bsi_attention_backward_debug_phases(buf)
    L.check(lib.bsi_attention_backward_bf16(dqkv.data_ptr(), lse.data_ptr(), dsum.data_ptr(), qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), B, 256, 16, 64, 0.0, 0, 1, st))
    torch.cuda.synchronize()
    lib.bsi_attention_backward_debug_phases(buf)
    n = max(1, buf[15])
    res["softmax_clk_per_item"] = dict(zip(["wait_S_dP", "math", "wait_acc", "readout"], [round(buf[i] / n) for i in range(4)]))
    res["control_clk_per_item"] = dict(zip(["wait_inputs", "wait_P", "wait_readout", "rest"], [round(buf[8 + i] / n) for i in range(4)]))
print(json.dumps(res), flush=True)
