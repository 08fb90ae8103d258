#!/usr/bin/env python
"""Attention kernel A/B: time bsi_attention_bf16 at the DiT-L shape (B = 256, T = 256, 16 heads of 64) for the variant selected by
BSI_ATT_VARIANT, check it against torch SDPA, and time SDPA next to it.  One JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bsi_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda", 0)
lib, st = L.load(), torch.cuda.current_stream().cuda_stream
B = int(os.environ.get("PROBE_B", 256))
M = B * 256
torch.manual_seed(0)
qkv = (torch.randn(M, 3072, device=dev) * float(os.environ.get("PROBE_SCALE", 1.0))).bfloat16()
o = torch.empty(M, 1024, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B * 16 * 256, device=dev)


def timeit(fn, warmup=3, iters=20):
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


ms = timeit(lambda: L.check(lib.bsi_attention_bf16(o.data_ptr(), qkv.data_ptr(), B, 256, 16, 64, st)))
ms_lse = timeit(lambda: L.check(lib.bsi_attention_lse_bf16(o.data_ptr(), lse.data_ptr(), qkv.data_ptr(), B, 256, 16, 64, st)))
q, k, v = qkv.reshape(B, 256, 3, 16, 64).permute(2, 0, 3, 1, 4).contiguous()
ms_sdpa = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
ref = torch.nn.functional.scaled_dot_product_attention(q[:8].float(), k[:8].float(), v[:8].float()).permute(0, 2, 1, 3).reshape(8 * 256, 1024)
L.check(lib.bsi_attention_lse_bf16(o.data_ptr(), lse.data_ptr(), qkv.data_ptr(), B, 256, 16, 64, st))
torch.cuda.synchronize()
err = float((o[: 8 * 256].float() - ref).abs().max())
rel = float((o[: 8 * 256].float() - ref).norm() / ref.norm())
lse_ref = torch.logsumexp(q[:8].float() @ k[:8].float().transpose(-1, -2) / 8.0, dim=-1) * 1.4426950408889634
lse_err = float((lse.reshape(B, 16, 256)[:8] - lse_ref).abs().max())
fl = 4.0 * B * 16 * 256 * 256 * 64
if os.environ.get("BSI_ATT_VARIANT") in ("9", "5"):
    import ctypes

    buf = (ctypes.c_ulonglong * 14)()
    lib.bsi_attention_debug_phases(buf)
    L.check(lib.bsi_attention_bf16(o.data_ptr(), qkv.data_ptr(), B, 256, 16, 64, st))
    torch.cuda.synchronize()
    lib.bsi_attention_debug_phases(buf)
    items = max(1, buf[5])
    print(json.dumps(dict(phases_clk_per_item=dict(zip(["wait_S", "pass1", "pass2", "wait_O", "epilogue"], [round(buf[i] / items, 1) for i in range(5)])),
                          control_clk_per_item=dict(zip(["wait_qk_ofree", "S_latency", "wait_P0", "wait_P1", "PV_tail"], [round(buf[8 + i] / items, 1) for i in range(5)])), items=items)))
print(json.dumps(dict(variant=os.environ.get("BSI_ATT_VARIANT", "0"), B=B, us=ms * 1e3, us_lse=ms_lse * 1e3, tflops=fl / ms / 1e9, sdpa_us=ms_sdpa * 1e3, max_abs_err=err,
                      rel_l2=rel, lse_max_err=lse_err)), flush=True)
