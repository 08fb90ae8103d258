#!/usr/bin/env python
"""Timing of the native VDM U-Net forward (cifar10-vdm configuration) and its conv GEMM on one B200 (development aid)."""
import ctypes, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bsi_b200 import _lib as L
from bsi_b200.models import DenoisingVDMUNet, NyquistPositionalEmbedding
from bsi_b200.nn import FourierFeatures

dev = torch.device("cuda", 0)

def timeit(fn, warmup=2, iters=5):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters

B = int(os.environ.get("PROBE_B", 256)); levels = int(os.environ.get("PROBE_LEVELS", 32))
torch.manual_seed(0)
m = DenoisingVDMUNet((3, 32, 32), NyquistPositionalEmbedding(32, 100), "silu", 128, levels, 4, n_attention_heads=1, dropout=0.1,
                     fourier_features=FourierFeatures(n_min=6, n_max=8)).to(dev).eval().requires_grad_(False)
mu, t = torch.randn(B, 3, 32, 32, device=dev), torch.rand(B, device=dev)
lib = L.load()
with torch.inference_mode():
    y = m(mu, t); torch.cuda.synchronize()
    c0 = lib.bsi_launch_counter(); m(mu, t); launches = lib.bsi_launch_counter() - c0
    ms = timeit(lambda: m(mu, t))
    L.check(lib.bsi_profile_gemm_begin()); m(mu, t)
    g_ms, g_fl, g_n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int32()
    L.check(lib.bsi_profile_gemm_end(ctypes.byref(g_ms), ctypes.byref(g_fl), ctypes.byref(g_n)))
fl = B * 53.47e9 * (levels / 32)
print(json.dumps(dict(kernel="unet_forward", B=B, levels=levels, ms=ms, tflops=fl / ms / 1e9, launches=launches, gemm_ms=g_ms.value, gemm_launches=g_n.value,
                      gemm_tflops=g_fl.value / g_ms.value / 1e9, finite=bool(torch.isfinite(y).all()))), flush=True)
