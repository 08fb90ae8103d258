"""Weight-gradient GEMM probe at the DiT-L shapes (M = 65536 tokens, or PROBE_M) vs cuBLAS (torch.matmul of the transposed operand):
stream-K work distribution (splits = 0) against the round-robin (tile, split) items of round 1."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from bsi_b200 import _lib as L  # noqa: E402

lib = L.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
M = int(os.environ.get("PROBE_M", 65536))


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for name, N, K in (("qkv", 3072, 1024), ("out", 1024, 1024), ("mlp1", 4096, 1024), ("mlp2", 1024, 4096)):
    dy = torch.randn(M, N, device=dev).bfloat16()
    x = torch.randn(M, K, device=dev).bfloat16()
    dw = torch.zeros(N, K, device=dev)
    ms = timeit(lambda: L.check(lib.bsi_gemm_wgrad_bf16(dw.data_ptr(), dy.data_ptr(), x.data_ptr(), M, N, K, N, K, K, 0, st)))
    tiles = ((N + 255) // 256) * ((K + 255) // 256)
    splits_r1 = max(1, min(-(-2 * 74 // tiles), (M // 64) // 8))  # round 1's rule: two waves of the 74 CTA pairs
    ms_items = timeit(lambda: L.check(lib.bsi_gemm_wgrad_bf16(dw.data_ptr(), dy.data_ptr(), x.data_ptr(), M, N, K, N, K, K, splits_r1, st)))
    ms_cublas = timeit(lambda: torch.matmul(dy.t(), x))
    fl = 2.0 * M * N * K
    print(json.dumps(dict(kernel=f"wgrad_{name}", M=M, N=N, K=K, ms=ms, tflops=fl / ms / 1e9, items_splits=splits_r1, items_ms=ms_items, items_tflops=fl / ms_items / 1e9, cublas_ms=ms_cublas, cublas_tflops=fl / ms_cublas / 1e9)), flush=True)
    del dy, x, dw
