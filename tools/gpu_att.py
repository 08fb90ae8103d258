"""Attention-only probe: times the T=256 attention kernels (tcgen05, mma.sync) against SDPA."""
import json
import sys

import torch

sys.path.insert(0, ".")
from bsi_b200 import _lib as L  # noqa: E402

lib = L.load()
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
st = torch.cuda.current_stream().cuda_stream
qkv = torch.randn(B * 256, 3072, device=dev).bfloat16()
o = torch.empty(B * 256, 1024, device=dev, dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


res = {}
for mode, name in ((0, "tcgen05"), (1, "mma_sync")):
    lib.bsi_attention_force_legacy(mode)
    res[name + "_ms"] = timeit(lambda: L.check(lib.bsi_attention_bf16(o.data_ptr(), qkv.data_ptr(), B, 256, 16, 64, st)))
lib.bsi_attention_force_legacy(0)
q, k, v = qkv.reshape(B, 256, 3, 16, 64).permute(2, 0, 3, 1, 4).contiguous()
res["sdpa_ms"] = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
print(json.dumps(res))
