"""GroupNorm+SiLU probe at the cifar10-vdm shape: [256 images][1024 px][128 ch] fp32 -> bf16."""
import json
import sys

import torch

sys.path.insert(0, ".")
from bsi_b200 import _lib as L  # noqa: E402

lib = L.load()
dev = torch.device("cuda:0")
B, HW, C = 256, 1024, 128
xs = [torch.randn(B, HW, C, device=dev) for _ in range(3)]  # rotate inputs: 134 MB each, beyond L2
act = torch.empty(B, HW, C, device=dev, dtype=torch.bfloat16)
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
st = torch.cuda.current_stream().cuda_stream


def run(i):
    L.check(lib.bsi_groupnorm_act_bf16(act.data_ptr(), None, xs[i % 3].data_ptr(), gamma.data_ptr(), beta.data_ptr(), B, HW, C, 4, 1e-5, 1, st))


for i in range(6):
    run(i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(30):
    run(i)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 30
ref = torch.nn.functional.silu(torch.nn.functional.group_norm(xs[2].permute(0, 2, 1), 32, gamma, beta, 1e-5)).permute(0, 2, 1)
run(2)
err = float((act.float() - ref).abs().max())
print(json.dumps(dict(kernel="groupnorm_silu", ms=ms, gbps=B * HW * C * 6 / ms / 1e6, max_err=err)))
