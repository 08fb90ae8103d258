"""Config-5 probe: one optimisation step of DiT-L/4 on imagenet64 shapes through the native training path --
BSI.train_loss(x).mean().backward() + fused clip/AdamW/EMA -- timed with CUDA events.  Prints one JSON line.

    python tools/gpu_train.py [batch] [depth]
"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from bsi_b200 import BSI, Discretization  # noqa: E402
from bsi_b200 import optim as NO  # noqa: E402
from bsi_b200.models import DenoisingDiT  # noqa: E402
from bsi_b200.nn import FourierFeatures  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 24
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = DenoisingDiT((3, 64, 64), 4, 1024, depth, 16, dropout=None, fourier_features=FourierFeatures(n_min=6, n_max=8)).to(dev).train()
with torch.no_grad():
    for blk in model.dit.blocks:  # adaLN-Zero would make every block the identity
        torch.nn.init.normal_(blk.adaLN_modulation[-1].weight, std=0.02)
        torch.nn.init.normal_(blk.adaLN_modulation[-1].bias, std=0.02)
bsi = BSI(model, data_shape=(3, 64, 64), k=256, discretization=Discretization.image_8bit(), lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6,
          preconditioning="edm").to(dev)
ema = NO.create_ema(model, beta=0.9999, update_after_step=1000, update_every=1)
opt = NO.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
opt.attach_ema(ema)
gen = torch.Generator(device=dev).manual_seed(2)
x = torch.randint(0, 256, (B, 3, 64, 64), device=dev, generator=gen).float() * (2 / 255) - 1


def step():
    opt.zero_grad()
    loss = bsi.train_loss(x, gen).mean()
    loss.backward()
    opt.step()
    ema.update()
    return loss


for _ in range(2):
    loss = step()
torch.cuda.synchronize()
torch.cuda.reset_peak_memory_stats()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3
t0 = time.perf_counter()
a.record()
for _ in range(n):
    loss = step()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
flops = B * 3 * (161.26e9 * depth / 24 + 0.352e9)  # forward + dgrad + wgrad per sample
print(json.dumps(dict(what="imagenet64-dit train step (native path)", batch=B, depth=depth, ms=ms, samples_per_s=B / ms * 1e3, tflops=flops / ms / 1e9,
                      wall_ms=(time.perf_counter() - t0) / n * 1e3, peak_mem_gb=torch.cuda.max_memory_allocated() / 2**30, loss=float(loss),
                      params=sum(p.numel() for p in model.parameters()))))

if len(sys.argv) > 3 and sys.argv[3] == "profile":
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    rows = sorted(((e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"),
                  key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    for name, ms_k, cnt in rows[:32]:
        print(f"{ms_k:8.2f} ms {100 * ms_k / tot:5.1f}% n={cnt:5d}  {name[:110]}")
    print(f"total device time {tot:.1f} ms")
