#!/usr/bin/env python
"""Idle time inside one training step (imagenet64-dit, batch 128, one GPU): CUPTI kernel records of a step driven from Python --
how much of the step is the GPU waiting for the host?  Prints the span, the busy time and the largest gaps with the kernels around them."""
import collections
import json
import os
import re
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bsi_b200 import BSI, Discretization  # noqa: E402
from bsi_b200 import optim as NO  # noqa: E402
from bsi_b200.models import DenoisingDiT  # noqa: E402
from bsi_b200.nn import FourierFeatures  # noqa: E402

dev = torch.device("cuda", 0)
B = int(os.environ.get("PROBE_B", 128))
torch.manual_seed(0)
model = DenoisingDiT((3, 64, 64), 4, 1024, 24, 16, dropout=0.05, fourier_features=FourierFeatures(n_min=6, n_max=8)).to(dev).train()
with torch.no_grad():
    for blk in model.dit.blocks:
        torch.nn.init.normal_(blk.adaLN_modulation[-1].weight, std=0.02)
bsi = BSI(model, data_shape=(3, 64, 64), k=256, discretization=Discretization.image_8bit(), lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6, preconditioning="edm").to(dev)
ema = NO.create_ema(model, beta=0.9999, update_after_step=1000, update_every=1)
opt = NO.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01, max_grad_norm=1.0)
opt.attach_ema(ema)
opt.attach_model(model)
gen = torch.Generator(device=dev).manual_seed(2)
x = torch.randint(0, 256, (B, 3, 64, 64), device=dev, generator=gen).float() * (2 / 255) - 1


def step():
    opt.zero_grad()
    bsi.train_loss(x, gen).mean().backward()
    opt.step()
    ema.update()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
short = lambda n: re.sub(r"\(.*", "", n).replace("void ", "").replace("bsi::", "")[:44]
busy = sum(e["dur"] for e in ev)
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
gaps = []
for a, b in zip(ev, ev[1:]):
    g = b["ts"] - (a["ts"] + a["dur"])
    if g > 0:
        gaps.append((g, short(a["name"]), short(b["name"])))
print(f"train step batch {B}: {ms:.2f} ms unprofiled; profiled span {span / 1e3:.2f} ms, device busy {busy / 1e3:.2f} ms, idle {(span - busy) / 1e3:.2f} ms "
      f"({100 * (span - busy) / span:.1f} %), {len(ev)} device activities, {sum(1 for g in gaps if g[0] > 20)} gaps > 20 us")
by = collections.defaultdict(lambda: [0, 0.0])
for g, a, b in gaps:
    by[(a, b)][0] += 1
    by[(a, b)][1] += g
print("largest idle totals by (kernel before -> kernel after):")
for (a, b), (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {a:44s} -> {b:44s} n={n:4d} total={t / 1e3:7.3f} ms avg={t / n:7.1f} us")
per = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    per[short(e["name"])][0] += 1
    per[short(e["name"])][1] += e["dur"]
print("device time per kernel:")
for name, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"  {name:44s} n={n:4d} total={t / 1e3:8.3f} ms avg={t / n:8.1f} us")
