#!/usr/bin/env python
"""Kernel timeline of the captured sampler step (imagenet64-dit, batch 256): per-kernel device time and the idle gaps between
consecutive kernels inside CUDA-graph replays, from CUPTI activity records (torch.profiler; nsys is not in the image).

    python tools/gpu_timeline.py [--k 4] [--batch 256] [--config imagenet64-dit] > gpurun_out/timeline.txt

Answers the question the per-kernel ncu list cannot: how much of a step is NOT inside any kernel (launch gaps, prologues that
cannot overlap the previous kernel's tail), and after which kernels the gaps sit.  Run once with BSI_PDL=0 and once with BSI_PDL=1.
"""
import argparse
import collections
import json
import os
import re
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bsi_b200 import BSI, Discretization  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=4)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--config", default="imagenet64-dit")
ap.add_argument("--depth", type=int, default=None)
args = ap.parse_args()

dev = torch.device("cuda", 0)
a = type("A", (), {})()
a.config, a.cfg = args.config, bench.CONFIGS[args.config]
a.depth, a.batch, a.k = args.depth or a.cfg["depth"], args.batch, args.k
model = bench.build_model(a).to(dev)
bsi = BSI(model, data_shape=a.cfg["shape"], lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6, k=a.k, preconditioning="edm", discretization=Discretization.image_8bit()).to(dev)

from torch.profiler import ProfilerActivity, profile  # noqa: E402

with torch.inference_mode():
    bsi.sample(a.batch, seed=1)  # capture
    bsi.sample(a.batch, seed=2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    bsi.sample(a.batch, seed=3)
    e1.record()
    torch.cuda.synchronize()
    wall_ms = e0.elapsed_time(e1)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        bsi.sample(a.batch, seed=4)
        torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
short = lambda n: re.sub(r"\(.*", "", n).replace("void ", "").replace("bsi::", "")[:48]
busy = sum(e["dur"] for e in ev)
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
per = collections.OrderedDict()
gap_after = collections.defaultdict(lambda: [0, 0.0])
for i, e in enumerate(ev):
    p = per.setdefault(short(e["name"]), [0, 0.0])
    p[0] += 1
    p[1] += e["dur"]
    if i + 1 < len(ev):
        g = ev[i + 1]["ts"] - (e["ts"] + e["dur"])
        ga = gap_after[(short(e["name"]), short(ev[i + 1]["name"]))]
        ga[0] += 1
        ga[1] += g
print(f"PDL={os.environ.get('BSI_PDL', '0')} config={a.config} batch={a.batch} k={a.k} depth={a.depth}: sample() {wall_ms:.2f} ms (CUDA events, unprofiled); "
      f"profiled span {span / 1e3:.2f} ms, kernels busy {busy / 1e3:.2f} ms, idle between kernels {(span - busy) / 1e3:.2f} ms = {100 * (span - busy) / span:.1f} %, "
      f"{len(ev)} kernels")
print("\nper kernel (us):")
for name, (n, d) in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print(f"  {name:48s} n={n:5d} total={d / 1e3:9.3f} ms share={100 * d / span:5.1f}% avg={d / n:8.1f}")
print("\nidle gap after kernel -> next kernel (us):")
for (a_, b_), (n, g) in sorted(gap_after.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"  {a_:40s} -> {b_:40s} n={n:5d} total={g / 1e3:8.3f} ms avg={g / n:7.2f}")
