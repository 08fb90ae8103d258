#!/usr/bin/env python
"""Host-side cost of one training step (the step is host-bound at small per-GPU batches): cProfile of step() with the device kept busy."""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = [sys.argv[0]]
os.environ.setdefault("PROBE_B", "128")
src = open(os.path.join(ROOT, "tools", "gpu_train_timeline.py")).read().split("for _ in range(3):\n    step()\ntorch.cuda.synchronize()")[0]
exec(compile(src, "setup", "exec"))
for _ in range(3):
    step()
torch.cuda.synchronize()
import time

t0 = time.perf_counter()
step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host time to enqueue one step: {t_host * 1e3:.1f} ms; until the device is done: {t_all * 1e3:.1f} ms")
pr = cProfile.Profile()
pr.enable()
step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
