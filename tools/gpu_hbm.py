#!/usr/bin/env python
"""Launch every HBM-bound kernel of the path once more at the imagenet64 sizes (bench.py's roofline_hbm cases): the ncu target for
profiles/ncu_full_r02_hbm_kernels.csv.  Prints the CUDA-event timings as JSON."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda", 0)
peak = bench.measured_peaks()[1]
if os.environ.get("HBM_ONE_LAUNCH"):  # ncu target: one warm-up and one timed launch per kernel
    bench._time_ms = (lambda orig: (lambda fn, warmup, iters, flush=None: orig(fn, 1, 1, flush)))(bench._time_ms)
print(json.dumps(bench.hbm_kernel_rooflines(dev, peak)), flush=True)
