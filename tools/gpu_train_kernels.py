#!/usr/bin/env python
"""The HBM-bound kernels of the training step in isolation (M = 32768 token rows of 1024: batch 128): achieved bandwidth with the
L2 flushed before every launch, on fixed buffers (``--rotate N`` walks through N different buffer sets spread over tens of GB, like
the activations of a real step).  One JSON line per kernel."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bsi_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rotate", type=int, default=1)
ap.add_argument("--batch", type=int, default=128)
args = ap.parse_args()
dev = torch.device("cuda", 0)
lib, st = L.load(), torch.cuda.current_stream().cuda_stream
B, T, D = args.batch, 256, 1024
M = B * T
torch.manual_seed(0)
flush = torch.zeros(96 << 20, dtype=torch.float32, device=dev)  # 384 MB, READ before every launch: the L2 is left full of clean lines
ap_dirty = os.environ.get("FLUSH_DIRTY", "0") == "1"  # write instead: the L2 is left full of dirty lines that the timed kernel has to evict
mods = torch.randn(B, 6 * D, device=dev) * 0.1
none = L.RowRef(None, 0, 0)
sets = []
for _ in range(args.rotate):
    sets.append(dict(x=torch.randn(M, D, device=dev), br=torch.randn(M, D, device=dev).bfloat16(), xo=torch.empty(M, D, device=dev),
                     a=torch.empty(M, D, device=dev, dtype=torch.bfloat16), dx=torch.randn(M, D, device=dev), da=torch.randn(M, D, device=dev).bfloat16(),
                     parts=torch.empty(2, M // 32, D, device=dev), dbr=torch.empty(M, D, device=dev, dtype=torch.bfloat16),
                     dgate=torch.empty(B, D, device=dev), dbias=torch.empty(B, D, device=dev)))
ref = lambda j: L.rowref(mods, 6 * D, 0, j * D)


def gate_ln(s, drop=0.0):
    L.check(lib.bsi_gate_residual_layernorm_bf16(s["a"].data_ptr(), s["xo"].data_ptr(), s["x"].data_ptr(), s["br"].data_ptr(), ref(2), ref(3), ref(4), None, None,
                                                 T, M, D, 1e-5, drop, 1234, st))


def ln_bwd(s, drop=0.0):
    L.check(lib.bsi_layernorm_mod_backward(s["dx"].data_ptr(), s["parts"][0].data_ptr(), s["parts"][1].data_ptr(), s["da"].data_ptr(), s["x"].data_ptr(), ref(4), None,
                                           T, 32, M, D, 1e-5, drop, 1234, st))


def gate_bwd(s):
    L.check(lib.bsi_gate_residual_backward(s["dbr"].data_ptr(), s["dgate"].data_ptr(), s["dbias"].data_ptr(), s["dx"].data_ptr(), s["br"].data_ptr(), ref(5), T, B, D, st))


def gate_bwd_rows(s):
    L.check(lib.bsi_gate_residual_backward_rows(s["dbr"].data_ptr(), s["parts"][0].data_ptr(), s["parts"][1].data_ptr(), s["dx"].data_ptr(), s["br"].data_ptr(), ref(5),
                                                T, 32, M, D, st))


def ln_fwd(s):
    L.check(lib.bsi_layernorm_mod_bf16(s["a"].data_ptr(), s["x"].data_ptr(), ref(0), ref(1), None, None, None, T, M, D, 1e-5, st))


cases = [
    ("k_gate_residual_layernorm", lambda s: gate_ln(s), M * D * (4 + 2 + 4 + 2)),
    ("k_gate_residual_layernorm dropout", lambda s: gate_ln(s, 0.05), M * D * (4 + 2 + 4 + 2)),
    ("k_layernorm_mod_backward", lambda s: ln_bwd(s), M * D * (4 + 2 + 4 + 4)),
    ("k_layernorm_mod_backward dropout", lambda s: ln_bwd(s, 0.05), M * D * (4 + 2 + 4 + 4)),
    ("k_gate_residual_backward", gate_bwd, M * D * (4 + 2 + 2)),
    ("k_gate_residual_backward_pipe", gate_bwd_rows, M * D * (4 + 2 + 2)),
    ("k_layernorm_mod (inference kernel, same rows)", ln_fwd, M * D * (4 + 2)),
]
only = os.environ.get("CASES")  # comma-separated substrings: run only the matching cases (ncu captures)
n_iter = int(os.environ.get("ITERS", 12))
for name, fn, nbytes in cases:
    if only and not any(o in name for o in only.split(",")):
        continue
    for i in range(1 if only else 3):
        fn(sets[i % len(sets)])
    torch.cuda.synchronize()
    tot, n = 0.0, n_iter
    for i in range(n):
        flush.add_(1.0) if ap_dirty else flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(sets[i % len(sets)])
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    us = tot / n * 1e3
    print(json.dumps({"kernel": name, "rows": M, "rotate": args.rotate, "flush": "dirty" if ap_dirty else "clean", "us": round(us, 1), "algorithmic_MB": round(nbytes / 1e6, 1), "GBps": round(nbytes / us / 1e3, 0),
                      "frac_of_6543": round(nbytes / us / 1e3 / 6543, 3)}), flush=True)
