"""Shared helpers for the -m gpu tests (CUDA path vs oracle / torch fp32)."""

import os

import torch

import helpers as H
from bsi_b200 import _lib as L

O = H.O
OUT_DIR = os.path.join(H.ROOT, "gpurun_out")


def dev():
    return torch.device("cuda", 0)


def report(name: str, got: torch.Tensor, ref: torch.Tensor, rtol: float, atol: float):
    """assert_close with a diagnostic that localises the failure (rows/cols pattern) for kernel debugging."""
    got32, ref32 = got.detach().float().cpu(), ref.detach().float().cpu()
    diff = (got32 - ref32).abs()
    tol = atol + rtol * ref32.abs()
    bad = diff > tol
    if not bool(bad.any()) and bool(torch.isfinite(got32).all()):
        return
    msg = [f"{name}: {int(bad.sum())}/{bad.numel()} mismatched, max abs err {float(diff.max()):.4g}, ref max {float(ref32.abs().max()):.4g}"]
    msg.append(f"non-finite in got: {int((~torch.isfinite(got32)).sum())}")
    if bad.ndim == 2:
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        msg.append(f"bad rows: {rows.numel()} (first {rows[:16].tolist()}), bad cols: {cols.numel()} (first {cols[:16].tolist()})")
    idx = bad.nonzero()[:5]
    for i in idx:
        t = tuple(i.tolist())
        msg.append(f"  at {t}: got {float(got32[t]):.6g} ref {float(ref32[t]):.6g}")
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, "test_diagnostics.txt"), "a") as fh:
        fh.write("\n".join(msg) + "\n")
    raise AssertionError("\n".join(msg))


def call(fn_name: str, *args):
    lib = L.load()
    L.check(getattr(lib, fn_name)(*args), fn_name)


def sync():
    torch.cuda.synchronize()
