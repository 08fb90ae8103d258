"""Optimizer-side probe at DiT-L size (478.9 M parameters): native two-launch step vs the reference's torch sequence
(clip_grad_norm_ + AdamW(fused=True) + _foreach_lerp_).  Prints one JSON line per arm with achieved HBM GB/s."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from bsi_b200 import optim as NO  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
# DiT-L/4-like parameter list: 24 blocks of (qkv, out, mlp1, mlp2, adaLN) + small tensors
shapes = []
for _ in range(24):
    shapes += [(3072, 1024), (3072,), (1024, 1024), (1024,), (4096, 1024), (4096,), (1024, 4096), (1024,), (6144, 1024), (6144,)]
shapes += [(1024, 336), (1024,), (256, 1024), (48, 1024), (48,), (1024, 256), (1024,), (1024, 1024), (1024,)]
numel = sum(torch.Size(s).numel() for s in shapes)


def make():
    return torch.nn.ParameterList([torch.nn.Parameter(torch.randn(s, device=dev) * 0.02) for s in shapes])


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


# native
ps = make()
holder = torch.nn.Module()
holder.ps = ps
ema = NO.EMA(holder, beta=0.9999, update_after_step=0, update_every=1, include_online_model=False)
opt = NO.AdamW(ps.parameters(), lr=1e-3, weight_decay=0.01, max_grad_norm=1.0, bf16_copy=True)
opt.attach_ema(ema)
gsrc = torch.randn(opt._g.numel, device=dev) * 1e-3


def native_step():
    opt._g.flat.copy_(gsrc)  # stands in for backward; not counted below
    opt.step()
    ema.update()


def copy_only():
    opt._g.flat.copy_(gsrc)


ms = timeit(native_step) - timeit(copy_only)
bytes_native = numel * (4 + 42)
print(json.dumps(dict(arm="native", params=numel, ms=ms, algorithmic_bytes=bytes_native, gbps=bytes_native / ms / 1e6, launches=3)), flush=True)
del opt, ema, ps, holder, gsrc
torch.cuda.empty_cache()

# reference sequence on the same GPU
ps = make()
ema_ps = [p.detach().clone() for p in ps]
ropt = torch.optim.AdamW(ps.parameters(), lr=1e-3, weight_decay=0.01, fused=True)
grads = [torch.randn_like(p) * 1e-3 for p in ps]


def set_grads():
    for p, g in zip(ps, grads):
        p.grad = g.clone()


def ref_step():
    set_grads()
    torch.nn.utils.clip_grad_norm_(ps.parameters(), 1.0)
    ropt.step()
    torch._foreach_lerp_(ema_ps, [p.data for p in ps], 1.0 - 0.9999)


ms_ref = timeit(ref_step) - timeit(set_grads)
print(json.dumps(dict(arm="torch clip+AdamW(fused)+foreach_lerp", params=numel, ms=ms_ref, speedup=ms_ref / ms)), flush=True)
