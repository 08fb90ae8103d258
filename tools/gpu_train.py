"""Config-5 probe: optimisation steps of DiT-L/4 on imagenet64 shapes through the native training path --
BSI.train_loss(x).mean().backward() [+ one NCCL all-reduce over the flat gradient arena] + fused clip/AdamW/EMA -- timed with
CUDA events, max over ranks.  The global batch is fixed (strong scaling, like the reference's data loader which divides the batch
by the world size, bsi/data/h5image.py:309-312); a rank whose share exceeds --micro accumulates gradients over micro-batches.

    python tools/gpu_train.py [--global-batch 1024] [--micro 256] [--depth 24] [--steps 3] [--profile]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/gpu_train.py ...
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from bsi_b200 import BSI, Discretization  # noqa: E402
from bsi_b200 import optim as NO  # noqa: E402
from bsi_b200.models import DenoisingDiT  # noqa: E402
from bsi_b200.nn import FourierFeatures  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--global-batch", type=int, default=128)
ap.add_argument("--micro", type=int, default=256)
ap.add_argument("--depth", type=int, default=24)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--no-sink", action="store_true", help="gradients through autograd, one all-reduce after backward")
ap.add_argument("--overlap", action="store_true", help="start each block's all-reduce during the backward (default: one all-reduce after it)")
ap.add_argument("--dropout", type=float, default=0.0, help="0.05 in config/experiment/imagenet64.yaml")
a = ap.parse_args()

world, rank, local_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local_rank)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
local = a.global_batch // world
micro = min(a.micro, local)
assert local % micro == 0 and a.global_batch % world == 0

torch.manual_seed(0)  # identical replicas on every rank
model = DenoisingDiT((3, 64, 64), 4, 1024, a.depth, 16, dropout=a.dropout or None, fourier_features=FourierFeatures(n_min=6, n_max=8)).to(dev).train()
with torch.no_grad():
    for blk in model.dit.blocks:  # adaLN-Zero would make every block the identity
        torch.nn.init.normal_(blk.adaLN_modulation[-1].weight, std=0.02)
        torch.nn.init.normal_(blk.adaLN_modulation[-1].bias, std=0.02)
bsi = BSI(model, data_shape=(3, 64, 64), k=256, discretization=Discretization.image_8bit(), lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6,
          preconditioning="edm").to(dev)
ema = NO.create_ema(model, beta=0.9999, update_after_step=1000, update_every=1)
opt = NO.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
opt.attach_ema(ema)
if not a.no_sink:
    opt.attach_model(model, overlap=a.overlap)  # weight gradients straight into the arena; --overlap: per-block all-reduce during the backward
gen = torch.Generator(device=dev).manual_seed(2 + rank)
x = torch.randint(0, 256, (local, 3, 64, 64), device=dev, generator=gen).float() * (2 / 255) - 1


def step():
    opt.zero_grad()
    total = 0.0
    for i in range(0, local, micro):
        # mean over the global batch = sum of micro-batch sums / global batch (after the all-reduce's 1/world)
        loss = bsi.train_loss(x[i : i + micro], gen).sum() * (world / a.global_batch)
        if i + micro < local:
            with opt.no_sync():  # gradient accumulation: only the last micro-batch starts the all-reduces
                loss.backward()
        else:
            loss.backward()
        total += loss.detach()
    if world > 1:
        opt.all_reduce_grads()
    opt.step()
    ema.update()
    return total


for _ in range(2):
    loss = step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.reset_peak_memory_stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms = float(ms)
flops = a.global_batch * 3 * (161.26e9 * a.depth / 24 + 0.352e9)  # forward + dgrad + wgrad per sample
if rank == 0:
    print(json.dumps(dict(what="imagenet64-dit train step (native path)", no_sink=a.no_sink, overlap=a.overlap, n_gpus=world, global_batch=a.global_batch, per_gpu_batch=local, micro_batch=micro,
                          depth=a.depth, dropout=a.dropout, ms_per_step=ms, samples_per_s=a.global_batch / ms * 1e3, tflops_total=flops / ms / 1e9,
                          tflops_per_gpu=flops / ms / 1e9 / world, peak_mem_gb=torch.cuda.max_memory_allocated() / 2**30, loss=float(loss),
                          grad_norm=float(opt.total_grad_norm()), params=sum(p.numel() for p in model.parameters()))), flush=True)

if a.profile and rank == 0:
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
if world > 1:
    dist.destroy_process_group()
