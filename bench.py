#!/usr/bin/env python
"""Headline benchmark: BSI.sample samples/sec on the imagenet64-dit configuration (BASELINE.json configs[3]).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU (oracle port)

One "step" = one full BSI.sample call: n = 256 samples of 3x64x64 per GPU through the k = 256 step sampler
(257 DiT-L/4 denoiser forwards + 256 fused posterior updates), random-init weights, synthetic noise.
Multi-GPU runs shard the samples over the ranks (weak scaling: 256 per GPU, global sample index keys the
noise, no data-path collective) and first run a correctness self-check of every N > 1 path ("multi_gpu_check").
Besides the headline the line carries, outside the timed region: "roofline_hbm" (the elementwise / reduction kernels
against the measured HBM peak), "elbo" (configs[3]: elbo(x[256], 1, 10), data points sharded) and "train_step"
(configs[4]: global batch 1024 data parallel, strong scaling).  Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs that fit one GPU; flops = algorithmic forward flops / sample / denoiser call
CONFIGS = {
    "imagenet64-dit": dict(kind="dit", shape=(3, 64, 64), patch=4, batch=256, depth=24, flops=161.26e9, label="DiT-L/4"),  # headline (configs[3])
    "imagenet32-dit": dict(kind="dit", shape=(3, 32, 32), patch=2, batch=512, depth=24, flops=161.11e9, label="DiT-L/2"),
    "cifar10-vdm": dict(kind="unet", shape=(3, 32, 32), patch=0, batch=256, depth=32, flops=53.47e9, label="VDM U-Net dim 128"),
}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=2)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="native", choices=["native", "reference"])
    # development overrides (the judged configuration is the default)
    p.add_argument("--config", default="imagenet64-dit", choices=sorted(CONFIGS))
    p.add_argument("--batch", type=int, default=None, help="samples per GPU (default: the configuration's batch)")
    p.add_argument("--k", type=int, default=256)
    p.add_argument("--depth", type=int, default=None, help="DiT depth / U-Net levels (default: the configuration's)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-side", action="store_true", help="skip the ELBO / training-step / HBM-kernel side measurements")
    a = p.parse_args()
    a.cfg = CONFIGS[a.config]
    a.batch = a.batch or a.cfg["batch"]
    a.depth = a.depth or a.cfg["depth"]
    return a


# ------------------------------------------------------------------------------------------ workload
def build_model(a):
    """Denoiser of the configuration (reference config/experiment/{imagenet64,imagenet32,cifar10-vdm}.yaml), random init as
    SURVEY §8(d): torch.manual_seed(0) construction; the DiT's adaLN output layers are re-randomised N(0, 0.02^2) with seed 1
    (adaLN-Zero would make every block the identity)."""
    from bsi_b200.models import DenoisingDiT, DenoisingVDMUNet, NyquistPositionalEmbedding
    from bsi_b200.nn import FourierFeatures

    torch.manual_seed(0)
    cfg = a.cfg
    if cfg["kind"] == "unet":
        m = DenoisingVDMUNet(cfg["shape"], NyquistPositionalEmbedding(32, 100), "silu", 128, a.depth, 4, n_attention_heads=1, dropout=0.1,
                             fourier_features=FourierFeatures(n_min=6, n_max=8))
        return m.eval().requires_grad_(False)
    m = DenoisingDiT(cfg["shape"], cfg["patch"], 1024, a.depth, 16, dropout=0.05, fourier_features=FourierFeatures(n_min=6, n_max=8))
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for blk in m.dit.blocks:
            blk.adaLN_modulation[2].weight.normal_(0.0, 0.02, generator=g)
            blk.adaLN_modulation[2].bias.normal_(0.0, 0.02, generator=g)
    return m.eval().requires_grad_(False)


def workload_name(a):
    c = a.cfg
    tag = "" if (a.batch, a.k, a.depth) == (c["batch"], 256, c["depth"]) else " [REDUCED development run]"
    shape = "x".join(str(v) for v in c["shape"])
    return f"{a.config} {c['label']} depth/levels {a.depth} (random init) BSI.sample k={a.k}, batch {a.batch} of {shape} per GPU{tag}"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )  # fmt: skip
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return d.get("bf16_tflops_sustained") or d.get("bf16_tflops"), d.get("hbm_gbs"), "measured (MEASURED_PEAKS.json, sustained cuBLAS bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"


# ------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_sample_rate(a, model_cpu_sd=None, batch=8, k_cpu=2):
    """Reference algorithm on the host cores: oracle.sample_with_noise with the oracle DiT, `k_cpu` steps at `batch`,
    extrapolated linearly to k steps (a full run is hours).  Returns (samples/s, seconds, cores)."""
    from oracle import bsi_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    shape = a.cfg["shape"]
    if a.cfg["kind"] == "unet":
        spec = O.UNetSpec(shape, dim=128, levels=a.depth)
        forward = lambda mu, tt: O.unet_forward(model_cpu_sd, spec, mu, tt)
    else:
        spec = O.DiTSpec(shape, a.cfg["patch"], 1024, a.depth, 16)
        forward = lambda mu, tt: O.dit_forward(model_cpu_sd, spec, mu, tt)
    if model_cpu_sd is None:
        model_cpu_sd = {k_: v.detach().float().cpu() for k_, v in build_model(a).state_dict().items()}
    consts = O.make_consts(1e-2, 1e6, 2e6)
    t = torch.linspace(0.0, 1.0, a.k + 1)[: k_cpu + 1].clone()
    eps = torch.randn((k_cpu + 1, batch, *shape), generator=torch.Generator().manual_seed(3))
    with torch.inference_mode():
        t0 = time.perf_counter()
        O.sample_with_noise(forward, consts, t, eps)
        dt = time.perf_counter() - t0
    per_sample_call = dt / (batch * (k_cpu + 1))
    return 1.0 / (per_sample_call * (a.k + 1)), dt, torch.get_num_threads(), f"k={k_cpu} steps ({k_cpu + 1} denoiser calls) at batch {batch}, fp32 eager PyTorch CPU, extrapolated linearly to k={a.k}"


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sd = {k_: v.detach().float().cpu() for k_, v in build_model(a).state_dict().items()}
    for _ in range(a.warmup):
        cpu_sample_rate(a, sd)
    t0 = time.perf_counter()
    rates = [cpu_sample_rate(a, sd) for _ in range(a.steps)]
    wall = time.perf_counter() - t0
    value = sum(r[0] for r in rates) / len(rates)
    line = {
        "impl": "reference", "metric": "BSI.sample samples/sec", "value": value, "unit": "samples/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * wall / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(a), "host": "CPU only; each step is a bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": rates[0][2], "kind": "port", "sample": rates[0][3]},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ native arm: side measurements
def _time_ms(fn, warmup: int, iters: int, flush=None) -> float:
    """Mean device time of fn() [ms] from CUDA events on the current stream; `flush` (a tensor larger than L2) is READ before every
    timed call so the kernel starts with none of its inputs in L2, as it does behind a denoiser forward.  (Round 2 first rewrote the
    buffer instead: that leaves 126 MB of dirty lines which the timed kernel has to write back -- 8-15 % of a 250 MB kernel's time.)"""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters


def hbm_kernel_rooflines(dev, peak_gbs: float, n: int = 256, n_measure: int = 10, shape=(3, 64, 64), patch: int = 4):
    """Achieved HBM bandwidth of the elementwise / reduction kernels of the path at the imagenet64 sizes: the sampler-side kernels
    on [n, D] (the per-step working set), the ELBO-side kernels on [n_measure * n, D].  bytes = ALGORITHMIC bytes per launch
    (x of the loss kernels is L2-resident across replicas and not counted).  L2 is flushed before every launch."""
    from bsi_b200 import _lib as L

    lib, st = L.load(), L.stream_ptr(dev)
    D = shape[0] * shape[1] * shape[2]
    R = n * n_measure
    g = torch.Generator(device=dev).manual_seed(11)
    flush = torch.zeros(96 * 1024 * 1024, device=dev)  # 384 MB > 126 MB L2
    mu_s, f_s = torch.randn(n, D, device=dev, generator=g), torch.randn(n, D, device=dev, generator=g)
    coef = torch.rand(4, 8, device=dev, generator=g) + 0.5
    x = (torch.randint(0, 256, (n, D), device=dev, generator=g).float() * (2 / 255) - 1).contiguous()
    # the reconstruction term's regime (bsi/bsi.py:224-228): mu = x + 1e-3 eps, x_hat = c_skip mu + c_out f with c_skip ~ 1, c_out = 1e-3
    mu = x.repeat(n_measure, 1) + 1e-3 * torch.randn(R, D, device=dev, generator=g)
    f = torch.randn(R, D, device=dev, generator=g)
    cs, co = torch.ones(R, device=dev), torch.full((R,), 1e-3, device=dev)
    out = torch.empty(R, device=dev)
    edges = torch.linspace(-1 - 1 / 255, 1 + 1 / 255, 257, device=dev)
    T = (shape[1] // patch) * (shape[2] // patch)
    cin = shape[0] * 7
    lda = (patch * patch * cin + 7) // 8 * 8
    A = torch.empty(n * T, lda, dtype=torch.bfloat16, device=dev)
    one = torch.ones(1, device=dev)
    xm = torch.empty(n * T, 1024, dtype=torch.bfloat16, device=dev)
    xs = torch.randn(n * T, 1024, device=dev, generator=g)
    mod = torch.rand(2, 1024, device=dev, generator=g) * 0.1  # one shift / scale row shared by all samples
    cases = [
        ("k_step_fused", n * D * 12, lambda: lib.bsi_step_fused(L.ptr(mu_s), L.ptr(f_s), L.ptr(coef), None, 1, 1, L.noise(seed=3, draw=1), None, None, n, D, st)),
        ("k_patch_operand_tiled", n * D * 4 + n * T * lda * 2, lambda: lib.bsi_dit_patch_operand(L.ptr(A), L.ptr(mu_s), L.rowref(one, 0), None, n, shape[0], shape[1], shape[2], patch, 6, 8, lda, st)),
        ("k_layernorm_mod", n * T * 1024 * 6, lambda: lib.bsi_layernorm_mod_bf16(L.ptr(xm), L.ptr(xs), L.rowref(mod, 0, 0, 0), L.rowref(mod, 0, 0, 1024), None, None, None, T, n * T, 1024, 1e-5, st)),
        ("k_sqerr_reduce", R * D * 8, lambda: lib.bsi_sqerr_reduce(L.ptr(out), L.ptr(x), L.ptr(mu), L.ptr(f), L.ptr(cs), L.ptr(co), R, n, D, st)),
        ("k_recon_reduce", R * D * 8, lambda: lib.bsi_recon_reduce(L.ptr(out), L.ptr(x), L.ptr(mu), L.ptr(f), L.ptr(cs), L.ptr(co), L.ptr(edges), 256, -1 - 1 / 255, 2 / 255,
                                                                   1414.2135, R, n, D, st)),
        ("k_q_sample", R * D * 8, lambda: lib.bsi_q_sample(L.ptr(mu), L.ptr(f), L.ptr(x), L.ptr(cs), L.ptr(co), L.ptr(cs), L.noise(seed=5), R, n, D, st)),
    ]
    rows = []
    for name, nbytes, fn in cases:
        ms = _time_ms(lambda: L.check(fn(), name), 2, 5, flush)
        rows.append({"kernel": name, "bytes": int(nbytes), "us": ms * 1e3, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak_gbs})
    return rows


def elbo_rate(a, bsi, dev, world: int, rank: int):
    """BASELINE.json configs[3], ELBO half: elbo(x[256], n_recon=1, n_measure=10); N > 1 shards the 256 data points (strong scaling,
    bsi_b200.distributed.sharded_elbo: one final all_gather of the [n, B] losses).  Returns (ms per call, bpd mean)."""
    import torch.distributed as dist

    from bsi_b200.distributed import sharded_elbo

    shape = a.cfg["shape"]
    B = a.batch
    x = torch.randint(0, 256, (B, *shape), device=dev, generator=torch.Generator(device=dev).manual_seed(2)).float() * (2 / 255) - 1
    with torch.inference_mode():
        for i in range(2):
            e, bpd, _ = sharded_elbo(bsi, x, 1, 10, 100 + i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 3
        for i in range(iters):
            e, bpd, _ = sharded_elbo(bsi, x, 1, 10, 200 + i)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, float(bpd.mean()), bool(torch.isfinite(bpd).all())


def train_step_rate(a, dev, world: int, rank: int, global_batch: int = 1024, micro: int = 256):
    """BASELINE.json configs[4]: imagenet64-dit train_loss forward/backward, bf16 tensor-core operands, data parallel with the global
    batch FIXED at 1024 (strong scaling -- the reference's loader divides the batch by the world size, bsi/data/h5image.py:309-312);
    a rank accumulates its share over micro-batches.  One step = train_loss(x).mean().backward() + one NCCL all-reduce over
    the flat gradient arena + clip/AdamW/EMA, dropout 0.05 as in config/experiment/imagenet64.yaml."""
    import torch.distributed as dist

    from bsi_b200 import BSI, Discretization
    from bsi_b200 import optim as NO
    from bsi_b200.models import DenoisingDiT
    from bsi_b200.nn import FourierFeatures

    shape = a.cfg["shape"]
    local = global_batch // world
    micro = min(micro, local)
    torch.manual_seed(0)  # identical replicas on every rank
    model = DenoisingDiT(shape, a.cfg["patch"], 1024, a.depth, 16, dropout=0.05, fourier_features=FourierFeatures(n_min=6, n_max=8)).to(dev).train()
    with torch.no_grad():
        for blk in model.dit.blocks:
            torch.nn.init.normal_(blk.adaLN_modulation[-1].weight, std=0.02)
            torch.nn.init.normal_(blk.adaLN_modulation[-1].bias, std=0.02)
    bsi = BSI(model, data_shape=shape, k=a.k, discretization=Discretization.image_8bit(), lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6, preconditioning="edm").to(dev)
    ema = NO.create_ema(model, beta=0.9999, update_after_step=1000, update_every=1)
    opt = NO.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01, max_grad_norm=1.0)
    opt.attach_ema(ema)
    opt.attach_model(model)
    gen = torch.Generator(device=dev).manual_seed(2 + rank)
    x = torch.randint(0, 256, (local, *shape), device=dev, generator=gen).float() * (2 / 255) - 1

    def step():
        opt.zero_grad()
        total = 0.0
        for i in range(0, local, micro):
            loss = bsi.train_loss(x[i : i + micro], gen).sum() * (world / global_batch)
            if i + micro < local:
                with opt.no_sync():
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
    iters = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    finite = bool(torch.isfinite(loss))
    del opt, ema, bsi, model
    torch.cuda.empty_cache()
    return ms, local, micro, finite


def multi_gpu_check(a, model, dev, world: int, rank: int):
    """Correctness of the N > 1 paths on the hardware the driver runs on (its GPU test box has one GPU), before anything is timed:
    (i)   sharded_sample + all_gather == rank 0's single-batch rows, bit for bit;
    (ii)  sharded_elbo   + all_gather == rank 0's single-GPU elbo, bit for bit (one lambda grid for the whole batch);
    (iii) one data-parallel optimisation step of a depth-2 DiT: the all-reduced gradient arena / world == the mean of the ranks'
          local gradients, and parameters + EMA weights are bit-identical on all ranks after step().
    Raises on any mismatch."""
    import torch.distributed as dist

    from bsi_b200 import BSI, Discretization
    from bsi_b200 import optim as NO
    from bsi_b200.distributed import sharded_elbo, sharded_sample
    from bsi_b200.models import DenoisingDiT
    from bsi_b200.nn import FourierFeatures

    shape = a.cfg["shape"]
    hyper = dict(lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6, preconditioning="edm", discretization=Discretization.image_8bit())
    res = {"world": world}
    bsi4 = BSI(model, data_shape=shape, k=4, **hyper).to(dev)
    n_total = 2 * world + 1  # ragged split
    with torch.inference_mode():
        full = sharded_sample(bsi4, n_total, seed=123)
        single = bsi4.sample(n_total, seed=123)
        ok = torch.tensor([int(torch.equal(full, single))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        res["sharded_sample_equals_single_gpu"] = bool(ok.item())
        x = torch.randint(0, 256, (n_total, *shape), device=dev, generator=torch.Generator(device=dev).manual_seed(9)).float() * (2 / 255) - 1
        e, bpd, ex = sharded_elbo(bsi4, x, 1, 2, 321)
        e1, bpd1, ex1 = bsi4.elbo(x, 1, 2, torch.Generator(device=dev).manual_seed(321))
        ok = torch.tensor([int(torch.equal(bpd, bpd1) and torch.equal(ex["l_measure"], ex1["l_measure"]))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        res["sharded_elbo_equals_single_gpu"] = bool(ok.item())
    del bsi4
    # (iii) DDP semantics (bsi/tasks/bsi.py:163-166): same replica everywhere, different data, same lambda grid / noise seed
    torch.manual_seed(0)
    if a.cfg["kind"] == "dit":
        small = DenoisingDiT(shape, a.cfg["patch"], 256, 2, 4, dropout=None, fourier_features=FourierFeatures(n_min=6, n_max=8)).to(dev).train()
        with torch.no_grad():
            for blk in small.dit.blocks:
                torch.nn.init.normal_(blk.adaLN_modulation[-1].weight, std=0.02)
        sb = BSI(small, data_shape=shape, k=8, **hyper).to(dev)
        ema = NO.create_ema(small, beta=0.9, update_after_step=0, update_every=1)
        opt = NO.AdamW(small.parameters(), lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
        opt.attach_ema(ema)
        opt.attach_model(small)
        xb = torch.randint(0, 256, (4, *shape), device=dev, generator=torch.Generator(device=dev).manual_seed(50 + rank)).float() * (2 / 255) - 1
        opt.zero_grad()
        with opt.no_sync():
            sb.train_loss(xb, torch.Generator(device=dev).manual_seed(7)).mean().backward()
        local = opt._g.flat.clone()
        opt.zero_grad()
        sb.train_loss(xb, torch.Generator(device=dev).manual_seed(7)).mean().backward()
        opt.all_reduce_grads()
        pieces = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(pieces, local)
        mean = torch.stack(pieces).double().mean(0)
        got = opt._g.flat.double() / world
        err = float((got - mean).abs().max() / mean.abs().max())
        res["ddp_grad_vs_mean_of_rank_grads_max_rel"] = err
        opt.step()
        ema.update()
        state = torch.cat((opt._p.flat, opt._ema_arena.flat))
        pieces = [torch.empty_like(state) for _ in range(world)]
        dist.all_gather(pieces, state)
        res["ranks_bit_identical_after_step"] = all(torch.equal(pieces[0], p_) for p_ in pieces[1:])
        res["grad_norm"] = float(opt.total_grad_norm())
        bad = err > 1e-5 or not res["ranks_bit_identical_after_step"]
        del opt, ema, sb, small
    else:
        bad = False
    torch.cuda.synchronize()
    if bad or not res["sharded_sample_equals_single_gpu"] or not res["sharded_elbo_equals_single_gpu"]:
        raise RuntimeError(f"multi-GPU self-check failed on rank {rank}: {res}")
    return res


# ------------------------------------------------------------------------------------------ native arm
def run_native(a):
    import torch.distributed as dist

    from bsi_b200 import BSI, Discretization
    from bsi_b200 import _lib as L
    from bsi_b200.distributed import gather_rows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.load()
    if lib.bsi_device_arch() != 100:
        raise RuntimeError(f"bsi_b200 kernels are built for sm_100a; device reports sm_{lib.bsi_device_arch()}")

    model = build_model(a).to(dev)
    shape = a.cfg["shape"]
    bsi = BSI(model, data_shape=shape, lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6, k=a.k, preconditioning="edm",
              discretization=Discretization.image_8bit()).to(dev)
    n, D = a.batch, shape[0] * shape[1] * shape[2]
    check = multi_gpu_check(a, model, dev, world, rank) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(i):
        return bsi.sample(n, seed=1000 + i, sample_offset=rank * n)

    with torch.inference_mode():
        for i in range(a.warmup):
            one_step(i)
        # ---- timed region: device-resident inputs --------------------------------------------------------------
        clocks = ClockSampler(local)
        barrier()
        c0 = lib.bsi_launch_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            out = one_step(a.warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        host_launches = lib.bsi_launch_counter() - c0
        clock_info = clocks.stop()
        # ---- end-to-end: schedule from pinned host memory in; final all_gather of the shards (N > 1); samples to pinned host memory out
        t_host = torch.linspace(0.0, 1.0, a.k + 1).pin_memory()
        out_host = torch.empty((n, *shape), dtype=torch.float32).pin_memory()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(a.steps):
            t_dev = t_host.to(dev, non_blocking=True)
            res = bsi.sample(n, t=t_dev, seed=2000 + i, sample_offset=rank * n)
            if world > 1:
                everything = gather_rows(res, world * n)  # the one collective of the sampling path (north_star: "a final gather")
            out_host.copy_(res, non_blocking=True)
        f1.record()
        barrier()
        ms_e2e = f0.elapsed_time(f1)
        finite = bool(torch.isfinite(out).all()) and bool(torch.isfinite(out_host).all())

        # ---- kernel accounting: launches of one eager forward; GEMM time from CUDA events around every GEMM launch ----
        k_prof = min(3, a.k)
        kk, lam, coef, c_in, t_rows = bsi._step_table(torch.linspace(0.0, 1.0, k_prof + 1, device=dev))
        c1 = lib.bsi_launch_counter()
        model(out, torch.ones(n, device=dev))
        # minus the conditioning launches of a stand-alone forward (time embedding + 2 GEMMs for the DiT, + 3 for the U-Net)
        fwd_launches = lib.bsi_launch_counter() - c1 - (4 if a.cfg["kind"] == "unet" else 3)
        torch.cuda.synchronize()
        L.check(lib.bsi_profile_gemm_begin())
        model.sample_loop(n, torch.rsqrt(lam[:1]).contiguous(), coef, c_in, t_rows, kk, 7, rank * n, 1, use_graph=False)
        g_ms, g_fl, g_n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int32()
        L.check(lib.bsi_profile_gemm_end(ctypes.byref(g_ms), ctypes.byref(g_fl), ctypes.byref(g_n)))

    peak_tf, peak_hbm, peak_src = measured_peaks()
    full_size = (a.batch, a.k, a.depth) == (a.cfg["batch"], 256, a.cfg["depth"])
    side = {}
    if a.cfg["kind"] == "dit" and not a.no_side:
        hbm_rows = hbm_kernel_rooflines(dev, peak_hbm, n=a.batch, shape=shape, patch=a.cfg["patch"]) if rank == 0 else None
        el_ms, el_bpd, el_ok = elbo_rate(a, bsi, dev, world, rank)
        model._scratch.clear()  # conditioning tables / workspaces of the evaluation model: make room for the training step
        bsi._plans.clear()
        torch.cuda.empty_cache()
        gb = 1024 if full_size else 8 * world
        tr_ms, tr_local, tr_micro, tr_ok = train_step_rate(a, dev, world, rank, global_batch=gb, micro=256 if full_size else 4)
        side = dict(hbm=hbm_rows, el_ms=el_ms, el_bpd=el_bpd, el_ok=el_ok, tr_ms=tr_ms, tr_local=tr_local, tr_micro=tr_micro, tr_ok=tr_ok, gb=gb)

    times = torch.tensor([ms, ms_e2e, side.get("el_ms", 0.0), side.get("tr_ms", 0.0)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, ms_e2e, el_ms, tr_ms = (float(v) for v in times)
    if rank == 0:
        value = world * n * a.steps / (ms / 1e3)
        e2e = world * n * a.steps / (ms_e2e / 1e3)
        gemm_tf = g_fl.value / g_ms.value / 1e9 if g_ms.value > 0 else 0.0
        step_flops = n * (a.k + 1) * a.cfg["flops"] * a.depth / a.cfg["depth"]
        # kernels executed per sample() call: init + 3 conditioning + k x (forward + step + advance) + final forward + combine
        # (the step graph is captured once per (n, k) and reused, so there is no per-call warm-up forward any more)
        per_call = 1 + (4 if a.cfg["kind"] == "unet" else 3) + a.k * (fwd_launches + 2) + fwd_launches + 1
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get(a.config, {}).get("k_gemm_bf16_dram_bytes_per_launch")
        line = {
            "metric": "BSI.sample samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": workload_name(a), "parallelism": f"sample-sharded x{world}, no data-path collective (e2e adds the final all_gather)",
                "l2": "working set per step (bf16 weights + GBs of activations per denoiser forward) exceeds the 126 MB L2; no explicit flush",
                "precision": "bf16 tensor-core operands, fp32 accumulation, fp32 belief state / residual stream / losses",
                "whole_step_tflops_per_gpu": step_flops / (ms / a.steps) / 1e9, "outputs_finite": finite,
            },
            "roofline": {
                "bound": "tensor", "kernel": "k_gemm_bf16 (tcgen05; implicit-GEMM convolutions for the U-Net)", "achieved": gemm_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": gemm_tf / peak_tf if peak_tf else None, "traffic": traffic, "peak_source": peak_src,
                "how": f"CUDA events around each of {g_n.value} GEMM launches of a {k_prof}-step eager sampler pass at the benchmark batch (sum flops / sum time); "
                       "traffic = mean dram bytes per GEMM launch from the committed ncu --set full capture (profiles/ncu_traffic.json)",
                "gemm_share_of_step": (g_ms.value / (k_prof + 1)) * (a.k + 1) / (ms / a.steps),
            },
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": (a.k + 1) * 4, "d2h_bytes_per_step": n * D * 4},
            "gpu_launches": int(per_call * a.steps), "host_enqueued_launches": int(host_launches), "clocks": clock_info,
            "multi_gpu_check": check,
        }  # fmt: skip
        if side:
            B = a.batch
            el_flops = (1 + 10) * B * (a.cfg["flops"] + 0.352e9) * a.depth / a.cfg["depth"]
            tr_flops = side["gb"] * 3 * (a.cfg["flops"] + 0.352e9) * a.depth / a.cfg["depth"]
            line["roofline_hbm"] = {"peak_gbs": peak_hbm, "l2": "evicted before every launch (a 384 MB buffer is read: clean lines, no write-backs inside the timed kernel)", "kernels": side["hbm"]}
            line["elbo"] = {
                "workload": f"{a.config} elbo(x[{B}], n_recon=1, n_measure=10), data points sharded x{world} (strong scaling, final all_gather)",
                "value": B / el_ms * 1e3, "unit": "data points/s", "ms_per_call": el_ms, "tflops_per_gpu": el_flops / el_ms / 1e9 / world,
                "bpd_mean": side["el_bpd"], "finite": side["el_ok"],
            }
            line["train_step"] = {
                "workload": f"{a.config} train_loss fwd/bwd + gradient all-reduce + clip/AdamW/EMA, global batch {side['gb']} over {world} GPU(s) "
                            f"({side['tr_local']} per GPU in micro-batches of {side['tr_micro']}), dropout 0.05" + ("" if full_size else " [REDUCED development run]"),
                "ms_per_step": tr_ms, "samples_per_s": side["gb"] / tr_ms * 1e3, "tflops_per_gpu": tr_flops / tr_ms / 1e9 / world, "scaling": "strong",
                "all_reduce": "one NCCL all-reduce over the flat gradient arena after the backward (per-block collectives during the backward are 2.6 ms slower on 8 GPUs)" if world > 1 else "none (1 GPU)",
                "finite": side["tr_ok"],
            }
        if world == 1 and not a.no_cpu_baseline:
            v, dt, cores, sample = cpu_sample_rate(a)
            line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample, "seconds": dt}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (sm_100a): the bsi_b200 path has no CPU fallback; use --impl reference for the CPU arm")
        run_native(args)
