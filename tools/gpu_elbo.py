#!/usr/bin/env python
"""Timing of BSI.elbo(x[256], n_recon=1, n_measure=10) on the imagenet64-dit configuration (BASELINE.json configs[3], ELBO half)
and of the loss-reduction kernels alone (HBM roofline).  Development aid; prints JSON lines."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bsi_b200 import BSI, Discretization
from bsi_b200 import _lib as L

dev = torch.device("cuda", 0)
a = type("A", (), {})()
a.config = "imagenet64-dit"; a.cfg = bench.CONFIGS[a.config]; a.depth = int(os.environ.get("PROBE_DEPTH", 24)); a.batch = 256; a.k = 256
model = bench.build_model(a).to(dev)
bsi = BSI(model, data_shape=(3, 64, 64), lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6, k=256, preconditioning="edm", discretization=Discretization.image_8bit()).to(dev)
x = (torch.randint(0, 256, (256, 3, 64, 64), device=dev, generator=torch.Generator(device=dev).manual_seed(2)).float() * (2 / 255) - 1)

def timeit(fn, warmup=2, iters=3):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters

with torch.inference_mode():
    gen = torch.Generator(device=dev).manual_seed(0)
    ms = timeit(lambda: bsi.elbo(x, 1, 10, gen))
    e, b, ex = bsi.elbo(x, 1, 10, gen)
    fl = 2816 * 161.61e9 * a.depth / 24
    print(json.dumps(dict(what="elbo(x[256],1,10) imagenet64-dit", ms=ms, data_points_per_s=256 / ms * 1e3, tflops=fl / ms / 1e9, bpd_mean=float(b.mean()), finite=bool(torch.isfinite(b).all()))), flush=True)
    # loss reductions alone at the measure-term size [10*256, 12288]
    R, B, D = 2560, 256, 12288
    lib, st = L.load(), torch.cuda.current_stream().cuda_stream
    mu, f = torch.randn(R, D, device=dev), torch.randn(R, D, device=dev)
    cs, co = torch.rand(R, device=dev), torch.rand(R, device=dev)
    out = torch.empty(R, device=dev)
    xf = x.reshape(B, D).contiguous()
    ms_sq = timeit(lambda: L.check(lib.bsi_sqerr_reduce(out.data_ptr(), xf.data_ptr(), mu.data_ptr(), f.data_ptr(), cs.data_ptr(), co.data_ptr(), R, B, D, st)), iters=20)
    disc = Discretization.image_8bit(); edges = disc.bin_boundaries(dev, torch.float32)
    ms_rc = timeit(lambda: L.check(lib.bsi_recon_reduce(out.data_ptr(), xf.data_ptr(), mu.data_ptr(), f.data_ptr(), cs.data_ptr(), co.data_ptr(), edges.data_ptr(), 256,
                                                        disc.range[0], disc.dx, 1414.2135, R, B, D, st)), iters=20)
    gamma, sigma = torch.rand(R, device=dev), torch.rand(R, device=dev)
    ms_q = timeit(lambda: L.check(lib.bsi_q_sample(mu.data_ptr(), f.data_ptr(), xf.data_ptr(), gamma.data_ptr(), sigma.data_ptr(), cs.data_ptr(), L.noise(seed=1), R, B, D, st)), iters=20)
    by = R * D * 8
    print(json.dumps(dict(what="loss kernels at [2560,12288]", sqerr_ms=ms_sq, sqerr_gbs=by / ms_sq / 1e6, recon_ms=ms_rc, recon_gbs=by / ms_rc / 1e6, q_sample_ms=ms_q,
                          q_sample_gbs=by / ms_q / 1e6)), flush=True)
