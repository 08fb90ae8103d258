"""Two-GPU data-parallel optimisation step (needs 2 CUDA devices; skipped on a single-GPU box): one NCCL sum all-reduce over
the flat gradient arena + 1/world folded into the fused AdamW kernel must equal DDP's averaged gradients fed to the oracle."""

import os
import socket

import pytest
import torch

import helpers as H
from gpu_util import report
from oracle import optim_oracle as OO

pytestmark = pytest.mark.gpu
STEPS = 3


def _rank_grad(i, step, rank):
    return H.optim_grad(i, 2 * step + rank)


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from bsi_b200 import optim as NO

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)

    class Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(H.det_uniform(f"optim.p{i}", shp)) for i, shp in enumerate(H.OPTIM_SHAPES)])

    model = Holder().to(dev)
    ema = NO.EMA(model, include_online_model=False, **H.OPTIM_EMA)
    opt = NO.AdamW(model.parameters(), max_grad_norm=H.OPTIM_MAX_NORM, **H.OPTIM_HYPER)
    opt.attach_ema(ema)
    norms = []
    for step in range(STEPS):
        opt.zero_grad()
        for i, p in enumerate(model.ps):
            p.grad.add_(_rank_grad(i, step, rank).to(dev))
        opt.all_reduce_grads()
        opt.step()
        ema.update()
        norms.append(float(opt.total_grad_norm()))
    torch.cuda.synchronize()
    q.put((rank, [p.detach().cpu().numpy() for p in model.ps], [p.detach().cpu().numpy() for p in ema.ema_model.ps], norms))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_step_equals_ddp_average():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = {r: ([torch.from_numpy(p) for p in ps], [torch.from_numpy(e) for e in es], ns) for r, ps, es, ns in (q.get(timeout=300) for _ in procs)}
    [p.join(60) for p in procs]
    # ranks stay bit-identical replicas
    for a, b in zip(res[0][0] + res[0][1], res[1][0] + res[1][1]):
        assert torch.equal(a, b)
    assert res[0][2] == res[1][2]
    # and equal the single-process reference on the rank-averaged gradients (what DistributedDataParallel hands the optimizer)
    params = [H.det_uniform(f"optim.p{i}", shp) for i, shp in enumerate(H.OPTIM_SHAPES)]
    side = OO.OptimizerSide(params, max_norm=H.OPTIM_MAX_NORM, ema=OO.EMASchedule(**H.OPTIM_EMA), **H.OPTIM_HYPER)
    for step in range(STEPS):
        side.step([(_rank_grad(i, step, 0) + _rank_grad(i, step, 1)) * 0.5 for i in range(len(params))])
    for i in range(len(params)):
        report(f"ddp param {i}", res[0][0][i], side.params[i], 2e-6, 1e-8)
        report(f"ddp ema {i}", res[0][1][i], side.ema[i], 2e-6, 1e-8)


def _train_worker(rank, world, port, q, mode="allreduce"):
    """One data-parallel optimisation step of the native DiT: each rank differentiates train_loss on its half of the batch."""
    import torch.distributed as dist

    from bsi_b200 import BSI, Discretization
    from bsi_b200 import optim as NO
    from bsi_b200.models import DenoisingDiT

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    spec = H.O.DiTSpec((3, 32, 32), 2, 128, 1, 2, fourier=None)
    model = DenoisingDiT(spec.data_shape, spec.patch, spec.dim, spec.depth, spec.heads, dropout=None, fourier_features=None)
    model.load_state_dict(H.det_state_dict(H.dit_shapes(spec), seed=1))
    model = model.to(dev).train()
    bsi = BSI(model, data_shape=spec.data_shape, k=16, discretization=Discretization.image_8bit(), lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6,
              preconditioning="edm").to(dev)
    bsi.noise_source = "torch"
    opt = NO.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
    if mode == "sink":
        opt.attach_model(model, overlap=True)  # gradients straight into the arena, per-block all-reduce started during the backward
    net = model
    if mode == "ddp" and world > 1:  # the reference's own arrangement (BSITraining.configure_ddp, bsi/tasks/bsi.py:163-166)
        from torch.nn.parallel import DistributedDataParallel

        net = DistributedDataParallel(model, device_ids=[rank], static_graph=True)
    x = H.det_images("mt.x", 8, spec.data_shape, seed=2).to(dev)
    # both layouts see the same per-sample lambdas and noise: the 8-sample single-process run and the two 4-sample halves
    lam = H.det_uniform("mt.lam", (1, 8)).to(dev).abs() * 5 + 0.05
    eps = H.det_uniform("mt.eps", (8, *spec.data_shape)).to(dev)
    lo, hi = (0, 8) if world == 1 else (4 * rank, 4 * rank + 4)

    def local_loss():
        # train_loss with injected lambda / noise (bsi/bsi.py:291-310): 0.5 * reciprocal_pdf(lambda) * ||x - x_hat||^2
        lam_l = lam[:, lo:hi]
        t = bsi.p_lambda.cdf(lam_l).flatten()
        c_skip, c_out, c_in = bsi._edm_preconditioning(t)
        mu = x[lo:hi] * ((lam_l - bsi.lambda_0) / lam_l).reshape(-1, 1, 1, 1) + torch.rsqrt(lam_l).reshape(-1, 1, 1, 1) * eps[lo:hi]
        f = net(mu * c_in.reshape(-1, 1, 1, 1), t)
        x_hat = c_skip.reshape(-1, 1, 1, 1) * mu + c_out.reshape(-1, 1, 1, 1) * f
        err = (x[lo:hi] - x_hat).square().flatten(1).sum(1)
        return (0.5 * bsi.p_lambda.reciprocal_pdf(lam_l).flatten() * err).sum() / 8  # mean over the GLOBAL batch

    opt.zero_grad()
    (local_loss() * world).backward()  # all_reduce_grads averages over ranks, so each rank contributes world * (its share of the mean)
    if world > 1 and mode != "ddp":
        opt.all_reduce_grads()  # (DistributedDataParallel has already averaged the gradients in its hooks)
    grads = [(p.grad * opt._grad_scale).cpu().numpy() for p in model.parameters()]  # what the optimizer kernel is about to consume
    opt.step()
    torch.cuda.synchronize()
    # numpy arrays are pickled by value (torch tensors travel as shared-memory handles that die with this process)
    q.put((world, rank, [p.detach().cpu().numpy() for p in model.parameters()], grads, float(opt.total_grad_norm())))
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("mode", ["allreduce", "sink", "ddp"], ids=["allreduce_after_backward", "overlapped_per_block", "torch_ddp_wrapper"])
def test_native_dit_data_parallel_step_equals_single_process(mode):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    out = {}
    for world in (1, 2):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        procs = [ctx.Process(target=_train_worker, args=(r, world, port, q, mode)) for r in range(world)]
        [p.start() for p in procs]
        for _ in procs:
            w, r, params, grads, norm = q.get(timeout=300)
            out[(w, r)] = ([torch.from_numpy(p) for p in params], [torch.from_numpy(g) for g in grads], norm)
        [p.join(60) for p in procs]
    (_, g1, n_single), (p0, g0, n0), (p1, _, n1) = out[(1, 0)], out[(2, 0)], out[(2, 1)]
    assert n0 == n1 and abs(n0 - n_single) <= 2e-3 * n_single
    for b, c in zip(p0, p1):
        assert torch.equal(b, c)  # replicas stay bit-identical through the step
    # averaged gradients of the two halves == gradients of the whole batch (Adam's first step divides by |g|, so the
    # parameters themselves are ill-conditioned where g is ~0: the gradients are the meaningful comparison)
    # Tolerance 1e-2 relative L2: the adaLN chain runs under bf16 autocast, so its weight gradients are rounded to bf16 per rank
    # before the all-reduce (as in the reference's bf16 mixed-precision training); the native wgrad GEMMs accumulate in fp32.
    for a, b in zip(g1, g0):
        assert float((a - b).norm()) <= 1e-2 * float(a.norm()) + 1e-9
