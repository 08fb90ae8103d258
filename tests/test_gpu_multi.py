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
    q.put((rank, [p.detach().cpu() for p in model.ps], [p.detach().cpu() for p in ema.ema_model.ps], norms))
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
    res = {r: (ps, es, ns) for r, ps, es, ns in (q.get(timeout=300) for _ in procs)}
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
