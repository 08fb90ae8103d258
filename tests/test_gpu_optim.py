"""Optimizer side of the training step (clip -> AdamW -> EMA) on the native kernels vs the reference fixture and the oracle.

Tolerance: the CUDA kernel mirrors torch's single-tensor AdamW op by op with fp32 rounding after each op, so individual
steps agree with the CPU reference to an ulp or two; the global gradient norm is summed in a different order (one flat
arena instead of a norm of per-tensor norms), which perturbs the clip coefficient in the last bit.  rtol 2e-6 / atol 1e-8
over 14 steps covers both; bin-exact comparisons are used where the operation is a single rounding (EMA lerp, bf16 copy)."""

import ctypes

import pytest
import torch

import helpers as H
from gpu_util import call, dev, report, sync
from bsi_b200 import _lib as L
from bsi_b200 import optim as NO
from oracle import optim_oracle as OO

pytestmark = pytest.mark.gpu
RTOL, ATOL = 2e-6, 1e-8


class Holder(torch.nn.Module):
    def __init__(self, shapes, tag="optim.p"):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(H.det_uniform(f"{tag}{i}", shp)) for i, shp in enumerate(shapes)])


def run_fixture(fused_ema: bool, stray_grads: bool):
    model = Holder(H.OPTIM_SHAPES).to(dev())
    ema = NO.EMA(model, include_online_model=False, **H.OPTIM_EMA)
    opt = NO.AdamW(model.parameters(), max_grad_norm=H.OPTIM_MAX_NORM, fused=True, **H.OPTIM_HYPER)
    if fused_ema:
        opt.attach_ema(ema)
    hist = {"params": [], "ema": [], "decay": [], "norm": []}
    for step in range(H.OPTIM_STEPS):
        opt.zero_grad()
        for i, p in enumerate(model.ps):
            g = H.optim_grad(i, step).to(dev())
            if stray_grads:
                p.grad = g  # replaces the arena view, as zero_grad(set_to_none=True) + backward would
            else:
                p.grad.add_(g)  # accumulates like autograd does
        opt.step()
        ema.update()
        hist["decay"].append(ema.get_current_decay())
        hist["norm"].append(opt.total_grad_norm().cpu())
        hist["params"].append([p.detach().cpu().clone() for p in model.ps])
        hist["ema"].append([p.detach().cpu().clone() for p in ema.ema_model.ps])
    return hist, opt, model


def test_fixture_matches_reference_classes():
    g = H.load_golden("optim.pt")
    hist, opt, model = run_fixture(fused_ema=True, stray_grads=False)
    assert hist["decay"] == g["decay"]
    for step in range(H.OPTIM_STEPS):
        coef = torch.clamp(H.OPTIM_MAX_NORM / (hist["norm"][step] + 1e-6), max=1.0)
        report(f"clip coefficient step {step}", coef, g["coef"][step].reshape(1), 1e-6, 0)
        for i in range(len(H.OPTIM_SHAPES)):
            report(f"param {i} step {step}", hist["params"][step][i], g["params"][step][i], RTOL, ATOL)
            report(f"ema {i} step {step}", hist["ema"][step][i], g["ema"][step][i], RTOL, ATOL)
    for i, p in enumerate(model.ps):
        report(f"exp_avg {i}", opt.state[p]["exp_avg"], g["exp_avg"][i], RTOL, ATOL)
        report(f"exp_avg_sq {i}", opt.state[p]["exp_avg_sq"], g["exp_avg_sq"][i], RTOL, 1e-12)
        assert float(opt.state[p]["step"]) == H.OPTIM_STEPS
        assert float(p.grad.abs().max()) == 0.0  # the step leaves the gradient arena zeroed


def test_standalone_ema_and_replaced_grads_are_identical_to_fused_path():
    a, _, _ = run_fixture(fused_ema=True, stray_grads=False)
    b, _, _ = run_fixture(fused_ema=False, stray_grads=True)
    for step in range(H.OPTIM_STEPS):
        for i in range(len(H.OPTIM_SHAPES)):
            assert torch.equal(a["params"][step][i], b["params"][step][i]), (step, i)
            assert torch.equal(a["ema"][step][i], b["ema"][step][i]), (step, i)


def test_large_arena_vs_oracle_and_state_dict_roundtrip():
    shapes = [(1024, 1027), (4099,), (384, 1024), (3,)]
    model = Holder(shapes, tag="optim.big").to(dev())
    cpu_params = [H.det_uniform(f"optim.big{i}", shp) for i, shp in enumerate(shapes)]
    hyper = dict(lr=3e-3, betas=(0.8, 0.99), eps=1e-8, weight_decay=0.05)
    sched = OO.EMASchedule(beta=0.999, update_after_step=1, update_every=1)
    side = OO.OptimizerSide(cpu_params, max_norm=2.0, ema=sched, **hyper)
    ema = NO.EMA(model, beta=0.999, update_after_step=1, update_every=1, include_online_model=False)
    opt = NO.AdamW(model.parameters(), max_grad_norm=2.0, bf16_copy=True, **hyper)
    opt.attach_ema(ema)
    for step in range(4):
        grads = [H.det_uniform(f"optim.bigg{i}.{step}", shp) * (1e-3 if step == 1 else 1e-2) for i, shp in enumerate(shapes)]
        for p, g in zip(model.ps, grads):
            p.grad.copy_(g.to(dev()))
        opt.step()
        ema.update()
        # clip coefficient from the fp64 norm: torch's CPU fp32 norm over 1M elements is itself off by ~8e-6 (summation order),
        # the CUDA two-pass sum is checked against fp64 to 1e-6 right here
        norm64 = torch.sqrt(sum(g.double().square().sum() for g in grads)).float()
        report(f"grad norm step {step}", opt.total_grad_norm(), norm64.reshape(1), 1e-6, 0)
        side.step(grads, coef=torch.clamp(2.0 / (norm64 + 1e-6), max=1.0))
        if step == 1:  # checkpoint in the middle: a fresh optimizer restored from state_dict must continue identically
            sd = opt.state_dict()
            assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and sd["param_groups"][0]["lr"] == 3e-3
    sync()
    for i, p in enumerate(model.ps):
        report(f"big param {i}", p, side.params[i], RTOL, ATOL)
        report(f"big ema {i}", list(ema.ema_model.parameters())[i], side.ema[i], RTOL, ATOL)
        assert torch.equal(opt.bf16_params()[i], p.detach().bfloat16())
    # restore: moments and step counter come back, next step matches an uninterrupted run
    model2 = Holder(shapes, tag="optim.big").to(dev())
    opt2 = NO.AdamW(model2.parameters(), max_grad_norm=2.0, **hyper)
    opt2.load_state_dict(opt.state_dict())
    with torch.no_grad():
        for p2, p in zip(model2.ps, model.ps):
            p2.copy_(p)
    g5 = [H.det_uniform(f"optim.bigg{i}.5", shp) * 1e-2 for i, shp in enumerate(shapes)]
    for o, m in ((opt, model), (opt2, model2)):
        for p, g in zip(m.ps, g5):
            p.grad.copy_(g.to(dev()))
        o.step()
    for p2, p in zip(model2.ps, model.ps):
        assert torch.equal(p2, p)


def test_ema_lerp_matches_torch_cuda_bitwise():
    n = 1 << 20
    e = H.det_uniform("optim.e", (n,)).to(dev())
    p = H.det_uniform("optim.q", (n,)).to(dev())
    for w in (1e-4, 0.2152, 0.5, 0.63, 1.0):
        ref = e.clone().lerp_(p, w)
        got = e.clone()
        call("bsi_ema_update", L.ptr(got), L.ptr(p), n, w, 2, L.stream_ptr())
        sync()
        assert torch.equal(got, ref), w
    got = e.clone()
    call("bsi_ema_update", L.ptr(got), L.ptr(p), n, 0.0, 1, L.stream_ptr())
    sync()
    assert torch.equal(got, p)


def test_grad_sumsq_deterministic_and_accurate():
    lib = L.load()
    n = 479_000_064 // 8  # an eighth of DiT-L's parameter count keeps the test light
    g = torch.randn(n, device=dev(), generator=torch.Generator(device=dev()).manual_seed(5)) * 0.01
    ws = torch.empty(int(lib.bsi_grad_sumsq_workspace_floats()), device=dev())
    outs = []
    for _ in range(2):
        out = torch.zeros(1, device=dev())
        call("bsi_grad_sumsq", L.ptr(out), L.ptr(ws), L.ptr(g), n, L.stream_ptr())
        sync()
        outs.append(out.cpu())
    assert torch.equal(outs[0], outs[1])
    ref = float(g.double().square().sum())
    assert abs(float(outs[0]) - ref) <= 2e-6 * ref


def test_bad_arguments_fail_loudly():
    lib = L.load()
    x = torch.zeros(6, device=dev())
    a = L.AdamWArgs(param=L.ptr(x), grad=L.ptr(x), exp_avg=L.ptr(x), exp_avg_sq=L.ptr(x), numel=6, step=1, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8)
    assert lib.bsi_adamw_ema_step(ctypes.byref(a), L.stream_ptr()) != 0 and b"multiple of 4" in lib.bsi_last_error()
    with pytest.raises(RuntimeError, match="fp32 CUDA"):
        NO.AdamW([torch.nn.Parameter(torch.zeros(4))])
    with pytest.raises(NotImplementedError):
        NO.AdamW([torch.nn.Parameter(torch.zeros(4, device=dev()))], amsgrad=True)


def test_training_step_through_bsi_train_loss():
    """One full optimisation step as the reference's BSITraining runs it (bsi/tasks/bsi.py:186-198): train_loss -> backward
    (autograd accumulates straight into the gradient arena) -> clip + AdamW + EMA, against the oracle fed the same gradients."""
    from test_gpu_bsi_api import make

    bsi, sd = make()
    params = list(bsi.model.parameters())
    ema = NO.create_ema(bsi.model, beta=0.9999, update_after_step=0, update_every=1)
    opt = NO.AdamW(params, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
    opt.attach_ema(ema)
    cpu = [p.detach().cpu().clone() for p in params]
    side = OO.OptimizerSide(cpu, max_norm=1.0, ema=OO.EMASchedule(beta=0.9999, update_after_step=0, update_every=1))
    x = H.det_images("toy.x", 32, (3, 32, 32), seed=2).to(dev())
    gen = torch.Generator(device=dev()).manual_seed(3)
    for it in range(3):
        opt.zero_grad()
        state = gen.get_state()
        loss = bsi.train_loss(x, gen).mean()
        loss.backward()
        for i, p in enumerate(params):
            assert p.grad.data_ptr() == opt._g.view(i).data_ptr()  # accumulated in place, no gather needed
        gen.set_state(state)
        ref_grads = torch.autograd.grad(bsi.train_loss(x, gen).mean(), params)
        for p, g in zip(params, ref_grads):
            assert torch.equal(p.grad, g)
        opt.step()
        ema.update()
        side.step([g.cpu() for g in ref_grads])
        for i, p in enumerate(params):
            report(f"iteration {it} param {i}", p, side.params[i], RTOL, ATOL)
            report(f"iteration {it} ema {i}", list(ema.ema_model.parameters())[i], side.ema[i], RTOL, ATOL)
    assert ema.step == 3 and ema.initted and float(opt.state[params[0]]["step"]) == 3


def test_evaluation_after_native_step_sees_the_new_weights():
    """The fused optimizer / EMA kernels write parameters through raw pointers (no torch version bump): the packed bf16 arena of
    the native denoiser must still be rebuilt, for the online model and for the EMA copy (validation during training)."""
    from test_gpu_dit import SPECS, build

    spec = SPECS["small64"]
    m, sd = build(spec)
    mu = H.det_uniform("st.mu", (2, *spec.data_shape)).to(dev())
    t = torch.tensor([0.3, 0.8], device=dev())
    m.requires_grad_(True)
    ema = NO.create_ema(m, beta=0.9, update_after_step=0, update_every=1).to(dev())
    opt = NO.AdamW(m.parameters(), lr=5e-2, weight_decay=0.0, max_grad_norm=None)
    opt.attach_ema(ema)
    with torch.no_grad():
        y0 = m(mu, t).clone()
        e0 = ema.ema_model(mu, t).clone()  # packs the EMA copy's arena as well
    for it in range(3):  # step 1 copies (mode 1, buffer copy_ bumps versions); steps 2-3 lerp through the kernel only
        for i, p in enumerate(m.parameters()):
            p.grad.copy_(H.det_uniform(f"st.g{i}.{it}", tuple(p.shape)).to(dev()))
        opt.step()
        ema.update()
    with torch.no_grad():
        y1 = m(mu, t)
        e1 = ema.ema_model(mu, t)
        sync()
    new_sd = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
    ema_sd = {k: v.detach().float().cpu() for k, v in ema.ema_model.state_dict().items()}
    O = H.O
    ref1 = O.dit_forward(new_sd, spec, mu.cpu(), t.cpu())
    refe = O.dit_forward(ema_sd, spec, mu.cpu(), t.cpu())
    rel = lambda a, b: float((a.cpu() - b).norm() / b.norm())
    assert rel(y0, ref1) > 5e-2, "the optimizer steps were meant to move the output"
    assert rel(y1, ref1) < 1.5e-2, f"online model evaluated with stale packed weights: {rel(y1, ref1)}"
    assert rel(e1, refe) < 1.5e-2, f"EMA model evaluated with stale packed weights: {rel(e1, refe)}"
    assert rel(e0, refe) > 1e-2
