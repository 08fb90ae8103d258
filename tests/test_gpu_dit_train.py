"""Differentiable native DiT (bsi_b200/models/dit_train.py, SURVEY §8 a23 first version) vs the CPU oracle under autograd:
forward value, the gradient of every parameter (adaLN and time embedding included), and train_loss end to end.

Tolerance: bf16 tensor-core operands with fp32 accumulation against an fp32 reference: relative L2 error per tensor
< 3e-2 for gradients (two chained bf16 roundings per GEMM input), < 1e-2 for the forward output."""

import pytest
import torch

import helpers as H
from gpu_util import dev, sync
from bsi_b200 import BSI, Discretization
from bsi_b200.models import DenoisingDiT
from bsi_b200.nn import FourierFeatures

O = H.O
pytestmark = pytest.mark.gpu
HYPER = dict(lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6, preconditioning="edm")
C32 = O.make_consts(1e-2, 1e6, 2e6)

SPECS = {
    "small64": O.DiTSpec((3, 64, 64), 4, 128, 2, 2),
    "nofourier32": O.DiTSpec((3, 32, 32), 2, 128, 1, 2, fourier=None),
    "wide": O.DiTSpec((3, 32, 32), 2, 256, 1, 4),
    "L2x64": O.DiTSpec((3, 64, 64), 4, 1024, 2, 16),  # DiT-L width and head count (the 2-CTA 256x256 GEMM tiles, 16 heads), 2 blocks
}


def build(spec, seed=1):
    ff = None if spec.fourier is None else FourierFeatures(n_min=spec.fourier[0], n_max=spec.fourier[1])
    m = DenoisingDiT(spec.data_shape, spec.patch, spec.dim, spec.depth, spec.heads, dropout=None, fourier_features=ff)
    sd = H.det_state_dict(H.dit_shapes(spec), seed=seed)
    m.load_state_dict(sd)
    return m.to(dev()).train(), sd


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("name", list(SPECS))
def test_forward_and_all_gradients_vs_oracle_autograd(name):
    spec = SPECS[name]
    m, sd = build(spec)
    B = 3
    mu = 1.5 * H.det_uniform(f"dt.{name}.mu", (B, *spec.data_shape))
    t = torch.tensor([0.2, 0.55, 0.9])
    w = H.det_uniform(f"dt.{name}.w", (B, *spec.data_shape))
    y = m(mu.to(dev()), t.to(dev()))
    assert y.requires_grad
    (y * w.to(dev())).sum().backward()
    sync()
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    y_ref = O.dit_forward(ref_sd, spec, mu, t)
    (y_ref * w).sum().backward()
    assert rel(y, y_ref) < 1e-2
    worst = {}
    for k, p in m.state_dict(keep_vars=True).items():
        assert p.grad is not None, k
        worst[k] = rel(p.grad, ref_sd[k].grad)
    bad = {k: v for k, v in worst.items() if not v < 3e-2}
    assert not bad, f"gradient mismatch (relative L2): {bad}"


def test_inference_path_is_unchanged():
    spec = SPECS["small64"]
    m, sd = build(spec)
    mu = H.det_uniform("dt.inf.mu", (2, *spec.data_shape)).to(dev())
    t = torch.tensor([0.3, 0.9], device=dev())
    with torch.no_grad():
        y_inf = m(mu, t)
    y_tr = m(mu, t)
    assert not y_inf.requires_grad and y_tr.requires_grad
    assert rel(y_tr, y_inf) < 1e-2  # same kernels, fused differently (fp32 adaLN chain in the training path)


def test_dropout_in_training_mode():
    """dropout=0.05 as in config/experiment/imagenet64.yaml:39: train() applies the two dropout sites (stateless masks keyed by a
    per-call seed), eval() is the identity, a fixed seed reproduces forward and gradients bit for bit."""
    from bsi_b200.models.dit_train import forward_train

    spec = SPECS["small64"]
    m, sd = build(spec)
    ff = FourierFeatures(n_min=6, n_max=8)
    md = DenoisingDiT(spec.data_shape, spec.patch, spec.dim, spec.depth, spec.heads, dropout=0.05, fourier_features=ff)
    md.load_state_dict(sd)
    md = md.to(dev()).train()
    mu = H.det_uniform("dt.drop.mu", (2, *spec.data_shape)).to(dev())
    t = torch.tensor([0.3, 0.9], device=dev())
    y_plain = m(mu, t)
    y1 = forward_train(md, mu, t, None, seed=11)
    y1.square().sum().backward()
    g1 = [p.grad.clone() for p in md.parameters()]
    md.zero_grad()
    y2 = forward_train(md, mu, t, None, seed=11)
    y2.square().sum().backward()
    y3 = forward_train(md, mu, t, None, seed=12)
    assert torch.equal(y1, y2) and all(torch.equal(a, p.grad) for a, p in zip(g1, md.parameters()))
    assert not torch.equal(y1, y3)
    assert 1e-3 < rel(y1, y_plain) < 0.5  # dropout perturbs the output, moderately at p = 0.05
    assert all(torch.isfinite(g).all() and float(g.abs().max()) > 0 for g in g1)
    torch.manual_seed(5)
    ya = md(mu, t)
    torch.manual_seed(5)
    yb = md(mu, t)
    assert torch.equal(ya, yb) and not torch.equal(ya, md(mu, t))  # seeds come from torch's CPU generator
    md.eval()
    assert rel(md(mu, t), y_plain) == 0.0  # eval(): no dropout, same result as the dropout-free model


def test_train_loss_backward_through_native_dit():
    """BSI.train_loss(x).mean().backward() (bsi/bsi.py:291-310, bsi/tasks/bsi.py:186-194) with the native denoiser, against the
    oracle's train_loss on the same lambda grid and noise."""
    spec = SPECS["nofourier32"]
    m, sd = build(spec)
    bsi = BSI(m, data_shape=spec.data_shape, k=16, discretization=Discretization.image_8bit(), **HYPER).to(dev())
    bsi.noise_source = "torch"
    x = H.det_images("dt.x", 8, spec.data_shape, seed=2)
    loss = bsi.train_loss(x.to(dev()), torch.Generator(device=dev()).manual_seed(3))
    loss.mean().backward()
    gen = torch.Generator(device=dev()).manual_seed(3)
    off, perm = torch.rand((), device=dev(), generator=gen).cpu(), torch.randperm(8, device=dev(), generator=gen).cpu()
    eps = torch.randn((1, 8, *spec.data_shape), device=dev(), generator=gen).cpu()[0]
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.train_loss_with(lambda mu, t: O.dit_forward(ref_sd, spec, mu, t), C32, x, O.lam_of_t(C32, O.ld_times(1, 8, off, perm))[0], eps)
    ref.mean().backward()
    assert rel(loss, ref) < 2e-2
    bad = {k: rel(p.grad, ref_sd[k].grad) for k, p in m.state_dict(keep_vars=True).items()}
    bad = {k: v for k, v in bad.items() if not v < 5e-2}
    assert not bad, f"train_loss gradient mismatch (relative L2): {bad}"


def test_gradient_sink_equals_autograd_path():
    """AdamW.attach_model: the backward accumulates weight gradients straight into the optimizer's arena (no autograd += pass);
    the arena must hold exactly what the autograd path produces, also when two backward passes accumulate."""
    from bsi_b200 import optim as NO

    spec = SPECS["small64"]
    mu = 1.5 * H.det_uniform("dt.sink.mu", (2, *spec.data_shape)).to(dev())
    t = torch.tensor([0.25, 0.8], device=dev())
    w = H.det_uniform("dt.sink.w", (2, *spec.data_shape)).to(dev())
    grads = {}
    for sink in (False, True):
        m, _ = build(spec)
        opt = NO.AdamW(m.parameters(), lr=1e-3)
        if sink:
            opt.attach_model(m)
        for _ in range(2):  # gradient accumulation over two passes
            (m(mu, t) * w).sum().backward()
        sync()
        grads[sink] = [p.grad.clone() for p in m.parameters()]
        assert all(p.grad.data_ptr() == opt._g.view(i).data_ptr() for i, p in enumerate(m.parameters()))
    for a, b in zip(grads[False], grads[True]):
        assert rel(b, a) < 1e-5  # same kernels; only the fp32 accumulation order of the split-M partial tiles differs
