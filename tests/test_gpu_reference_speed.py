"""The reference algorithm in eager PyTorch on the same B200 (the oracle's functional DiT-L/4 forward on CUDA, under bf16 autocast
and in TF32) next to the native engine: the "existing Blackwell kernels" bar of SURVEY §8(d).  The oracle may only be imported from
tests/, so this comparison lives here; it records both timings in gpurun_out/reference_on_gpu.json and checks that the outputs agree
and that the native forward is not slower than PyTorch's library kernels (cuBLAS + flash attention)."""

import json
import os

import pytest
import torch

import helpers as H
from gpu_util import OUT_DIR, dev
from bsi_b200.models import DenoisingDiT
from bsi_b200.nn import FourierFeatures

O = H.O
pytestmark = pytest.mark.gpu


def _time(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def test_dit_l_forward_native_vs_pytorch_on_the_same_gpu():
    B = 256
    spec = O.DiTSpec((3, 64, 64), 4, 1024, 24, 16)
    torch.manual_seed(0)
    m = DenoisingDiT(spec.data_shape, spec.patch, spec.dim, spec.depth, spec.heads, dropout=None, fourier_features=FourierFeatures(n_min=6, n_max=8))
    with torch.no_grad():
        for blk in m.dit.blocks:
            torch.nn.init.normal_(blk.adaLN_modulation[-1].weight, std=0.02)
            torch.nn.init.normal_(blk.adaLN_modulation[-1].bias, std=0.02)
    m = m.to(dev()).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    mu = torch.randn((B, *spec.data_shape), device=dev())
    t = torch.rand(B, device=dev())
    res = {}
    with torch.inference_mode():
        y_native = m(mu, t)
        res["native_ms"] = _time(lambda: m(mu, t))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y_bf16 = O.dit_forward(sd, spec, mu, t).float()
            res["pytorch_bf16_autocast_ms"] = _time(lambda: O.dit_forward(sd, spec, mu, t))
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            res["pytorch_tf32_ms"] = _time(lambda: O.dit_forward(sd, spec, mu, t), reps=2)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = False
    flops = B * 161.61e9
    res.update(batch=B, native_tflops=flops / res["native_ms"] / 1e9, pytorch_bf16_tflops=flops / res["pytorch_bf16_autocast_ms"] / 1e9,
               speedup_vs_bf16_autocast=res["pytorch_bf16_autocast_ms"] / res["native_ms"], speedup_vs_tf32=res["pytorch_tf32_ms"] / res["native_ms"],
               rel_l2_native_vs_pytorch_bf16=float((y_native - y_bf16).norm() / y_bf16.norm()))
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, "reference_on_gpu.json"), "w") as fh:
        json.dump(res, fh)
    print(json.dumps(res))
    assert res["rel_l2_native_vs_pytorch_bf16"] < 2e-2
    assert res["speedup_vs_bf16_autocast"] > 1.0, res


def test_dit_l_forward_backward_native_vs_pytorch_autograd():
    """Same comparison for the training direction (config 5's denoiser work): forward + backward of DiT-L/4 at batch 64."""
    B = 64
    spec = O.DiTSpec((3, 64, 64), 4, 1024, 24, 16)
    torch.manual_seed(0)
    m = DenoisingDiT(spec.data_shape, spec.patch, spec.dim, spec.depth, spec.heads, dropout=None, fourier_features=FourierFeatures(n_min=6, n_max=8))
    with torch.no_grad():
        for blk in m.dit.blocks:
            torch.nn.init.normal_(blk.adaLN_modulation[-1].weight, std=0.02)
            torch.nn.init.normal_(blk.adaLN_modulation[-1].bias, std=0.02)
    m = m.to(dev()).train()
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and k in dict(m.named_parameters())) for k, v in m.state_dict().items()}
    mu = torch.randn((B, *spec.data_shape), device=dev())
    t = torch.rand(B, device=dev())
    w = torch.randn_like(mu)

    def native():
        m.zero_grad(set_to_none=True)
        (m(mu, t) * w).sum().backward()

    def pytorch():
        for v in sd.values():
            v.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = O.dit_forward(sd, spec, mu, t)
        (y.float() * w).sum().backward()

    res = {"batch": B, "native_fwd_bwd_ms": _time(native), "pytorch_bf16_autocast_fwd_bwd_ms": _time(pytorch)}
    res["speedup"] = res["pytorch_bf16_autocast_fwd_bwd_ms"] / res["native_fwd_bwd_ms"]
    g_native = m.dit.blocks[11].mlp[0].weight.grad
    g_ref = sd["dit.blocks.11.mlp.0.weight"].grad
    res["rel_l2_grad_mlp_block11"] = float((g_native - g_ref).norm() / g_ref.norm())
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, "reference_on_gpu_train.json"), "w") as fh:
        json.dump(res, fh)
    print(json.dumps(res))
    assert res["rel_l2_grad_mlp_block11"] < 5e-2


def test_unet_forward_native_vs_pytorch_on_the_same_gpu():
    """cifar10-vdm's VDM U-Net (dim 128, 32 levels, batch 256): the oracle's functional forward on CUDA (cuDNN convolutions, channels-last
    is PyTorch's business) under bf16 autocast vs the native engine."""
    from bsi_b200.models import DenoisingVDMUNet, NyquistPositionalEmbedding

    B = 256
    spec = O.UNetSpec((3, 32, 32), dim=128, levels=32)
    m = DenoisingVDMUNet(spec.data_shape, NyquistPositionalEmbedding(32, 100), "silu", 128, 32, 4, n_attention_heads=1, dropout=0.1,
                         fourier_features=FourierFeatures(n_min=6, n_max=8))
    m.load_state_dict(H.det_state_dict(H.unet_shapes(spec), seed=1))
    m = m.to(dev()).eval().requires_grad_(False)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    mu = torch.randn((B, *spec.data_shape), device=dev())
    t = torch.rand(B, device=dev())
    res = {"batch": B}
    with torch.inference_mode():
        y_native = m(mu, t)
        res["native_ms"] = _time(lambda: m(mu, t))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y_ref = O.unet_forward(sd, spec, mu, t).float()
            res["pytorch_bf16_autocast_ms"] = _time(lambda: O.unet_forward(sd, spec, mu, t))
    res["speedup_vs_bf16_autocast"] = res["pytorch_bf16_autocast_ms"] / res["native_ms"]
    res["rel_l2_native_vs_pytorch_bf16"] = float((y_native - y_ref).norm() / y_ref.norm())
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, "reference_on_gpu_unet.json"), "w") as fh:
        json.dump(res, fh)
    print(json.dumps(res))
    assert res["rel_l2_native_vs_pytorch_bf16"] < 5e-2
