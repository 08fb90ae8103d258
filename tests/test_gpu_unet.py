"""Native VDM U-Net path: implicit-GEMM convolutions, GroupNorm operand kernel, d=128 attention, the engine and BSI on top of it."""

import ctypes
import math

import pytest
import torch
import torch.nn.functional as F

import helpers as H
from gpu_util import call, dev, report, sync
from bsi_b200 import BSI, Discretization
from bsi_b200 import _lib as L
from bsi_b200.models import DenoisingVDMUNet, NyquistPositionalEmbedding
from bsi_b200.nn import FourierFeatures

O = H.O
pytestmark = pytest.mark.gpu
C32 = O.make_consts(1e-2, 1e6, 2e6)
HYPER = dict(lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6, preconditioning="edm")


def rnd(tag, shape, scale=1.0):
    return (scale * H.det_uniform(tag, shape)).to(dev())


def pack_weight(w, cpad=None):
    """fp32 [N][Cin][kh][kw] -> bf16 [N][taps][cpad] through the library's packing kernel."""
    N, cin, kh, kw = w.shape
    cpad = cpad or cin
    out = torch.zeros((N, kh * kw, cpad), dtype=torch.bfloat16, device=dev())
    wc = w.contiguous()
    call("bsi_pack_conv_weight", L.ptr(out), L.ptr(wc), N, cin, kh * kw, cpad, 0, cpad, L.stream_ptr())
    sync()
    return out


def conv(x1, w_packed, y, bias, epi, taps, x2=None, resid=None, scale=None, shift=None, gn_partial=None):
    a = L.ConvArgs()
    B, Hh, Ww, C1 = x1.shape
    a.X1, a.X2, a.W, a.Y, a.bias, a.resid = L.ptr(x1), L.ptr(x2), L.ptr(w_packed), L.ptr(y), L.ptr(bias), L.ptr(resid)
    a.B, a.H, a.Wd, a.C1, a.C2, a.N, a.taps, a.ldc, a.epilogue = B, Hh, Ww, C1, (x2.shape[3] if x2 is not None else 0), w_packed.shape[0], taps, y.shape[-1], epi
    a.scale = scale if scale is not None else L.RowRef(None, 0, 0)
    a.shift = shift if shift is not None else L.RowRef(None, 0, 0)
    a.step_ptr = None
    a.gn_partial = L.ptr(gn_partial)
    call("bsi_conv_bf16", ctypes.byref(a), L.stream_ptr())
    sync()


def ref_conv(x_nhwc, w, bias, pad):
    y = F.conv2d(x_nhwc.float().permute(0, 3, 1, 2), w.bfloat16().float(), bias, padding=pad)
    return y.permute(0, 2, 3, 1).reshape(-1, w.shape[0])


@pytest.mark.parametrize("cg", [2, 1])
def test_conv3x3_and_1x1_vs_torch(cg):
    call("bsi_gemm_force_cta_group", cg)
    try:
        B, Hh, Ww, C, N = 3, 32, 32, 128, 128
        x = rnd("cv.x", (B, Hh, Ww, C)).bfloat16()
        w = rnd("cv.w", (N, C, 3, 3), 1 / math.sqrt(9 * C))
        bias = rnd("cv.b", (N,), 0.1)
        y = torch.full((B * Hh * Ww, N), float("nan"), device=dev())
        conv(x, pack_weight(w), y, bias, L.EPI_BIAS_F32, 9)
        report(f"conv3x3 cg{cg}", y, ref_conv(x, w, bias, 1), 2e-3, 2e-3)
        # N = 384 (attention qkv), bf16 out
        w3 = rnd("cv.w3", (384, C, 3, 3), 1 / math.sqrt(9 * C))
        b3 = rnd("cv.b3", (384,), 0.1)
        y3 = torch.zeros((B * Hh * Ww, 384), dtype=torch.bfloat16, device=dev())
        conv(x, pack_weight(w3), y3, b3, L.EPI_BIAS_BF16, 9)
        report(f"conv3x3 N=384 bf16 cg{cg}", y3, ref_conv(x, w3, b3, 1), 1e-2, 1e-2)
        # two sources (channel concat) 3x3 with modulation + SiLU epilogue, and 1x1 over the concat
        x2 = rnd("cv.x2", (B, Hh, Ww, C)).bfloat16()
        wc = rnd("cv.wc", (N, 2 * C, 3, 3), 1 / math.sqrt(18 * C))
        table = rnd("cv.t", (B, 2 * N), 0.5)
        h = torch.zeros((B * Hh * Ww, N), dtype=torch.bfloat16, device=dev())
        conv(x, pack_weight(wc), h, bias, L.EPI_MOD_SILU_BF16, 9, x2=x2, scale=L.rowref(table, 2 * N, 0, 0), shift=L.rowref(table, 2 * N, 0, N))
        pre = ref_conv(torch.cat((x, x2), dim=3), wc, bias, 1).reshape(B, Hh * Ww, N)
        ref = F.silu(table[:, None, N:] + (1 + table[:, None, :N]) * pre).reshape(-1, N)
        report(f"conv3x3 two-source mod+silu cg{cg}", h, ref, 1e-2, 1e-2)
        w1 = rnd("cv.w1", (N, 2 * C, 1, 1), 1 / math.sqrt(2 * C))
        y1 = torch.zeros((B * Hh * Ww, N), device=dev())
        conv(x, pack_weight(w1), y1, bias, L.EPI_BIAS_F32, 1, x2=x2)
        report(f"conv1x1 two-source cg{cg}", y1, ref_conv(torch.cat((x, x2), dim=3), w1, bias, 0), 2e-3, 2e-3)
        # residual epilogue: Y = resid + conv(x) (gate = 1), resid separate from Y and in place
        r0 = rnd("cv.r", (B * Hh * Ww, N))
        yo = torch.zeros_like(r0)
        conv(x, pack_weight(w), yo, bias, L.EPI_GATE_RESID_F32, 9, resid=r0)
        report(f"conv3x3 + residual (out of place) cg{cg}", yo, r0 + ref_conv(x, w, bias, 1), 2e-3, 2e-3)
        yi = r0.clone()
        conv(x, pack_weight(w), yi, bias, L.EPI_GATE_RESID_F32, 9)
        report(f"conv3x3 + residual (in place) cg{cg}", yi, r0 + ref_conv(x, w, bias, 1), 2e-3, 2e-3)
        # padded input channels (encode conv: 21 -> 64)
        xe = torch.zeros((B, Hh, Ww, 64), dtype=torch.bfloat16, device=dev())
        xe[..., :21] = rnd("cv.xe", (B, Hh, Ww, 21)).bfloat16()
        we = rnd("cv.we", (N, 21, 3, 3), 1 / math.sqrt(9 * 21))
        ye = torch.zeros((B * Hh * Ww, N), device=dev())
        conv(xe, pack_weight(we, 64), ye, bias, L.EPI_BIAS_F32, 9)
        report(f"conv3x3 padded channels cg{cg}", ye, ref_conv(xe[..., :21], we, bias, 1), 2e-3, 2e-3)
    finally:
        call("bsi_gemm_force_cta_group", 0)


@pytest.mark.parametrize("cg", [2, 1])
@pytest.mark.parametrize("Hh,Ww", [(16, 16), (16, 8), (4, 64)], ids=["w16", "w8", "w64_no_row_reuse"])
def test_conv3x3_other_image_widths(Hh, Ww, cg):
    """The narrow-N convolution stages one haloed A box per (dx, channel block) and reads the three dy taps through smem
    descriptors W pixels apart: check the other legal row widths, and W = 64 where the box does not fit and every tap is loaded."""
    call("bsi_gemm_force_cta_group", cg)
    try:
        B, C, N = 5, 128, 128
        x = rnd(f"cw.x{Ww}", (B, Hh, Ww, C)).bfloat16()
        w = rnd("cw.w", (N, C, 3, 3), 1 / math.sqrt(9 * C))
        bias = rnd("cw.b", (N,), 0.1)
        y = torch.full((B * Hh * Ww, N), float("nan"), device=dev())
        conv(x, pack_weight(w), y, bias, L.EPI_BIAS_F32, 9)
        report(f"conv3x3 W={Ww} cg{cg}", y, ref_conv(x, w, bias, 1), 2e-3, 2e-3)
    finally:
        call("bsi_gemm_force_cta_group", 0)


@pytest.mark.parametrize("cg", [2, 1])
def test_groupnorm_statistics_from_the_conv_epilogue(cg):
    """The residual-stream convolution leaves per-tile (sum, sum of squares) of every 4-channel group behind; GroupNorm + SiLU from
    those statistics (one streaming pass) must equal torch's GroupNorm of the tensor the convolution wrote, for groups of 4 and 8."""
    call("bsi_gemm_force_cta_group", cg)
    try:
        B, Hh, Ww, C, N = 3, 32, 32, 128, 128
        x = rnd("gs.x", (B, Hh, Ww, C)).bfloat16()
        w = rnd("gs.w", (N, C, 3, 3), 1 / math.sqrt(9 * C))
        bias = rnd("gs.b", (N,), 0.1)
        r0 = rnd("gs.r", (B * Hh * Ww, N), 2.0) + 0.3
        y = torch.zeros_like(r0)
        part = torch.full((B * Hh * Ww // 128, 32, 2), float("nan"), device=dev())
        conv(x, pack_weight(w), y, bias, L.EPI_GATE_RESID_F32, 9, resid=r0, gn_partial=part)
        report(f"conv output with statistics cg{cg}", y, r0 + ref_conv(x, w, bias, 1), 2e-3, 2e-3)
        tiles = y.reshape(-1, 128, 32, 4)  # [tile][pixel][group][channel]
        report("per-tile group sums", part[..., 0], tiles.sum(dim=(1, 3)), 1e-4, 1e-3)
        report("per-tile group sums of squares", part[..., 1], (tiles * tiles).sum(dim=(1, 3)), 1e-4, 1e-3)
        gamma, beta = rnd("gs.g", (C,)) + 1, rnd("gs.be", (C,), 0.2)
        yb = y.reshape(B, Hh * Ww, C)
        for cpg, silu in ((4, 1), (8, 1), (4, 0)):
            act = torch.zeros((B, Hh * Ww, C), dtype=torch.bfloat16, device=dev())
            raw = torch.zeros_like(act)
            call("bsi_groupnorm_apply_bf16", L.ptr(act), L.ptr(raw), L.ptr(y), L.ptr(part), L.ptr(gamma), L.ptr(beta), B, Hh * Ww, C, cpg, 1e-5, silu, L.stream_ptr())
            old = torch.zeros_like(act)
            call("bsi_groupnorm_act_bf16", L.ptr(old), None, L.ptr(y), L.ptr(gamma), L.ptr(beta), B, Hh * Ww, C, cpg, 1e-5, silu, L.stream_ptr())
            sync()
            ref = F.group_norm(yb.permute(0, 2, 1), C // cpg, gamma, beta, 1e-5).permute(0, 2, 1)
            ref = F.silu(ref) if silu else ref
            report(f"groupnorm from epilogue statistics cpg={cpg} silu={silu}", act, ref, 1e-2, 1e-2)
            report("agrees with the stand-alone kernel", act, old, 1e-2, 4e-3)
            report("raw bf16 copy", raw, yb, 4e-3, 4e-3)
    finally:
        call("bsi_gemm_force_cta_group", 0)


def test_groupnorm_act_and_input_and_decode():
    B, HW, C = 3, 1024, 128
    x = rnd("gn.x", (B, HW, C), 2.0) + 0.3
    gamma, beta = rnd("gn.g", (C,)) + 1, rnd("gn.b", (C,), 0.2)
    for cpg, silu in ((4, 1), (8, 1), (4, 0)):
        act = torch.zeros((B, HW, C), dtype=torch.bfloat16, device=dev())
        raw = torch.zeros_like(act)
        call("bsi_groupnorm_act_bf16", L.ptr(act), L.ptr(raw), L.ptr(x), L.ptr(gamma), L.ptr(beta), B, HW, C, cpg, 1e-5, silu, L.stream_ptr())
        sync()
        ref = F.group_norm(x.permute(0, 2, 1), C // cpg, gamma, beta, 1e-5).permute(0, 2, 1)
        ref = F.silu(ref) if silu else ref
        report(f"groupnorm cpg={cpg} silu={silu}", act, ref, 1e-2, 1e-2)
        report("raw bf16 copy", raw, x, 4e-3, 4e-3)
    mu = rnd("gn.mu", (B, 3, 32, 32), 1.5)
    sc = torch.tensor([0.5, 1.0, 2.0], device=dev())
    op = torch.full((B, 1024, 64), 7.0, dtype=torch.bfloat16, device=dev())
    call("bsi_unet_input_bf16", L.ptr(op), L.ptr(mu), L.rowref(sc, 1), None, B, 3, 1024, 6, 8, 64, L.stream_ptr())
    sync()
    ref_in = O.with_fourier((sc[:, None, None, None] * mu).cpu(), (6, 8)).permute(0, 2, 3, 1).reshape(B, 1024, 21)
    report("unet input operand", op[..., :21], ref_in, 1e-2, 1e-2)
    assert float(op[..., 21:].float().abs().max()) == 0.0
    xs = rnd("gn.xs", (B * 1024, C))
    wd, bd = rnd("gn.wd", (3, C), 0.1), rnd("gn.bd", (3,), 0.1)
    out = torch.zeros((B, 3, 32, 32), device=dev())
    call("bsi_unet_decode", L.ptr(out), L.ptr(xs), L.ptr(wd), L.ptr(bd), B, 1024, C, 3, L.stream_ptr())
    sync()
    report("decode 1x1", out, (xs @ wd.T + bd).reshape(B, 1024, 3).permute(0, 2, 1).reshape(B, 3, 32, 32), 1e-5, 1e-5)


def test_attention_d128():
    B, T, d = 2, 1024, 128
    qkv = rnd("a2.qkv", (B * T, 3 * d), 1.5).bfloat16()
    out = torch.zeros((B * T, d), dtype=torch.bfloat16, device=dev())
    call("bsi_attention_d128_bf16", L.ptr(out), L.ptr(qkv), B, T, L.stream_ptr())
    sync()
    q, k, v = qkv.float().reshape(B, T, 3, 1, d).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B * T, d)
    report("attention d128", out, ref, 2e-2, 1e-2)


def build_unet(levels=2):
    spec = O.UNetSpec((3, 32, 32), dim=128, levels=levels)
    m = DenoisingVDMUNet(spec.data_shape, NyquistPositionalEmbedding(32, 100), "silu", 128, levels, 4, n_attention_heads=1, dropout=0.1,
                         fourier_features=FourierFeatures(n_min=6, n_max=8))
    sd = H.det_state_dict(H.unet_shapes(spec), seed=1)
    assert set(m.state_dict()) == set(sd)
    m.load_state_dict(sd)
    return m.to(dev()).eval().requires_grad_(False), sd, spec


def test_unet_forward_vs_reference_golden():
    m, sd, spec = build_unet()
    g = H.load_golden("unet.pt")["dim128"]
    mu = 1.5 * H.det_uniform("unet.mu", (2, *spec.data_shape))
    with torch.inference_mode():
        y = m(mu.to(dev()), torch.tensor([0.2, 0.95], device=dev()))
        sync()
    rel = float((y.cpu() - g["y"]).norm() / g["y"].norm())
    assert rel < 1.5e-2, f"relative L2 error {rel} of the bf16 U-Net engine vs the fp32 reference"
    report("unet forward vs reference golden", y, g["y"], 5e-2, 3e-2 * float(g["y"].abs().max()))


def test_bsi_on_native_unet():
    m, sd, spec = build_unet()
    bsi = BSI(m, data_shape=spec.data_shape, k=8, discretization=Discretization.image_8bit(), **HYPER).to(dev())
    f = lambda mu, t: O.unet_forward(sd, spec, mu, t)
    x = H.det_images("unet.x", 4, spec.data_shape, seed=2)
    bsi.noise_source = "torch"
    with torch.inference_mode():
        gen = torch.Generator(device=dev()).manual_seed(5)
        e, b, ex = bsi.elbo(x.to(dev()), 1, 2, gen)
        gen = torch.Generator(device=dev()).manual_seed(5)
        eps_r = torch.randn((1, 4, *spec.data_shape), device=dev(), generator=gen).cpu()
        off, perm = torch.rand((), device=dev(), generator=gen).cpu(), torch.randperm(8, device=dev(), generator=gen).cpu()
        eps_m = torch.randn((2, 4, *spec.data_shape), device=dev(), generator=gen).cpu()
        l_r = O.recon_loss(f, C32, x, 1, eps_r, O.GRID_8BIT)
        l_m = O.inf_measure_loss(f, C32, x, O.lam_of_t(C32, O.ld_times(2, 4, off, perm)), eps_m)
        _, b_ref, _ = O.combine_elbo(l_r, l_m, 3072)
    assert float((b.cpu() - b_ref).abs().max()) < 2e-3, f"bpd {b.cpu().tolist()} vs oracle {b_ref.tolist()}"
    bsi.noise_source = "philox"
    with torch.inference_mode():
        a = bsi.sample(4, seed=3)
        sync()
        assert int(m.last_sampler_state["step"].item()) == 8
        k, lam, coef, c_in, t_rows = bsi._step_table(bsi.default_schedule)
        bb = m.sample_loop(4, torch.rsqrt(lam[:1]).contiguous(), coef, c_in, t_rows, k, 3, 0, 1, use_graph=False)
        sync()
    assert torch.isfinite(a).all() and torch.equal(a, bb), "graph replay and eager loop disagree"
