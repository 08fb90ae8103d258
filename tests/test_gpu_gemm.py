"""tcgen05 GEMM (+ fused epilogues), LayerNorm-modulate and attention kernels vs plain torch fp32 on the same bf16 inputs."""

import math

import pytest
import torch
import torch.nn.functional as F

import helpers as H
from gpu_util import call, dev, report, sync
from bsi_b200 import _lib as L

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[2, 1], ids=["cta_pair", "single_cta"])
def cta_group(request):
    """Every test of this file runs against both GEMM kernels: the cta_group::2 pair kernel and the single-CTA one."""
    call("bsi_gemm_force_cta_group", request.param)
    yield request.param
    call("bsi_gemm_force_cta_group", 0)


def rnd(tag, shape, scale=1.0):
    return (scale * H.det_uniform(tag, shape)).to(dev())


def gemm(A, W, C, bias, epi, **kw):
    a = L.GemmArgs()
    a.A, a.W, a.C, a.bias = L.ptr(A), L.ptr(W), L.ptr(C), L.ptr(bias)
    batch = kw.get("batch", 1)
    a.M, a.N, a.K = kw.get("M", A.shape[-2]), kw.get("N", W.shape[-2]), kw.get("K", A.shape[-1])
    a.lda, a.ldw, a.ldc = kw.get("lda", A.shape[-1]), kw.get("ldw", W.shape[-1]), kw.get("ldc", C.shape[-1])
    a.batch = batch
    a.stride_a, a.stride_w = kw.get("stride_a", a.M * a.lda), kw.get("stride_w", a.N * a.ldw)
    a.stride_c, a.stride_bias = kw.get("stride_c", a.M * a.ldc), kw.get("stride_bias", a.N)
    a.epilogue = epi
    a.gate = kw.get("gate", L.RowRef(None, 0, 0))
    a.step_ptr = kw.get("step_ptr", None)
    a.rows_per_sample = kw.get("rows_per_sample", 0)
    a.pos = kw.get("pos", None)
    a.patch, a.grid_w, a.channels = kw.get("patch", 0), kw.get("grid_w", 0), kw.get("channels", 0)
    a.aux = L.ptr(kw.get("aux", None))
    import ctypes

    call("bsi_gemm_bf16", ctypes.byref(a), L.stream_ptr())
    sync()


@pytest.mark.parametrize(
    "M,N,K",
    [(128, 256, 64), (128, 256, 256), (256, 512, 1024), (384, 1024, 4096), (257, 6144, 1024), (300, 48, 1024), (256, 1024, 336), (2, 384, 128), (2048, 3072, 1024)],
)
def test_gemm_bias_f32(M, N, K):
    A = rnd(f"g.a{M}{N}{K}", (M, K)).bfloat16()
    W = rnd(f"g.w{M}{N}{K}", (N, K), 1 / math.sqrt(K)).bfloat16()
    bias = rnd(f"g.b{M}{N}{K}", (N,), 0.1)
    C = torch.full((M, N), float("nan"), device=dev())
    gemm(A, W, C, bias, L.EPI_BIAS_F32)
    ref = A.float() @ W.float().T + bias
    report(f"gemm f32 {M}x{N}x{K}", C, ref, 2e-4, 2e-4)


def test_gemm_bf16_epilogues():
    M, N, K = 512, 768, 256
    A = rnd("e.a", (M, K)).bfloat16()
    W = rnd("e.w", (N, K), 2 / math.sqrt(K)).bfloat16()
    bias = rnd("e.b", (N,), 0.1)
    pre = A.float() @ W.float().T + bias
    for epi, fn in ((L.EPI_BIAS_BF16, lambda v: v), (L.EPI_BIAS_GELU_BF16, lambda v: F.gelu(v, approximate="tanh")), (L.EPI_BIAS_SILU_BF16, F.silu)):
        C = torch.zeros((M, N), dtype=torch.bfloat16, device=dev())
        gemm(A, W, C, bias, epi)
        report(f"gemm bf16 epilogue {epi}", C, fn(pre), 1e-2, 4e-3)
    # no bias
    C = torch.zeros((M, N), dtype=torch.bfloat16, device=dev())
    gemm(A, W, C, None, L.EPI_BIAS_BF16)
    report("gemm no bias", C, A.float() @ W.float().T, 1e-2, 4e-3)


@pytest.mark.parametrize("M,N,K", [(512, 768, 256), (1000, 4096, 1024), (32768, 1024, 512)])
def test_gemm_training_epilogues_gelu_dual_and_gelu_grad(M, N, K):
    """Training path of the MLP (autograd of dit.py:71-76): the first Linear writes its bf16 pre-activation AND gelu of that rounded
    value from one accumulator tile; the data-gradient GEMM multiplies by gelu'(pre) in its epilogue.  Ragged M covers the TMA
    clipping of both outputs / the auxiliary load."""
    A = rnd(f"t.a{M}", (M, K)).bfloat16()
    W = rnd(f"t.w{M}", (N, K), 2 / math.sqrt(K)).bfloat16()
    bias = rnd(f"t.b{M}", (N,), 0.1)
    pre_ref = (A.float() @ W.float().T + bias).bfloat16()
    act = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=dev())
    pre = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=dev())
    gemm(A, W, act, bias, L.EPI_BIAS_GELU_DUAL_BF16, aux=pre)
    report("dual: pre-activation", pre, pre_ref, 1e-2, 4e-3)
    report("dual: gelu of the stored pre-activation", act, F.gelu(pre.float(), approximate="tanh"), 1e-2, 2e-3)
    # backward: dpre = (dY W2) * gelu'(pre), against autograd of F.gelu on the stored pre-activation
    dY = rnd(f"t.dy{M}", (M, K)).bfloat16()
    x = pre.float().requires_grad_(True)
    up = dY.float() @ W.float().T
    (F.gelu(x, approximate="tanh") * up).sum().backward()
    dpre = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=dev())
    gemm(dY, W, dpre, None, L.EPI_MUL_GELU_GRAD_BF16, aux=pre)
    report("gelu-grad epilogue", dpre, x.grad, 2e-2, 1e-2 * float(up.abs().mean()))
    rel = float((dpre.float() - x.grad).norm() / x.grad.norm())
    assert rel < 5e-3, f"relative L2 error {rel}"
    with pytest.raises(L.BsiNativeError):
        gemm(A[:128], W, act[:128], bias, L.EPI_BIAS_GELU_DUAL_BF16, aux=pre[:128])  # single-CTA variant is not built: fails loudly


def test_gemm_gate_residual_pos_unpatch():
    T, B, N, K = 256, 3, 256, 128
    M = B * T
    A = rnd("r.a", (M, K)).bfloat16()
    W = rnd("r.w", (N, K), 1 / math.sqrt(K)).bfloat16()
    bias = rnd("r.b", (N,), 0.1)
    x0 = rnd("r.x", (M, N))
    pre = A.float() @ W.float().T + bias
    # per-sample gate rows [B, 3N] with the gate in columns [N, 2N)
    table = rnd("r.g", (B, 3 * N))
    x = x0.clone()
    gemm(A, W, x, bias, L.EPI_GATE_RESID_F32, gate=L.rowref(table, 3 * N, 0, N), rows_per_sample=T)
    gate = table[:, N : 2 * N].repeat_interleave(T, dim=0)
    report("gate+residual per-sample", x, x0 + gate * pre, 2e-4, 2e-4)
    # broadcast gate row selected by the device step counter (sampler mode): rows [steps, 3N]
    step = torch.tensor([2], dtype=torch.int32, device=dev())
    x = x0.clone()
    gemm(A, W, x, bias, L.EPI_GATE_RESID_F32, gate=L.rowref(table, 0, 3 * N, N), rows_per_sample=T, step_ptr=L.ptr(step))
    report("gate+residual step-indexed", x, x0 + table[2, N : 2 * N] * pre, 2e-4, 2e-4)
    pos = rnd("r.p", (T, N))
    C = torch.zeros((M, N), device=dev())
    gemm(A, W, C, bias, L.EPI_POS_F32, pos=L.ptr(pos), rows_per_sample=T)
    report("pos epilogue", C, pre + pos.repeat(B, 1), 2e-4, 2e-4)
    # unpatchify: N = p*p*C, tokens on a 16x16 grid
    p, ch, gw = 4, 3, 16
    Wd = rnd("r.wd", (p * p * ch, K), 1 / math.sqrt(K)).bfloat16()
    bd = rnd("r.bd", (p * p * ch,), 0.1)
    out = torch.zeros((B, ch, 64, 64), device=dev())
    gemm(A, Wd, out, bd, L.EPI_UNPATCH_F32, ldc=p * p * ch, rows_per_sample=T, patch=p, grid_w=gw, channels=ch)
    y = (A.float() @ Wd.float().T + bd).cpu()
    report("unpatchify epilogue", out, H.O.unpatchify(y.reshape(B, T, -1), p, 16, gw), 2e-4, 2e-4)


def test_gemm_batched_shared_and_strided_a():
    Lb, M, N, K = 3, 130, 512, 256
    A = rnd("b.a", (M, K)).bfloat16()
    W = rnd("b.w", (Lb, N, K), 1 / math.sqrt(K)).bfloat16()
    bias = rnd("b.b", (Lb, N), 0.1)
    C = torch.zeros((Lb, M, N), device=dev())
    gemm(A, W, C, bias, L.EPI_BIAS_F32, batch=Lb, stride_a=0, M=M, N=N, K=K)
    ref = torch.einsum("mk,lnk->lmn", A.float(), W.float()) + bias[:, None]
    report("batched gemm shared A", C, ref, 2e-4, 2e-4)
    # A for batch l = columns [l*K, (l+1)*K) of a [M, Lb*K] matrix (the adaLN second Linear layout)
    A2 = rnd("b.a2", (M, Lb * K)).bfloat16()
    gemm(A2, W, C, bias, L.EPI_BIAS_F32, batch=Lb, stride_a=K, lda=Lb * K, M=M, N=N, K=K)
    ref = torch.einsum("mlk,lnk->lmn", A2.float().reshape(M, Lb, K), W.float()) + bias[:, None]
    report("batched gemm strided A", C, ref, 2e-4, 2e-4)


def test_gemm_many_tiles_persistent():
    # more tiles than SMs so every CTA loops (phase bookkeeping across tiles) — the QKV shape at batch 32
    M, N, K = 8192, 3072, 1024
    A = rnd("p.a", (M, K)).bfloat16()
    W = rnd("p.w", (N, K), 1 / math.sqrt(K)).bfloat16()
    bias = rnd("p.b", (N,), 0.1)
    C = torch.zeros((M, N), dtype=torch.bfloat16, device=dev())
    gemm(A, W, C, bias, L.EPI_BIAS_BF16)
    report("persistent gemm", C, A.float() @ W.float().T + bias, 1e-2, 4e-3)


@pytest.mark.parametrize("dim", [128, 1024])
def test_layernorm_modulate(dim):
    T, B = 256, 3
    M = B * T
    x = rnd(f"ln.x{dim}", (M, dim), 3.0) + 0.5
    table = rnd(f"ln.t{dim}", (B, 6 * dim), 0.5)
    out = torch.zeros((M, dim), dtype=torch.bfloat16, device=dev())
    call("bsi_layernorm_mod_bf16", L.ptr(out), L.ptr(x), L.rowref(table, 6 * dim, 0, 0), L.rowref(table, 6 * dim, 0, dim), None, None, None, T, M, dim,
         1e-5, L.stream_ptr())
    sync()
    shift, scale = table[:, :dim].repeat_interleave(T, 0), table[:, dim : 2 * dim].repeat_interleave(T, 0)
    ref = torch.addcmul(shift, scale + 1, F.layer_norm(x, (dim,), eps=1e-5))
    report(f"layernorm+modulate {dim}", out, ref, 1e-2, 1e-2)
    gamma, beta = rnd("ln.g", (dim,)) + 1, rnd("ln.b", (dim,), 0.2)
    call("bsi_layernorm_mod_bf16", L.ptr(out), L.ptr(x), L.RowRef(None, 0, 0), L.RowRef(None, 0, 0), None, L.ptr(gamma), L.ptr(beta), T, M, dim, 1e-5,
         L.stream_ptr())
    sync()
    report(f"layernorm affine {dim}", out, F.layer_norm(x, (dim,), gamma, beta, 1e-5), 1e-2, 1e-2)


@pytest.mark.parametrize("legacy", [0, 1], ids=["tcgen05", "mma_sync"])
@pytest.mark.parametrize("heads,B", [(2, 2), (16, 3)])
def test_attention(heads, B, legacy):
    T, hd = 256, 64
    call("bsi_attention_force_legacy", legacy)
    dim = heads * hd
    qkv = rnd(f"at.{heads}", (B * T, 3 * dim), 2.0).bfloat16()
    out = torch.zeros((B * T, dim), dtype=torch.bfloat16, device=dev())
    call("bsi_attention_bf16", L.ptr(out), L.ptr(qkv), B, T, heads, hd, L.stream_ptr())
    sync()
    q, k, v = qkv.float().reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B * T, dim)
    call("bsi_attention_force_legacy", 0)
    report(f"attention h{heads} legacy={legacy}", out, ref, 2e-2, 1e-2)


def test_patch_operand_and_time_embed():
    B, C, Hh, Ww, p = 2, 3, 64, 64, 4
    mu = rnd("po.mu", (B, C, Hh, Ww), 1.5)
    scale = torch.tensor([0.5, 1.0], device=dev())
    cin = C * 7
    P = p * p * cin
    A = torch.full((B * 256, P), 7.0, dtype=torch.bfloat16, device=dev())
    call("bsi_dit_patch_operand", L.ptr(A), L.ptr(mu), L.rowref(scale, 1), None, B, C, Hh, Ww, p, 6, 8, P, L.stream_ptr())
    sync()
    xin = (scale[:, None, None, None] * mu).cpu()
    ref = H.O.patchify(H.O.with_fourier(xin, (6, 8)), p).reshape(B * 256, P)
    report("patch operand", A, ref, 1e-2, 1e-2)
    # in32 geometry with pitch padding (84 -> 88)
    mu2 = rnd("po.mu2", (B, C, 32, 32), 1.5)
    A2 = torch.full((B * 256, 88), 7.0, dtype=torch.bfloat16, device=dev())
    one = torch.ones(1, device=dev())
    call("bsi_dit_patch_operand", L.ptr(A2), L.ptr(mu2), L.rowref(one, 0), None, B, C, 32, 32, 2, 6, 8, 88, L.stream_ptr())
    sync()
    ref2 = H.O.patchify(H.O.with_fourier(mu2.cpu(), (6, 8)), 2).reshape(B * 256, 84)
    report("patch operand in32", A2[:, :84], ref2, 1e-2, 1e-2)
    assert float(A2[:, 84:].float().abs().max()) == 0.0
    t = torch.tensor([0.0, 1 / 256, 0.5, 1.0], device=dev())
    sc, bi = H.O.nyquist_tables(1024, 1000)
    o32 = torch.empty(4, 1024, device=dev())
    sc_d, bi_d = sc.to(dev()), bi.to(dev())
    call("bsi_time_embed", None, L.ptr(o32), L.ptr(t), L.ptr(sc_d), L.ptr(bi_d), 4, 1024, L.stream_ptr())
    sync()
    report("time embed", o32, H.load_golden("embed.pt")["nyq1024_1000"], 1e-5, 2e-4)
