"""Backward building blocks for config 5 (not wired into an engine yet): the weight-gradient GEMM on tcgen05
with MN-major operands and split-M TMA reduce-add, and the data-gradient GEMM through the forward kernel with transposed
weights -- each against torch fp32 matmuls of the same bf16-rounded operands."""

import pytest
import torch
import torch.nn.functional as F

import helpers as H
from gpu_util import call, dev, report, sync
from bsi_b200 import _lib as L
from test_gpu_gemm import gemm, rnd

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K,splits", [(4096, 256, 256, 0), (4096, 256, 256, 1), (1000, 384, 320, 3), (8192, 1024, 1024, 0), (320, 136, 72, 0)])
def test_wgrad_vs_torch(M, N, K, splits):
    dy = rnd(f"wg.dy{M}", (M, N)).bfloat16()
    x = rnd(f"wg.x{M}", (M, K)).bfloat16()
    base = rnd(f"wg.acc{M}", (N, K))
    dw = base.clone()
    call("bsi_gemm_wgrad_bf16", L.ptr(dw), L.ptr(dy), L.ptr(x), M, N, K, N, K, K, splits, L.stream_ptr())
    sync()
    ref = base + dy.float().t() @ x.float()
    scale = float(ref.abs().max())
    report(f"wgrad {M}x{N}x{K} splits={splits}", dw, ref, 1e-3, 1e-3 * scale)
    # accumulating twice adds the same product again (autograd's +=)
    call("bsi_gemm_wgrad_bf16", L.ptr(dw), L.ptr(dy), L.ptr(x), M, N, K, N, K, K, splits, L.stream_ptr())
    sync()
    report("wgrad second accumulation", dw, base + 2 * (dy.float().t() @ x.float()), 1e-3, 2e-3 * scale)


def test_wgrad_strided_operands_and_bad_arguments():
    M, N, K = 2048, 256, 192
    big_dy = rnd("wg.sdy", (M, 3 * N)).bfloat16()  # dY is a column slice of a packed [M, 3N] gradient (the QKV layout)
    x = rnd("wg.sx", (M, K)).bfloat16()
    dw = torch.zeros((N, K + 8), device=dev())
    dy = big_dy[:, N : 2 * N]
    call("bsi_gemm_wgrad_bf16", L.ptr(dw), dy.data_ptr(), L.ptr(x), M, N, K, 3 * N, K, K + 8, 0, L.stream_ptr())
    sync()
    report("wgrad strided", dw[:, :K], dy.float().t() @ x.float(), 1e-3, 1e-2)
    assert float(dw[:, K:].abs().max()) == 0.0
    lib = L.load()
    assert lib.bsi_gemm_wgrad_bf16(L.ptr(dw), dy.data_ptr(), L.ptr(x), M, N, 100, 3 * N, K, K + 8, 0, L.stream_ptr()) != 0
    assert b"multiples of 8" in lib.bsi_last_error()


def test_dgrad_through_forward_kernel_with_transposed_weights():
    """dX = dY W: the forward kernel computes A W'^T, so feeding W' = W^T (a bf16 [K][N] copy refreshed once per optimizer
    step) gives the data gradient with no new kernel."""
    M, N, K = 4096, 384, 256
    dy = rnd("dg.dy", (M, N)).bfloat16()
    w = rnd("dg.w", (N, K), 0.05)
    wt = w.t().contiguous().bfloat16()  # [K][N]
    dx = torch.zeros((M, K), device=dev())
    zero = torch.zeros(K, device=dev())
    gemm(dy, wt, dx, zero, L.EPI_BIAS_F32)
    report("dgrad", dx, dy.float() @ w.bfloat16().float(), 1e-3, 1e-3)


def test_gate_residual_forward_backward_vs_autograd():
    B, T, D = 3, 256, 384
    M = B * T
    x0 = rnd("gr.x", (M, D))
    br = rnd("gr.br", (M, D)).bfloat16()
    tab = rnd("gr.tab", (B, 6 * D), 0.7)  # gate lives inside the [B, 6D] modulation table (column block 2), like dit.py:89
    gate = L.rowref(tab, 6 * D, 0, 2 * D)
    x = x0.clone()
    call("bsi_gate_residual", L.ptr(x), L.ptr(x), L.ptr(br), gate, T, M, D, L.stream_ptr())
    x_new = torch.zeros_like(x0)
    call("bsi_gate_residual", L.ptr(x_new), L.ptr(x0), L.ptr(br), gate, T, M, D, L.stream_ptr())
    sync()
    g = tab[:, 2 * D : 3 * D].clone().requires_grad_(True)
    brf = br.float().requires_grad_(True)
    ref = torch.addcmul(x0.reshape(B, T, D), g[:, None], brf.reshape(B, T, D)).reshape(M, D)
    report("gate_residual", x, ref, 1e-6, 1e-6)
    assert torch.equal(x_new, x)  # out of place == in place
    dx = rnd("gr.dx", (M, D))
    ref.backward(dx)
    dbr = torch.zeros((M, D), dtype=torch.bfloat16, device=dev())
    dgate = torch.zeros((B, D), device=dev())
    dbias = torch.zeros((B, D), device=dev())
    call("bsi_gate_residual_backward", L.ptr(dbr), L.ptr(dgate), L.ptr(dbias), L.ptr(dx), L.ptr(br), gate, T, B, D, L.stream_ptr())
    sync()
    report("dbranch", dbr, brf.grad, 1e-2, 1e-3)
    report("dgate", dgate, g.grad, 1e-4, 1e-4)
    report("bias gradient of the branch", dbias.sum(0), dbr.float().sum(0), 1e-5, 1e-4)
    # the row-pipelined kernel (per-CTA partial sums over 32 consecutive rows): same dbranch bit for bit, same sums
    dbr2 = torch.zeros_like(dbr)
    parts = torch.zeros((2, M // 32, D), device=dev())
    call("bsi_gate_residual_backward_rows", L.ptr(dbr2), L.ptr(parts[0]), L.ptr(parts[1]), L.ptr(dx), L.ptr(br), gate, T, 32, M, D, L.stream_ptr())
    sync()
    assert torch.equal(dbr2, dbr)
    report("dgate from row partials", parts[0].view(B, T // 32, D).sum(1), g.grad, 1e-4, 1e-4)
    report("bias gradient from row partials", parts[1].sum(0), dbr.float().sum(0), 1e-5, 1e-4)
    assert L.load().bsi_gate_residual_backward_rows(L.ptr(dbr2), L.ptr(parts[0]), L.ptr(parts[1]), L.ptr(dx), L.ptr(br), gate, T, 48, M, D, L.stream_ptr()) != 0


def test_reduce_rows_several_jobs_in_one_launch():
    """dst[g] (+)= sum_r src[g * rows + r] for several partial buffers at once (the finishing pass of the row kernels' partial sums)."""
    a = rnd("rr.a", (6 * 8, 384))       # 6 groups of 8 partial rows -> a strided destination (column block of a wider table)
    b = rnd("rr.b", (1000, 1024))       # one group: a bias gradient, accumulated into an existing buffer
    c = rnd("rr.c", (2 * 3, 260))       # D not a multiple of 256
    tab = torch.zeros((6, 3 * 384), device=dev())
    acc0 = rnd("rr.acc", (1024,))
    acc = acc0.clone()
    out_c = torch.zeros((2, 260), device=dev())
    jobs = (L.ReduceJob * 3)(L.ReduceJob(L.ptr(a), tab[:, 384:768].data_ptr(), 6, 8, 384, 3 * 384, 0, 0),
                             L.ReduceJob(L.ptr(b), L.ptr(acc), 1, 1000, 1024, 1024, 1, 0),
                             L.ReduceJob(L.ptr(c), L.ptr(out_c), 2, 3, 260, 260, 0, 0))
    L.check(L.load().bsi_reduce_rows(jobs, 3, L.stream_ptr()))
    sync()
    report("per-group sums", tab[:, 384:768], a.view(6, 8, 384).sum(1), 1e-6, 1e-5)
    assert float(tab[:, :384].abs().max()) == 0 and float(tab[:, 768:].abs().max()) == 0
    report("accumulated column sums", acc, acc0 + b.double().sum(0).float(), 1e-5, 1e-4)
    report("ragged width", out_c, c.view(2, 3, 260).sum(1), 1e-6, 1e-5)
    assert L.load().bsi_reduce_rows(jobs, 13, L.stream_ptr()) != 0
    bad = (L.ReduceJob * 1)(L.ReduceJob(L.ptr(a), L.ptr(out_c), 1, 1, 386, 386, 0, 0))
    assert L.load().bsi_reduce_rows(bad, 1, L.stream_ptr()) != 0


def test_colsum_bf16():
    M, N = 1000, 384
    a = rnd("cs.a", (M, 2 * N)).bfloat16()[:, N:]  # strided view (pitch 2N)
    parts = torch.zeros(((M + 255) // 256, N), device=dev())
    call("bsi_colsum_bf16", L.ptr(parts), a.data_ptr(), M, N, 2 * N, 256, L.stream_ptr())
    sync()
    report("colsum", parts.sum(0), a.float().sum(0), 1e-5, 1e-4)


def test_gelu_forward_backward_vs_autograd():
    n = 8 * 5000
    pre = rnd("ge.pre", (n,), 4.0).bfloat16()
    up = rnd("ge.up", (n,)).bfloat16()
    h = torch.zeros_like(pre)
    call("bsi_gelu_bf16", L.ptr(h), L.ptr(pre), n, L.stream_ptr())
    p = pre.float().requires_grad_(True)
    ref = F.gelu(p, approximate="tanh")
    ref.backward(up.float())
    dpre = torch.zeros_like(pre)
    call("bsi_gelu_backward_bf16", L.ptr(dpre), L.ptr(up), L.ptr(pre), n, L.stream_ptr())
    sync()
    report("gelu", h, ref, 1e-2, 1e-3)
    report("gelu backward", dpre, p.grad, 1e-2, 1e-3)


@pytest.mark.parametrize("dim", [128, 1024])
def test_layernorm_mod_backward_vs_autograd(dim):
    B, T, rpc = 3, 256, 32
    M = B * T
    x = rnd(f"lb.x{dim}", (M, dim), 1.5) + 0.3
    da = rnd(f"lb.da{dim}", (M, dim)).bfloat16()
    tab = rnd(f"lb.t{dim}", (B, 6 * dim), 0.5)
    dx0 = rnd(f"lb.dx{dim}", (M, dim))
    # modulated variant
    xr = x.clone().requires_grad_(True)
    sc = tab[:, dim : 2 * dim].clone().requires_grad_(True)
    sh = tab[:, :dim].clone().requires_grad_(True)
    a = torch.addcmul(sh[:, None], sc[:, None] + 1, F.layer_norm(xr, (dim,), eps=1e-5).reshape(B, T, dim)).reshape(M, dim)
    a.backward(da.float())
    dx = dx0.clone()
    parts = torch.zeros((2, M // rpc, dim), device=dev())
    call("bsi_layernorm_mod_backward", L.ptr(dx), L.ptr(parts[0]), L.ptr(parts[1]), L.ptr(da), L.ptr(x), L.rowref(tab, 6 * dim, 0, dim), None, T, rpc, M, dim,
         1e-5, 0.0, 0, L.stream_ptr())
    sync()
    report(f"ln-mod dx {dim}", dx, dx0 + xr.grad, 1e-4, 1e-4)
    report(f"ln-mod dscale {dim}", parts[0].reshape(B, T // rpc, dim).sum(1), sc.grad, 1e-4, 1e-3)
    report(f"ln-mod dshift {dim}", parts[1].reshape(B, T // rpc, dim).sum(1), sh.grad, 1e-4, 1e-3)
    # affine variant (patch decoder LayerNorm)
    gamma = (rnd(f"lb.g{dim}", (dim,), 0.3) + 1).requires_grad_(True)
    beta = rnd(f"lb.b{dim}", (dim,), 0.2).requires_grad_(True)
    xr2 = x.clone().requires_grad_(True)
    F.layer_norm(xr2, (dim,), gamma, beta, 1e-5).backward(da.float())
    dx = torch.zeros_like(x)
    parts.zero_()
    call("bsi_layernorm_mod_backward", L.ptr(dx), L.ptr(parts[0]), L.ptr(parts[1]), L.ptr(da), L.ptr(x), L.RowRef(None, 0, 0), L.ptr(gamma.detach()), T, rpc, M, dim,
         1e-5, 0.0, 0, L.stream_ptr())
    sync()
    report(f"ln-affine dx {dim}", dx, xr2.grad, 1e-4, 1e-4)
    report(f"ln-affine dgamma {dim}", parts[0].sum(0), gamma.grad, 1e-4, 2e-3)
    report(f"ln-affine dbeta {dim}", parts[1].sum(0), beta.grad, 1e-4, 2e-3)


def test_cast_transpose_bf16():
    R, Cc = 100, 84  # ragged on both sides of the 32 x 32 tiles, padded pitches
    w = rnd("ct.w", (R, Cc))
    out = torch.zeros((R, 88), dtype=torch.bfloat16, device=dev())
    out_t = torch.zeros((Cc, 104), dtype=torch.bfloat16, device=dev())
    call("bsi_cast_transpose_bf16", L.ptr(out), L.ptr(out_t), L.ptr(w), R, Cc, 88, 104, L.stream_ptr())
    sync()
    assert torch.equal(out[:, :Cc], w.bfloat16()) and float(out[:, Cc:].abs().max()) == 0.0
    assert torch.equal(out_t[:, :R], w.bfloat16().t()) and float(out_t[:, R:].abs().max()) == 0.0


@pytest.mark.parametrize("B,T,heads", [(2, 256, 2), (3, 128, 4), (2, 512, 1)])
def test_attention_backward_vs_autograd(B, T, heads):
    hd, dim = 64, heads * 64
    qkv = rnd(f"ab.qkv{T}", (B * T, 3 * dim), 2.0).bfloat16()
    dout = rnd(f"ab.do{T}", (B * T, dim)).bfloat16()
    out = torch.zeros((B * T, dim), dtype=torch.bfloat16, device=dev())
    call("bsi_attention_bf16", L.ptr(out), L.ptr(qkv), B, T, heads, hd, L.stream_ptr())
    dqkv = torch.full((B * T, 3 * dim), float("nan"), dtype=torch.bfloat16, device=dev())
    ws = torch.zeros((2, B * heads * T), device=dev())
    call("bsi_attention_backward_bf16", L.ptr(dqkv), L.ptr(ws[0]), L.ptr(ws[1]), L.ptr(qkv), L.ptr(out), L.ptr(dout), B, T, heads, hd, 0.0, 0, 0, L.stream_ptr())
    sync()
    q, k, v = (t.detach().requires_grad_(True) for t in qkv.float().reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4))
    o = F.scaled_dot_product_attention(q, k, v)
    o.backward(dout.float().reshape(B, T, heads, hd).permute(0, 2, 1, 3))
    ref = torch.stack((q.grad, k.grad, v.grad)).permute(1, 3, 0, 2, 4).reshape(B * T, 3 * dim)
    lse_ref = torch.logsumexp((q @ k.transpose(-1, -2)).detach() / 8.0, dim=-1) * 1.4426950408889634
    report("log-sum-exp (base 2)", ws[0].reshape(B, heads, T), lse_ref, 1e-4, 1e-3)
    for name, sl in (("dq", slice(0, dim)), ("dk", slice(dim, 2 * dim)), ("dv", slice(2 * dim, 3 * dim))):
        got, want = dqkv[:, sl].float(), ref[:, sl]
        rel = float((got - want).norm() / want.norm())
        assert rel < 1.5e-2, f"{name}: relative L2 error {rel}"
        report(name, got, want, 5e-2, 2e-2 * float(want.abs().max()))


def test_layernorm_dropout_forward_backward_with_restated_mask():
    """nn.Dropout on the modulated activations (dit.py:101): the kernels' stateless mask is restated in Python (helpers.dropout_keep)
    and fed to torch autograd as an explicit tensor."""
    B, T, dim, rpc, p, seed = 2, 256, 256, 32, 0.3, 0xC0FFEE
    M = B * T
    x = rnd("ld.x", (M, dim), 1.2)
    tab = rnd("ld.t", (B, 6 * dim), 0.5)
    out = torch.zeros((M, dim), dtype=torch.bfloat16, device=dev())
    call("bsi_layernorm_mod_dropout_bf16", L.ptr(out), L.ptr(x), L.rowref(tab, 6 * dim, 0, 0), L.rowref(tab, 6 * dim, 0, dim), T, M, dim, 1e-5, p, seed,
         L.stream_ptr())
    keep = H.dropout_keep(seed, torch.arange(M * dim, dtype=torch.int64), p).reshape(M, dim).to(dev())
    assert abs(float(keep.float().mean()) - (1 - p)) < 5e-3
    xr = x.clone().requires_grad_(True)
    a = torch.addcmul(tab[:, None, :dim], tab[:, None, dim : 2 * dim] + 1, F.layer_norm(xr, (dim,), eps=1e-5).reshape(B, T, dim)).reshape(M, dim)
    ref = a * keep / (1 - p)
    sync()
    report("ln dropout forward", out, ref, 1e-2, 1e-2)
    assert torch.equal(out == 0, ~keep | (ref.bfloat16() == 0))  # exactly the restated mask
    da = rnd("ld.da", (M, dim)).bfloat16()
    ref.backward(da.float())
    dx = torch.zeros_like(x)
    parts = torch.zeros((2, M // rpc, dim), device=dev())
    call("bsi_layernorm_mod_backward", L.ptr(dx), L.ptr(parts[0]), L.ptr(parts[1]), L.ptr(da), L.ptr(x), L.rowref(tab, 6 * dim, 0, dim), None, T, rpc, M, dim,
         1e-5, p, seed, L.stream_ptr())
    sync()
    report("ln dropout backward dx", dx, xr.grad, 1e-4, 1e-4)


@pytest.mark.parametrize("B,T,heads,p", [(2, 256, 2, 0.05), (1, 128, 3, 0.5)])
def test_attention_dropout_forward_backward_with_restated_mask(B, T, heads, p):
    """Attention dropout (F.scaled_dot_product_attention(dropout_p), dit.py:43-44) forward and backward against torch autograd with
    the kernels' mask restated in Python."""
    hd, dim, seed = 64, heads * 64, 20261017
    qkv = rnd(f"ad.qkv{T}", (B * T, 3 * dim), 2.0).bfloat16()
    dout = rnd(f"ad.do{T}", (B * T, dim)).bfloat16()
    out = torch.zeros((B * T, dim), dtype=torch.bfloat16, device=dev())
    lse = torch.zeros(B * heads * T, device=dev())
    call("bsi_attention_dropout_bf16", L.ptr(out), L.ptr(lse), L.ptr(qkv), B, T, heads, hd, p, seed, L.stream_ptr())
    dqkv = torch.full((B * T, 3 * dim), float("nan"), dtype=torch.bfloat16, device=dev())
    ws = torch.zeros((2, B * heads * T), device=dev())
    call("bsi_attention_backward_bf16", L.ptr(dqkv), L.ptr(ws[0]), L.ptr(ws[1]), L.ptr(qkv), L.ptr(out), L.ptr(dout), B, T, heads, hd, p, seed, 0, L.stream_ptr())
    # fast path: the statistics saved by the forward kernel replace the recomputation pass -- same result up to the rounding of lse
    dqkv_fast = torch.full_like(dqkv, float("nan"))
    call("bsi_attention_backward_bf16", L.ptr(dqkv_fast), L.ptr(lse), L.ptr(ws[1]), L.ptr(qkv), L.ptr(out), L.ptr(dout), B, T, heads, hd, p, seed, 1, L.stream_ptr())
    sync()
    report("saved vs recomputed log-sum-exp", lse, ws[0], 1e-5, 1e-4)
    report("backward with saved statistics", dqkv_fast, dqkv, 2e-2, 4e-3)
    assert float((dqkv_fast.float() - dqkv.float()).norm() / dqkv.float().norm()) < 2e-3
    keep = H.attention_dropout_mask(seed, B, heads, T, p).to(dev())
    assert abs(float(keep.float().mean()) - (1 - p)) < 1e-2
    q, k, v = (t.detach().requires_grad_(True) for t in qkv.float().reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4))
    prob = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1) * keep / (1 - p)
    o = prob @ v
    o.backward(dout.float().reshape(B, T, heads, hd).permute(0, 2, 1, 3))
    o_ref = o.detach().permute(0, 2, 1, 3).reshape(B * T, dim)
    assert float((out.float() - o_ref).norm() / o_ref.norm()) < 1e-2
    ref = torch.stack((q.grad, k.grad, v.grad)).permute(1, 3, 0, 2, 4).reshape(B * T, 3 * dim)
    for name, sl in (("dq", slice(0, dim)), ("dk", slice(dim, 2 * dim)), ("dv", slice(2 * dim, 3 * dim))):
        got, want = dqkv[:, sl].float(), ref[:, sl]
        rel = float((got - want).norm() / want.norm())
        assert rel < 2e-2, f"{name}: relative L2 error {rel}"


@pytest.mark.parametrize("dim", [128, 1024])
def test_gate_residual_layernorm_fused_equals_the_two_kernels(dim):
    """x_out = x + gate * branch and LayerNorm(+modulation / affine, + dropout) of x_out in one pass == the two separate kernels, bit for bit."""
    B, T = 3, 256
    M = B * T
    x = rnd(f"gl.x{dim}", (M, dim), 1.3)
    br = rnd(f"gl.br{dim}", (M, dim)).bfloat16()
    tab = rnd(f"gl.t{dim}", (B, 6 * dim), 0.5)
    gate, shift, scale = (L.rowref(tab, 6 * dim, 0, j * dim) for j in (2, 3, 4))
    none = L.RowRef(None, 0, 0)
    gamma, beta = rnd(f"gl.g{dim}", (dim,), 0.3) + 1, rnd(f"gl.b{dim}", (dim,), 0.2)
    for variant in ("modulated", "dropout", "affine"):
        p, seed = (0.2, 77) if variant == "dropout" else (0.0, 0)
        x_sep = torch.zeros_like(x)
        call("bsi_gate_residual", L.ptr(x_sep), L.ptr(x), L.ptr(br), gate, T, M, dim, L.stream_ptr())
        a_sep = torch.zeros((M, dim), dtype=torch.bfloat16, device=dev())
        if variant == "affine":
            call("bsi_layernorm_mod_bf16", L.ptr(a_sep), L.ptr(x_sep), none, none, None, L.ptr(gamma), L.ptr(beta), T, M, dim, 1e-5, L.stream_ptr())
        elif variant == "dropout":
            call("bsi_layernorm_mod_dropout_bf16", L.ptr(a_sep), L.ptr(x_sep), shift, scale, T, M, dim, 1e-5, p, seed, L.stream_ptr())
        else:
            call("bsi_layernorm_mod_bf16", L.ptr(a_sep), L.ptr(x_sep), shift, scale, None, None, None, T, M, dim, 1e-5, L.stream_ptr())
        x_f, a_f = torch.zeros_like(x), torch.zeros_like(a_sep)
        aff = variant == "affine"
        call("bsi_gate_residual_layernorm_bf16", L.ptr(a_f), L.ptr(x_f), L.ptr(x), L.ptr(br), gate, none if aff else shift, none if aff else scale,
             L.ptr(gamma) if aff else None, L.ptr(beta) if aff else None, T, M, dim, 1e-5, p, seed, L.stream_ptr())
        sync()
        assert torch.equal(x_f, x_sep) and torch.equal(a_f, a_sep), variant


@pytest.mark.parametrize("T", [256, 128], ids=["tcgen05", "mma_sync"])
def test_attention_forward_with_saved_statistics(T):
    """bsi_attention_lse_bf16 == bsi_attention_bf16 bit for bit, plus the log2-sum-exp the backward's fast path consumes."""
    B, heads, hd = 3, 4, 64
    dim = heads * hd
    qkv = rnd(f"al.qkv{T}", (B * T, 3 * dim), 2.0).bfloat16()
    dout = rnd(f"al.do{T}", (B * T, dim)).bfloat16()
    out0 = torch.zeros((B * T, dim), dtype=torch.bfloat16, device=dev())
    out1 = torch.zeros_like(out0)
    lse = torch.zeros(B * heads * T, device=dev())
    call("bsi_attention_bf16", L.ptr(out0), L.ptr(qkv), B, T, heads, hd, L.stream_ptr())
    call("bsi_attention_lse_bf16", L.ptr(out1), L.ptr(lse), L.ptr(qkv), B, T, heads, hd, L.stream_ptr())
    sync()
    assert torch.equal(out0, out1)
    q, k, _ = qkv.float().reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
    ref = torch.logsumexp(q @ k.transpose(-1, -2) / 8.0, dim=-1) * 1.4426950408889634
    report("saved log2-sum-exp", lse.reshape(B, heads, T), ref, 1e-4, 1e-3)
    ws = torch.zeros((2, B * heads * T), device=dev())
    slow, fast = torch.zeros_like(qkv), torch.zeros_like(qkv)
    call("bsi_attention_backward_bf16", L.ptr(slow), L.ptr(ws[0]), L.ptr(ws[1]), L.ptr(qkv), L.ptr(out0), L.ptr(dout), B, T, heads, hd, 0.0, 0, 0, L.stream_ptr())
    call("bsi_attention_backward_bf16", L.ptr(fast), L.ptr(lse), L.ptr(ws[1]), L.ptr(qkv), L.ptr(out0), L.ptr(dout), B, T, heads, hd, 0.0, 0, 1, L.stream_ptr())
    sync()
    # the saved and the recomputed statistics differ in the last bits, which moves a bf16 gradient by at most one ulp of the tensor's range
    report("backward with saved statistics", fast, slow, 2e-2, 4e-3)
    assert float((fast.float() - slow.float()).norm() / slow.float().norm()) < 2e-3


def test_shared_memory_opt_in_is_never_lowered_by_another_entry_point():
    """Regression: bsi_attention_lse_bf16 (T = 128, 48 KB) and the mma.sync path of bsi_attention_bf16 (T = 512, 144 KB) launch the same
    kernel from two call sites; the smaller request must not lower the kernel's dynamic shared memory limit again."""
    heads, hd = 2, 64
    dim = heads * hd
    call("bsi_attention_force_legacy", 1)
    try:
        for T in (512, 128, 512):
            qkv = rnd(f"sm.qkv{T}", (T, 3 * dim)).bfloat16()
            out = torch.zeros((T, dim), dtype=torch.bfloat16, device=dev())
            lse = torch.zeros(heads * T, device=dev())
            call("bsi_attention_bf16", L.ptr(out), L.ptr(qkv), 1, T, heads, hd, L.stream_ptr())
            call("bsi_attention_lse_bf16", L.ptr(out), L.ptr(lse), L.ptr(qkv), 1, T, heads, hd, L.stream_ptr())
            sync()
            assert torch.isfinite(out.float()).all()
    finally:
        call("bsi_attention_force_legacy", 0)


@pytest.mark.parametrize("B,heads,p", [(2, 2, 0.0), (10, 16, 0.0), (3, 4, 0.1)], ids=["small", "two_items_per_cta", "dropout"])
def test_attention_backward_tcgen05_vs_autograd(B, heads, p):
    """The tcgen05 backward (T = 256, statistics saved by the forward; attention_bwd_sm100.cu) against torch autograd, with the
    dropout mask restated in Python; 160 (sample, head) items make some CTAs loop over two items."""
    T, hd, dim, seed = 256, 64, heads * 64, 20261017
    qkv = rnd(f"tb.qkv{B}", (B * T, 3 * dim), 2.0).bfloat16()
    dout = rnd(f"tb.do{B}", (B * T, dim)).bfloat16()
    out = torch.zeros((B * T, dim), dtype=torch.bfloat16, device=dev())
    lse = torch.zeros(B * heads * T, device=dev())
    if p > 0:
        call("bsi_attention_dropout_bf16", L.ptr(out), L.ptr(lse), L.ptr(qkv), B, T, heads, hd, p, seed, L.stream_ptr())
    else:
        call("bsi_attention_lse_bf16", L.ptr(out), L.ptr(lse), L.ptr(qkv), B, T, heads, hd, L.stream_ptr())
    dqkv = torch.full((B * T, 3 * dim), float("nan"), dtype=torch.bfloat16, device=dev())
    dsum = torch.zeros(B * heads * T, device=dev())
    call("bsi_attention_backward_bf16", L.ptr(dqkv), L.ptr(lse), L.ptr(dsum), L.ptr(qkv), L.ptr(out), L.ptr(dout), B, T, heads, hd, p, seed, 1, L.stream_ptr())
    sync()
    q, k, v = (t.detach().requires_grad_(True) for t in qkv.float().reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4))
    prob = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    if p > 0:
        prob = prob * H.attention_dropout_mask(seed, B, heads, T, p).to(dev()) / (1 - p)
    o = prob @ v
    o.backward(dout.float().reshape(B, T, heads, hd).permute(0, 2, 1, 3))
    ref = torch.stack((q.grad, k.grad, v.grad)).permute(1, 3, 0, 2, 4).reshape(B * T, 3 * dim)
    d_ref = (dout.float() * out.float()).reshape(B, T, heads, hd).sum(-1).permute(0, 2, 1)
    report("D = rowsum(dO o O)", dsum.reshape(B, heads, T), d_ref, 1e-5, 1e-5)
    assert torch.isfinite(dqkv.float()).all()
    for name, sl in (("dq", slice(0, dim)), ("dk", slice(dim, 2 * dim)), ("dv", slice(2 * dim, 3 * dim))):
        got, want = dqkv[:, sl].float(), ref[:, sl]
        rel = float((got - want).norm() / want.norm())
        assert rel < 1.5e-2, f"{name}: relative L2 error {rel}"
        report(name, got, want, 5e-2, 2e-2 * float(want.abs().max()))
