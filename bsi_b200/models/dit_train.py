"""Differentiable forward of the native DenoisingDiT (reference bsi/models/dit.py:87-103,174-181,225-233) for
``BSI.train_loss(x).mean().backward()`` -- SURVEY §8 a23, first version.

The inference engine (dit_engine.cu) fuses each block into a handful of kernels and keeps nothing; training needs the
intermediates, so this module drives the same C-ABI primitives one by one from Python, saves the bf16 operands of every
GEMM, and runs the backward with

  * ``bsi_gemm_bf16``        forward GEMMs, and the data gradients dX = dY W through transposed bf16 weight copies,
  * ``bsi_gemm_wgrad_bf16``  weight gradients dW += dY^T X (tcgen05, MN-major operands, split-M reduce-add),
  * ``bsi_layernorm_mod_bf16`` / ``bsi_layernorm_mod_backward``, ``bsi_gate_residual(_backward)``, ``bsi_gelu(_backward)_bf16``,
  * ``bsi_attention_bf16``   forward attention,

  * ``bsi_attention_backward_bf16``  attention backward (two deterministic mma.sync kernels that recompute the scores).

What is still PyTorch here, and why: the adaLN / time-embedding chain on ``[B, 6*dim]`` tensors (O(batch) work, 0.2 % of the
flops; differentiated by autograd through the ``mods`` argument, under bf16 autocast) and the final sums of per-CTA partial
reductions (``[B, chunks, D]`` buffers).  Every pass over a token-sized tensor runs on this repo's kernels.
Dropout (the reference's experiments train with 0.05) is fused into the attention and LayerNorm kernels through a stateless hash
mask (csrc/common.cuh ``dropout_keep``), which the backward kernels regenerate.
"""

from __future__ import annotations

import ctypes as C
import os

import torch
from torch import Tensor

from .. import _lib as L

_LN_ROWS_PER_CTA = 32
# A/B switch for measurements: 0 keeps GELU forward / backward as separate elementwise kernels (the round-1 path)
_FUSED_GELU = os.environ.get("BSI_TRAIN_FUSED_GELU", "1") != "0"


def _st(dev):
    return L.stream_ptr(dev)


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _pad_cols(t: Tensor, n: int) -> Tensor:
    """[rows][c] -> contiguous [rows][n] with zero columns appended."""
    if t.shape[1] == n:
        return t.contiguous()
    out = t.new_zeros((t.shape[0], n))
    out[:, : t.shape[1]] = t
    return out


def _pad_rows(t: Tensor, n: int) -> Tensor:
    if t.shape[0] == n:
        return t.contiguous()
    out = t.new_zeros((n, *t.shape[1:]))
    out[: t.shape[0]] = t
    return out


def _gemm(A: Tensor, W: Tensor, out: Tensor, bias: Tensor, epi: int, *, pos: Tensor | None = None, rows_per_sample: int = 0, aux: Tensor | None = None) -> None:
    """out[M][N] = epilogue(A[M][K] @ W[N][K]^T + bias); A, W bf16 row-major (pitch = last stride-1 dimension).
    aux: bf16 [M][N] second output (EPI_BIAS_GELU_DUAL_BF16) or saved pre-activation input (EPI_MUL_GELU_GRAD_BF16)."""
    a = L.GemmArgs()
    a.A, a.W, a.C, a.bias = A.data_ptr(), W.data_ptr(), out.data_ptr(), bias.data_ptr()
    a.M, a.N, a.K = A.shape[0], W.shape[0], A.shape[1]
    a.lda, a.ldw, a.ldc = A.stride(0), W.stride(0), out.stride(0)
    a.batch = 1
    a.stride_a, a.stride_w, a.stride_c, a.stride_bias = a.M * a.lda, a.N * a.ldw, a.M * a.ldc, a.N
    a.epilogue = epi
    a.gate = L.RowRef(None, 0, 0)
    a.step_ptr, a.rows_per_sample = None, rows_per_sample
    a.pos = pos.data_ptr() if pos is not None else None
    a.patch = a.grid_w = a.channels = 0
    a.aux = aux.data_ptr() if aux is not None else None
    L.check(L.load().bsi_gemm_bf16(C.byref(a), _st(A.device)), "bsi_gemm_bf16")


def _wgrad(dY: Tensor, X: Tensor, into: Tensor | None = None) -> Tensor:
    """dW[N][K] (+)= dY[M][N]^T @ X[M][K] (fp32): a fresh tensor, or accumulated in place into `into` (a gradient-arena slice)."""
    M, N, K = dY.shape[0], dY.shape[1], X.shape[1]
    dW = into if into is not None else torch.zeros((N, K), dtype=torch.float32, device=dY.device)
    L.check(L.load().bsi_gemm_wgrad_bf16(dW.data_ptr(), dY.data_ptr(), X.data_ptr(), M, N, K, dY.stride(0), X.stride(0), K, 0, _st(dY.device)),
            "bsi_gemm_wgrad_bf16")
    return dW


def _ln_mod(x: Tensor, shift: L.RowRef | None, scale: L.RowRef | None, T: int, gamma: Tensor | None = None, beta: Tensor | None = None,
            drop: tuple[float, int] = (0.0, 0)) -> Tensor:
    M, D = x.shape
    out = torch.empty((M, D), dtype=torch.bfloat16, device=x.device)
    none = L.RowRef(None, 0, 0)
    if drop[0] > 0:
        L.check(L.load().bsi_layernorm_mod_dropout_bf16(out.data_ptr(), x.data_ptr(), shift, scale, T, M, D, 1e-5, drop[0], drop[1], _st(x.device)),
                "bsi_layernorm_mod_dropout_bf16")
        return out
    L.check(L.load().bsi_layernorm_mod_bf16(out.data_ptr(), x.data_ptr(), shift or none, scale or none, None, L.ptr(gamma), L.ptr(beta), T, M, D, 1e-5,
                                            _st(x.device)), "bsi_layernorm_mod_bf16")
    return out


def _ln_mod_backward(dx_io: Tensor, da: Tensor, x: Tensor, scale: L.RowRef | None, T: int, gamma: Tensor | None = None,
                     drop: tuple[float, int] = (0.0, 0), shift_first: bool = False):
    """dx_io += dL/dx; returns the per-CTA partial sums [2 (dscale, dshift)][M / 32][D] -- (dshift, dscale), the order of the
    modulation vector's chunks, with ``shift_first``."""
    M, D = x.shape
    parts = torch.empty((2, (M + _LN_ROWS_PER_CTA - 1) // _LN_ROWS_PER_CTA, D), dtype=torch.float32, device=x.device)
    i_scale, i_shift = (1, 0) if shift_first else (0, 1)
    L.check(L.load().bsi_layernorm_mod_backward(dx_io.data_ptr(), parts[i_scale].data_ptr(), parts[i_shift].data_ptr(), da.data_ptr(), x.data_ptr(),
                                                scale or L.RowRef(None, 0, 0), L.ptr(gamma), T, _LN_ROWS_PER_CTA, M, D, 1e-5, drop[0], drop[1],
                                                _st(x.device)),
            "bsi_layernorm_mod_backward")
    return parts


class _ReduceBatch:
    """Collects partial-sum reductions (csrc/train_kernels.cu k_reduce_rows) and runs them in one launch: dst[g] (+)= sum_r src[g * rows + r].
    The partial buffers are kept alive until ``flush``."""

    MAX = 12

    def __init__(self, dev):
        self.dev, self.jobs, self.keep = dev, [], []

    def add(self, src: Tensor, dst: Tensor, groups: int, rows: int, accumulate: bool = False) -> None:
        D = src.shape[-1]
        assert src.is_contiguous() and src.dtype == torch.float32 and dst.dtype == torch.float32 and dst.stride(-1) == 1 and src.numel() == groups * rows * D
        ld = dst.stride(0) if dst.dim() > 1 else D
        self.jobs.append(L.ReduceJob(src.data_ptr(), dst.data_ptr(), groups, rows, D, ld, int(accumulate), 0))
        self.keep += [src, dst]
        if len(self.jobs) == self.MAX:
            self.flush()

    def flush(self) -> None:
        if self.jobs:
            self.jobs.sort(key=lambda j: -j.rows)  # the long (bias-gradient) reductions first: their CTAs are the kernel's critical path
            arr = (L.ReduceJob * len(self.jobs))(*self.jobs)
            L.check(L.load().bsi_reduce_rows(arr, len(self.jobs), _st(self.dev)), "bsi_reduce_rows")
        self.jobs, self.keep = [], []


class DiTTrainFunction(torch.autograd.Function):
    """out = DiT(in_scale * mu; mods, params).  Differentiable w.r.t. ``mods`` [L][B][6*dim] and the parameter list (not mu)."""

    @staticmethod
    def forward(ctx, model, mu: Tensor, in_scale: Tensor | None, drop: tuple[float, int], mods: Tensor, *params: Tensor):
        cfg = model._cfg
        dev = mu.device
        B, Cc, H, Wd = mu.shape
        p, D, depth, heads = cfg.patch, cfg.dim, cfg.depth, cfg.heads
        T = (H // p) * (Wd // p)
        M = B * T
        lib = L.load()
        it = iter(params)
        w_patch, b_patch = next(it), next(it)
        blocks = [tuple(next(it) for _ in range(8)) for _ in range(depth)]
        ln_g, ln_b, w_dec, b_dec = next(it), next(it), next(it), next(it)
        wts: list[Tensor] = []  # transposed bf16 weights, in parameter order, for the data-gradient GEMMs of the backward

        # bf16 / transposed bf16 copies of the weights are kept between calls while the weights do not change (micro-batches of one
        # optimisation step): keyed like NativeDenoiser._signature -- address, torch version counter, generation of the optimizer arena
        cache = model.__dict__.setdefault("_train_wcache", {})

        def stamp(w: Tensor):
            arena = getattr(w, "_bsi_arena", None)
            return (w.data_ptr(), w._version, arena[0].generation if arena is not None else 0)

        def bf(w: Tensor, pad_rows: int = 0, pad_cols: int = 0) -> Tensor:
            """bf16 copy [N'][K'] (zero padded) for the forward GEMM; its transpose [K'][N'] is produced in the same pass."""
            key, st_w = (id(w), pad_rows, pad_cols), stamp(w)
            hit = cache.get(key)
            if hit is not None and hit[0] == st_w:
                wts.append(hit[2])
                return hit[1]
            n, k = w.shape
            n_p, k_p = max(n, pad_rows), max(k, pad_cols)
            alloc = torch.zeros if (n_p != n or k_p != k) else torch.empty
            w16, wt16 = alloc((n_p, k_p), dtype=torch.bfloat16, device=dev), alloc((k_p, n_p), dtype=torch.bfloat16, device=dev)
            L.check(lib.bsi_cast_transpose_bf16(w16.data_ptr(), wt16.data_ptr(), w.detach().float().contiguous().data_ptr(), n, k, k_p, n_p, _st(dev)),
                    "bsi_cast_transpose_bf16")
            cache[key] = (st_w, w16, wt16)
            wts.append(wt16)
            return w16

        drop_p, drop_seed = drop
        with torch.cuda.device(dev):
            scale = torch.ones(1, dtype=torch.float32, device=dev) if in_scale is None else in_scale.detach().float().contiguous()
            K0 = _pad8(w_patch.shape[1])  # operand pitches are multiples of 16 bytes; the padding columns are zero on both sides
            a0 = torch.empty((M, K0), dtype=torch.bfloat16, device=dev)
            L.check(lib.bsi_dit_patch_operand(a0.data_ptr(), mu.detach().float().contiguous().data_ptr(), L.rowref(scale, 0 if in_scale is None else 1), None,
                                              B, Cc, H, Wd, p, cfg.fourier_n_min, cfg.fourier_n_max, K0, _st(dev)), "bsi_dit_patch_operand")
            x = torch.empty((M, D), dtype=torch.float32, device=dev)
            _gemm(a0, bf(w_patch, pad_cols=K0), x, b_patch.detach().float(), L.EPI_POS_F32, pos=model.dit.patch_pos_embedding.float().contiguous(),
                  rows_per_sample=T)
            ctx.mods_dtype = mods.dtype
            mods = mods.detach().float().contiguous()
            saved = []
            none = L.RowRef(None, 0, 0)

            def gate_ln(x_in: Tensor, br: Tensor, gate, shift, scale, gamma=None, beta=None, drop_site=(0.0, 0)):
                """x_out = x_in + gate * br and the LayerNorm (+ modulation / affine, + dropout) of x_out, one pass (train_kernels.cu)."""
                x_out = torch.empty_like(x_in)  # out of place: x_in stays alive as a saved activation
                a = torch.empty((M, D), dtype=torch.bfloat16, device=dev)
                L.check(lib.bsi_gate_residual_layernorm_bf16(a.data_ptr(), x_out.data_ptr(), x_in.data_ptr(), br.data_ptr(), gate, shift or none, scale or none,
                                                             L.ptr(gamma), L.ptr(beta), T, M, D, 1e-5, drop_site[0], drop_site[1], _st(dev)),
                        "bsi_gate_residual_layernorm_bf16")
                return x_out, a

            a1 = _ln_mod(x, L.rowref(mods[0], 6 * D, 0, 0), L.rowref(mods[0], 6 * D, 0, D), T)
            for l, (w_qkv, b_qkv, w_o, b_o, w_1, b_1, w_2, b_2) in enumerate(blocks):
                m = mods[l]
                ref = lambda j: L.rowref(m, 6 * D, 0, j * D)
                x_in = x
                qkv = torch.empty((M, 3 * D), dtype=torch.bfloat16, device=dev)
                _gemm(a1, bf(w_qkv), qkv, b_qkv.detach().float(), L.EPI_BIAS_BF16)
                att = torch.empty((M, D), dtype=torch.bfloat16, device=dev)
                lse = torch.empty(B * heads * T, dtype=torch.float32, device=dev)  # softmax statistics: spare the backward a recomputation pass
                if drop_p > 0:
                    L.check(lib.bsi_attention_dropout_bf16(att.data_ptr(), lse.data_ptr(), qkv.data_ptr(), B, T, heads, D // heads, drop_p,
                                                           _layer_seed(drop_seed, 2 * l), _st(dev)), "bsi_attention_dropout_bf16")
                else:
                    L.check(lib.bsi_attention_lse_bf16(att.data_ptr(), lse.data_ptr(), qkv.data_ptr(), B, T, heads, D // heads, _st(dev)),
                            "bsi_attention_lse_bf16")
                br1 = torch.empty((M, D), dtype=torch.bfloat16, device=dev)
                _gemm(att, bf(w_o), br1, b_o.detach().float(), L.EPI_BIAS_BF16)
                x_mid, a2 = gate_ln(x, br1, ref(2), ref(3), ref(4), drop_site=(drop_p, _layer_seed(drop_seed, 2 * l + 1)))
                pre = torch.empty((M, 4 * D), dtype=torch.bfloat16, device=dev)
                h = torch.empty_like(pre)
                if _FUSED_GELU and M > 128:  # one GEMM writes the pre-activation (kept for the backward) and its GELU from the same accumulator tile
                    _gemm(a2, bf(w_1), h, b_1.detach().float(), L.EPI_BIAS_GELU_DUAL_BF16, aux=pre)
                else:
                    _gemm(a2, bf(w_1), pre, b_1.detach().float(), L.EPI_BIAS_BF16)
                    L.check(lib.bsi_gelu_bf16(h.data_ptr(), pre.data_ptr(), pre.numel(), _st(dev)), "bsi_gelu_bf16")
                br2 = torch.empty((M, D), dtype=torch.bfloat16, device=dev)
                _gemm(h, bf(w_2), br2, b_2.detach().float(), L.EPI_BIAS_BF16)
                saved.append((x_in, a1, qkv, att, br1, x_mid, a2, pre, h, br2, lse))
                if l + 1 < depth:  # the MLP branch's residual update opens the next block's attention LayerNorm ...
                    nxt = mods[l + 1]
                    x, a1 = gate_ln(x_mid, br2, ref(5), L.rowref(nxt, 6 * D, 0, 0), L.rowref(nxt, 6 * D, 0, D))
                else:  # ... or the affine LayerNorm of the patch decoder
                    x, a_dec = gate_ln(x_mid, br2, ref(5), None, None, ln_g.detach().float().contiguous(), ln_b.detach().float().contiguous())
            n_out = w_dec.shape[0]
            Np = _pad8(n_out)
            y = torch.empty((M, Np), dtype=torch.float32, device=dev)
            _gemm(a_dec, bf(w_dec, pad_rows=Np), y, _pad_rows(b_dec.detach().float(), Np), L.EPI_BIAS_F32)
            gh, gw = H // p, Wd // p
            out = y[:, :n_out].reshape(B, gh, gw, p, p, Cc).permute(0, 5, 1, 3, 2, 4).reshape(B, Cc, H, Wd).contiguous()
        ctx.model, ctx.geom, ctx.drop = model, (B, Cc, H, Wd, T, M), drop
        ctx.saved_acts = (a0, x, a_dec, saved, wts)
        ctx.save_for_backward(mods, *params)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout: Tensor):
        model = ctx.model
        cfg = model._cfg
        B, Cc, H, Wd, T, M = ctx.geom
        p, D, depth, heads = cfg.patch, cfg.dim, cfg.depth, cfg.heads
        mods, *params = ctx.saved_tensors
        drop_p, drop_seed = ctx.drop
        a0, x_last, a_dec, saved, wts = ctx.saved_acts
        wt_blocks = [wts[1 + 4 * l : 5 + 4 * l] for l in range(depth)]  # [patch | (qkv, out, mlp1, mlp2) per block | decoder]
        wt_dec = wts[-1]
        dev = dout.device
        lib = L.load()
        it = iter(params)
        w_patch, b_patch = next(it), next(it)
        blocks = [tuple(next(it) for _ in range(8)) for _ in range(depth)]
        ln_g, ln_b, w_dec, b_dec = next(it), next(it), next(it), next(it)
        red = _ReduceBatch(dev)  # the partial-sum reductions of a block run as one launch at the end of the block

        def colsum_parts(t: Tensor) -> Tensor:
            """per-CTA partial column sums [ceil(M / 256)][N] of a bf16 matrix"""
            parts = torch.empty(((t.shape[0] + 255) // 256, t.shape[1]), dtype=torch.float32, device=dev)
            L.check(lib.bsi_colsum_bf16(parts.data_ptr(), t.data_ptr(), t.shape[0], t.shape[1], t.stride(0), 256, _st(dev)), "bsi_colsum_bf16")
            return parts

        def colsum(t: Tensor) -> Tensor:
            if t.dtype != torch.bfloat16:
                return t.sum(0, dtype=torch.float32)
            return colsum_parts(t).sum(0)

        zero_bias = torch.zeros(4 * D, dtype=torch.float32, device=dev)  # the data-gradient GEMMs have no bias: one shared zero vector
        zeros = lambda n: zero_bias[:n]
        grads: list[Tensor] = []
        sink = getattr(model, "_grad_sink", None)

        def emit_w(p: Tensor, dY: Tensor, X: Tensor, rows: int | None = None, cols: int | None = None):
            """Weight gradient of `p`: accumulated by the GEMM straight into the optimizer's arena when one is attached (returns
            None: autograd has nothing left to add), else returned as a tensor."""
            view = sink.grad_view(p) if sink is not None else None
            if view is not None and rows is None and cols is None:
                _wgrad(dY, X, into=view)
                return None
            g = _wgrad(dY, X)
            g = g[:rows] if rows is not None else g
            g = g[:, :cols] if cols is not None else g
            if view is not None:
                view.add_(g)
                return None
            return g

        def emit_b(p: Tensor, g: Tensor):
            view = sink.grad_view(p) if sink is not None else None
            if view is not None:
                view.add_(g)
                return None
            return g

        def emit_b_parts(p: Tensor, parts: Tensor):
            """Bias gradient = sum of all partial rows: accumulated into the optimizer's arena, or into a fresh tensor for autograd."""
            view = sink.grad_view(p) if sink is not None else None
            if parts.shape[-1] % 4 or (view is not None and view.data_ptr() % 16):
                return emit_b(p, parts.sum(0))
            if view is not None:
                red.add(parts, view, 1, parts.shape[0], accumulate=True)
                return None
            g = torch.empty(parts.shape[-1], dtype=torch.float32, device=dev)
            red.add(parts, g, 1, parts.shape[0])
            return g

        with torch.cuda.device(dev):
            gh, gw = H // p, Wd // p
            dy = dout.float().reshape(B, Cc, gh, p, gw, p).permute(0, 2, 4, 3, 5, 1).reshape(M, p * p * Cc)
            n_out = w_dec.shape[0]
            Np = _pad8(n_out)
            dy16 = _pad_cols(dy.to(torch.bfloat16), Np)
            g_wdec, g_bdec = emit_w(w_dec, dy16, a_dec, rows=n_out), emit_b(b_dec, colsum(dy))
            da = torch.empty((M, D), dtype=torch.bfloat16, device=dev)
            _gemm(dy16, wt_dec, da, zeros(D), L.EPI_BIAS_BF16)  # W^T [D][Np]: the forward kernel computes dY @ W
            dx = torch.zeros((M, D), dtype=torch.float32, device=dev)
            dgb = _ln_mod_backward(dx, da, x_last, None, T, ln_g.detach().float().contiguous()).sum(1)  # [2][D]: dgamma, dbeta
            tail = [emit_b(ln_g, dgb[0]), emit_b(ln_b, dgb[1]), g_wdec, g_bdec]
            if sink is not None:
                sink.grads_ready([ln_g, ln_b, w_dec, b_dec])
            # d(mods) is collected chunk by chunk, [layer][chunk (shift, scale, gate) x (msa, mlp)][B][D]: the kernels and the partial-sum
            # reductions write their [B][D] results in place, one permuting copy at the end builds [layer][B][6 D]
            dparts = torch.empty((depth, 6, B, D), dtype=torch.float32, device=dev)
            block_grads = []
            rows_ok = T % _LN_ROWS_PER_CTA == 0 and D % 128 == 0 and D <= 1024

            def gate_backward(dbr: Tensor, dgate_out: Tensor, br: Tensor, gate) -> Tensor:
                """dbr = gate * dx, dgate_out[B][D] = sum_t dx * br; returns partial column sums of dbr (their sum over dim 0 is the bias gradient)."""
                if not rows_ok:
                    dbias = torch.empty((B, D), dtype=torch.float32, device=dev)
                    L.check(lib.bsi_gate_residual_backward(dbr.data_ptr(), dgate_out.data_ptr(), dbias.data_ptr(), dx.data_ptr(), br.data_ptr(), gate, T, B, D,
                                                           _st(dev)), "bsi_gate_residual_backward")
                    return dbias
                parts = torch.empty((2, M // _LN_ROWS_PER_CTA, D), dtype=torch.float32, device=dev)
                L.check(lib.bsi_gate_residual_backward_rows(dbr.data_ptr(), parts[0].data_ptr(), parts[1].data_ptr(), dx.data_ptr(), br.data_ptr(), gate, T,
                                                            _LN_ROWS_PER_CTA, M, D, _st(dev)), "bsi_gate_residual_backward_rows")
                red.add(parts[0], dgate_out, B, T // _LN_ROWS_PER_CTA)
                return parts[1]

            for l in reversed(range(depth)):
                w_qkv, b_qkv, w_o, b_o, w_1, b_1, w_2, b_2 = blocks[l]
                x_in, a1, qkv, att, br1, x_mid, a2, pre, h, br2, lse = saved[l]
                wt_qkv, wt_o, wt_1, wt_2 = wt_blocks[l]
                m, dm = mods[l], dparts[l]
                ref = lambda j: L.rowref(m, 6 * D, 0, j * D)
                # per-sample (dshift, dscale) from the CTA partials, written into chunks [j, j + 2) of d(mods)
                part = lambda t, j: red.add(t, dm[j : j + 2].view(2 * B, D), 2 * B, T // _LN_ROWS_PER_CTA)
                # ---- MLP branch: x_out = x_mid + gate_mlp * (gelu(a2 W1^T + b1) W2^T + b2)
                dbr = torch.empty((M, D), dtype=torch.bfloat16, device=dev)
                dbias = gate_backward(dbr, dm[5], br2, ref(5))
                g_w2, g_b2 = emit_w(w_2, dbr, h), emit_b_parts(b_2, dbias)
                dh = torch.empty((M, 4 * D), dtype=torch.bfloat16, device=dev)
                if _FUSED_GELU and M > 128:  # d(pre) = (dbr W2) * gelu'(pre): the derivative is applied in the data-gradient GEMM's epilogue
                    _gemm(dbr, wt_2, dh, zeros(4 * D), L.EPI_MUL_GELU_GRAD_BF16, aux=pre)
                else:
                    _gemm(dbr, wt_2, dh, zeros(4 * D), L.EPI_BIAS_BF16)
                    L.check(lib.bsi_gelu_backward_bf16(dh.data_ptr(), dh.data_ptr(), pre.data_ptr(), dh.numel(), _st(dev)), "bsi_gelu_backward_bf16")
                g_w1, g_b1 = emit_w(w_1, dh, a2), emit_b_parts(b_1, colsum_parts(dh))
                _gemm(dh, wt_1, da, zeros(D), L.EPI_BIAS_BF16)
                part(_ln_mod_backward(dx, da, x_mid, ref(4), T, drop=(drop_p, _layer_seed(drop_seed, 2 * l + 1)), shift_first=True), 3)
                # ---- attention branch: x_mid = x_in + gate_msa * (attn(a1 Wqkv^T + b) Wo^T + b)
                dbias = gate_backward(dbr, dm[2], br1, ref(2))
                g_wo, g_bo = emit_w(w_o, dbr, att), emit_b_parts(b_o, dbias)
                datt = torch.empty((M, D), dtype=torch.bfloat16, device=dev)
                _gemm(dbr, wt_o, datt, zeros(D), L.EPI_BIAS_BF16)
                dqkv = _attention_backward(qkv, att, datt, B, T, heads, D // heads, (drop_p, _layer_seed(drop_seed, 2 * l)), lse)
                g_wqkv, g_bqkv = emit_w(w_qkv, dqkv, a1), emit_b_parts(b_qkv, colsum_parts(dqkv))
                _gemm(dqkv, wt_qkv, da, zeros(D), L.EPI_BIAS_BF16)
                part(_ln_mod_backward(dx, da, x_in, ref(1), T, shift_first=True), 0)
                block_grads.append([g_wqkv, g_bqkv, g_wo, g_bo, g_w1, g_b1, g_w2, g_b2])
                red.flush()
                if sink is not None:  # this block's eight tensors are final: their all-reduce can overlap the remaining layers
                    sink.grads_ready([w_qkv, b_qkv, w_o, b_o, w_1, b_1, w_2, b_2])
                saved[l] = None  # release this layer's activations
            red.flush()
            dx16 = dx.to(torch.bfloat16)
            grads = [emit_w(w_patch, dx16, a0, cols=w_patch.shape[1]), emit_b(b_patch, colsum(dx))]
            if sink is not None:
                sink.grads_ready([w_patch, b_patch])
            for bg in reversed(block_grads):
                grads += bg
            grads += tail
            dmods = dparts.permute(0, 2, 1, 3).reshape(depth, B, 6 * D).to(ctx.mods_dtype)
        return (None, None, None, None, dmods, *grads)


class _AdaLNChain(torch.autograd.Function):
    """mods[l] = Linear_2^l(SiLU(Linear_0^l(cond))) for all blocks at once (dit.py:79-81,90-92) on bf16 tensor cores, like the inference engine's
    conditioning chain (and the reference under bf16 autocast): two batched matmuls over the stacked adaLN weights instead of 24 x
    (Linear, SiLU, Linear).  The backward is written out because autograd's generic path costs more than the chain itself: it materialises
    the stacked weight gradients in fp32 (0.6 GB for DiT-L) and then adds 96 slices into the parameters' gradients.  Here the weight
    gradients dW^l = dY^l^T X^l go through the weight-gradient GEMM straight into the optimizer's arena when one is attached
    (TMA reduce-add is the "+="), the bias gradients through one multi-tensor add."""

    @staticmethod
    def forward(ctx, model, cond: Tensor, *ada: Tensor):
        nl = len(ada) // 4
        w0 = torch.stack([ada[4 * l] for l in range(nl)]).to(torch.bfloat16)      # [L, d, d]
        b0 = torch.stack([ada[4 * l + 1] for l in range(nl)]).to(torch.bfloat16)  # [L, d]
        w2 = torch.stack([ada[4 * l + 2] for l in range(nl)]).to(torch.bfloat16)  # [L, 6d, d]
        b2 = torch.stack([ada[4 * l + 3] for l in range(nl)]).to(torch.bfloat16)  # [L, 6d]
        c16 = cond.to(torch.bfloat16).contiguous()
        pre = torch.baddbmm(b0[:, None, :], c16[None].expand(nl, -1, -1), w0.transpose(1, 2))  # [L, B, d]
        h = torch.nn.functional.silu(pre)
        mods = torch.baddbmm(b2[:, None, :], h, w2.transpose(1, 2))  # [L, B, 6d]
        ctx.model, ctx.nl = model, nl
        ctx.params = ada
        ctx.save_for_backward(c16, pre, h, w0, w2)
        return mods

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dmods: Tensor):
        c16, pre, h, w0, w2 = ctx.saved_tensors
        nl, ada = ctx.nl, ctx.params
        sink = getattr(ctx.model, "_grad_sink", None)
        dm16 = dmods.to(torch.bfloat16).contiguous()
        db2 = dmods.sum(1, dtype=torch.float32)                 # [L, 6d]
        dh = torch.bmm(dm16, w2)                                # [L, B, d]
        sg = torch.sigmoid(pre.float())
        dpre = dh.float() * (sg * (1.0 + pre.float() * (1.0 - sg)))  # SiLU'
        dp16 = dpre.to(torch.bfloat16)
        db0 = dpre.sum(1)                                       # [L, d]
        dcond = torch.bmm(dp16, w0).sum(0, dtype=torch.float32)  # [B, d]
        views = [sink.grad_view(p) if sink is not None else None for p in ada]
        grads: list = [None] * len(ada)
        direct = all(v is not None for v in views) and h.shape[1] % 8 == 0
        if direct:
            for l in range(nl):
                _wgrad(dp16[l], c16, into=views[4 * l])
                _wgrad(dm16[l], h[l], into=views[4 * l + 2])
            torch._foreach_add_([views[4 * l + 1] for l in range(nl)], list(db0.unbind(0)))
            torch._foreach_add_([views[4 * l + 3] for l in range(nl)], list(db2.unbind(0)))
        else:
            dw0 = torch.bmm(dp16.transpose(1, 2), c16[None].expand(nl, -1, -1)).float()
            dw2 = torch.bmm(dm16.transpose(1, 2), h).float()
            for l in range(nl):
                for j, g in enumerate((dw0[l], db0[l], dw2[l], db2[l])):
                    if views[4 * l + j] is not None:
                        views[4 * l + j].add_(g)
                    else:
                        grads[4 * l + j] = g
        return (None, dcond, *grads)


def _attention_backward(qkv: Tensor, att: Tensor, datt: Tensor, B: int, T: int, heads: int, hd: int, drop: tuple[float, int] = (0.0, 0),
                        lse: Tensor | None = None) -> Tensor:
    """d(qkv) of att = attention(qkv) on the packed [B*T][3*dim] layout (two mma.sync kernels, attention_bwd.cu)."""
    dqkv = torch.empty_like(qkv)
    ws = torch.empty((2, B * heads * T), dtype=torch.float32, device=qkv.device)
    lse_ws = lse if lse is not None else ws[0]
    L.check(L.load().bsi_attention_backward_bf16(dqkv.data_ptr(), lse_ws.data_ptr(), ws[1].data_ptr(), qkv.data_ptr(), att.data_ptr(), datt.data_ptr(),
                                                 B, T, heads, hd, drop[0], drop[1], int(lse is not None), _st(qkv.device)), "bsi_attention_backward_bf16")
    return dqkv


def trainable_parameters(model) -> list[Tensor]:
    """The parameters DiTTrainFunction differentiates, in its argument order (adaLN / time embedding go through ``mods``)."""
    d = model.dit
    ps = [d.patch_encoder.weight, d.patch_encoder.bias]
    for blk in d.blocks:
        ps += [blk.attn.to_qkv.weight, blk.attn.to_qkv.bias, blk.attn.to_out.weight, blk.attn.to_out.bias,
               blk.mlp[0].weight, blk.mlp[0].bias, blk.mlp[2].weight, blk.mlp[2].bias]
    ps += [d.patch_decoder[0].weight, d.patch_decoder[0].bias, d.patch_decoder[1].weight, d.patch_decoder[1].bias]
    return ps


def _layer_seed(seed: int, k: int) -> int:
    """Distinct 32-bit stream key per (call, layer, site) from the call's seed (host-side integer hash)."""
    x = (seed + 0x9E3779B9 * (k + 1)) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & 0xFFFFFFFF
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & 0xFFFFFFFF
    return x ^ (x >> 16)


def forward_train(model, mu: Tensor, t: Tensor, in_scale: Tensor | None, seed: int | None = None) -> Tensor:
    """f(in_scale * mu, t) with autograd support for every parameter of the DiT.

    In ``train()`` mode with ``dropout=p`` the reference's two dropout sites are active (dit.py:43-44 on the attention
    probabilities, :101 on the MLP input); their masks are stateless hashes of a per-call seed drawn from torch's CPU
    generator (or ``seed``), regenerated in the backward kernels instead of being stored."""
    blk0 = model.dit.blocks[0]
    drop_p = float(blk0.dropout.p) if model.training and isinstance(blk0.dropout, torch.nn.Dropout) else 0.0
    if drop_p > 0 and seed is None:
        seed = int(torch.randint(0, 2**31 - 1, (1,)).item())  # CPU generator: no device synchronisation
    cond = model.dit.t_embedding(t.to(torch.float32))
    ada = []
    for blk in model.dit.blocks:
        ada += [blk.adaLN_modulation[0].weight, blk.adaLN_modulation[0].bias, blk.adaLN_modulation[2].weight, blk.adaLN_modulation[2].bias]
    mods = _AdaLNChain.apply(model, cond, *ada)  # [L, B, 6d]
    return DiTTrainFunction.apply(model, mu, in_scale, (drop_p, int(seed or 0)), mods, *trainable_parameters(model))
