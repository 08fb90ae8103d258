"""DenoisingDiT with the reference's constructor, parameter names and state_dict layout
(reference bsi/models/dit.py:184-233), executed by the sm_100a DiT engine of libbsi_b200.so.

The nn.Module tree below only *holds* the fp32 master parameters under the reference's keys
(`dit.patch_encoder.*`, `dit.blocks.{i}.attn.to_qkv.*`, ... — so reference checkpoints load
unchanged); `forward` never runs those sub-modules.  Parameters are packed to bf16 into a
device arena whenever they change, and the forward pass is a sequence of native kernels:
operand builder (patchify + Fourier features + c_in scaling) -> tcgen05 GEMMs with fused
bias / GELU / gate+residual / positional / unpatchify epilogues, LayerNorm+adaLN kernels and
fused attention.  `sample_loop` replays a CUDA graph of one sampler step k times.
"""

from __future__ import annotations

import ctypes as C
from functools import partial

import torch
from torch import Tensor, nn

from .. import _lib as L
from ._native import NativeDenoiser
from ..nn import MLP, FourierFeatures
from .pos_emb import NyquistPositionalEmbedding


class _AttentionParams(nn.Module):
    def __init__(self, dim: int, heads: int, dropout: float):
        super().__init__()
        self.heads, self.dropout = heads, dropout
        self.to_qkv = nn.Linear(dim, 3 * dim)
        self.to_out = nn.Linear(dim, dim)


class _BlockParams(nn.Module):
    """adaLN-Zero block parameters (reference bsi/models/dit.py:58-85)."""

    def __init__(self, size: int, heads: int, mlp_ratio: int, dropout: float | None):
        super().__init__()
        self.norm = nn.LayerNorm(size, elementwise_affine=False)
        self.attn = _AttentionParams(size, heads, dropout if dropout is not None else 0.0)
        self.dropout = nn.Dropout(dropout) if dropout is not None else nn.Identity()
        self.mlp = MLP(size, size, hidden_features=[mlp_ratio * size], actfn=partial(nn.GELU, approximate="tanh"))
        self.adaLN_modulation = nn.Sequential(nn.Linear(size, size), nn.SiLU(), nn.Linear(size, 6 * size))
        nn.init.zeros_(self.adaLN_modulation[-1].weight)
        nn.init.zeros_(self.adaLN_modulation[-1].bias)


class _DiTParams(nn.Module):
    """Parameter/buffer container matching reference `DiT` (bsi/models/dit.py:106-172)."""

    def __init__(self, input_size, patch_size, in_channels, out_channels, hidden_size, depth, heads, mlp_ratio, dropout):
        super().__init__()
        self.input_size, self.patch_size = tuple(input_size), patch_size
        self.in_channels, self.out_channels = in_channels, out_channels
        height, width = input_size
        gh, gw = height // patch_size, width // patch_size
        pos = NyquistPositionalEmbedding(hidden_size // 2, max(height, width))
        rows, cols = pos(torch.linspace(0, 1, gh)), pos(torch.linspace(0, 1, gw))
        table = torch.cat((rows.repeat_interleave(gw, dim=0), cols.repeat(gh, 1)), dim=1)
        self.register_buffer("patch_pos_embedding", table, persistent=False)
        self.t_embedding = NyquistPositionalEmbedding(hidden_size, 1000)
        self.patch_encoder = nn.Linear(patch_size**2 * in_channels, hidden_size)
        self.blocks = nn.ModuleList([_BlockParams(hidden_size, heads, mlp_ratio, dropout) for _ in range(depth)])
        self.patch_decoder = nn.Sequential(nn.LayerNorm(hidden_size), nn.Linear(hidden_size, patch_size**2 * out_channels))


class DenoisingDiT(NativeDenoiser):
    """Diffusion Transformer denoiser f(mu, t) -> x-shaped output, native on B200."""

    _api = "bsi_dit_"

    def __init__(self, data_shape, patch_size: int, dim: int, depth: int, heads: int, dropout: float | None = None,
                 fourier_features: FourierFeatures | None = None, **kwargs):
        super().__init__()
        self.data_shape = tuple(data_shape)
        assert len(self.data_shape) == 3, "Only works for 2D images"
        self.fourier_features = fourier_features
        channels = self.data_shape[0]
        in_channels = channels + (channels * fourier_features.n_features() if fourier_features is not None else 0)
        self.dit = _DiTParams(self.data_shape[1:], patch_size, in_channels, channels, dim, depth, heads, 4, dropout)
        self._cfg = L.DitConfig(channels, self.data_shape[1], self.data_shape[2], patch_size, dim, depth, heads,
                                fourier_features.n_min if fourier_features is not None else 0,
                                fourier_features.n_max if fourier_features is not None else -1, 0)
        self._init_native()

    # ---- evaluation precision (extension; the reference evaluates in fp32, bsi/lightning/plugins.py:7-24) ----------------------
    @property
    def precision(self) -> str:
        """"bf16" (default): bf16 tensor-core operands, fp32 accumulation and residual stream.  "fp32": every GEMM operand is split
        into three bf16 terms and attention runs in fp32 (csrc/exact_kernels.cu) -- fp32-level agreement with the reference
        (the 1e-5 tier of the trajectory test) at 3-5x the cost.  Inference only; training always uses the bf16 path."""
        return "fp32" if self._cfg.exact else "bf16"

    def set_precision(self, precision: str) -> "DenoisingDiT":
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        exact = 1 if precision == "fp32" else 0
        if exact != self._cfg.exact:
            self._cfg.exact = exact
            if self._engine:  # arena layout and workspaces differ: rebuild everything on the next call
                self._fn("destroy")(self._engine)
            self._engine, self._arena, self._packed_sig, self._scratch = None, None, None, {}
        return self

    def forward_scaled(self, mu: Tensor, t: Tensor, in_scale: Tensor | None) -> Tensor:
        """Inference: the fused engine.  Under autograd with trainable parameters: the differentiable path of
        ``dit_train`` (same kernels driven one by one, activations saved, native dgrad / wgrad GEMMs)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .dit_train import forward_train

            if not mu.is_cuda:
                raise L.BsiNativeError("DenoisingDiT runs on CUDA only (no CPU fallback)")
            if tuple(mu.shape[1:]) != self.data_shape:
                raise ValueError(f"expected input of shape [B, {self.data_shape}], got {tuple(mu.shape)}")
            return forward_train(self, mu, t, in_scale)
        return super().forward_scaled(mu, t, in_scale)

    def _named_tensors(self):
        yield from self.state_dict(keep_vars=True).items()
        yield "dit.patch_pos_embedding", self.dit.patch_pos_embedding
        yield "dit.t_embedding.scale", self.dit.t_embedding.scale
        yield "dit.t_embedding.bias", self.dit.t_embedding.bias

    def residual_stream(self, B: int) -> Tensor:
        """Debug: fp32 token stream [B*T, dim] after the last block of the previous forward."""
        dev = self._scratch["workspace"].device
        T = (self.data_shape[1] // self._cfg.patch) * (self.data_shape[2] // self._cfg.patch)
        out = torch.empty((B * T, self._cfg.dim), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L.check(L.load().bsi_dit_peek(self._engine, 0, out.data_ptr(), B, self._scratch["workspace"].data_ptr(), L.stream_ptr(dev)), "bsi_dit_peek")
        return out
