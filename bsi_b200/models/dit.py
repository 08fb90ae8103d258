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


class DenoisingDiT(nn.Module):
    """Diffusion Transformer denoiser f(mu, t) -> x-shaped output, native on B200."""

    bsi_native = True

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
                                fourier_features.n_max if fourier_features is not None else -1)
        self._engine = None
        self._arena = None
        self._packed_sig = None
        self._scratch: dict = {}

    # ---- engine / parameter arena ---------------------------------------------------------------
    def __del__(self):
        eng = getattr(self, "_engine", None)
        if eng:
            try:
                L.load().bsi_dit_destroy(eng)
            except Exception:
                pass

    def _named_tensors(self):
        yield from self.state_dict(keep_vars=True).items()
        yield "dit.patch_pos_embedding", self.dit.patch_pos_embedding
        yield "dit.t_embedding.scale", self.dit.t_embedding.scale
        yield "dit.t_embedding.bias", self.dit.t_embedding.bias

    def _signature(self):
        """(address, version) of every tensor the engine packs; a change triggers re-packing.

        Tensors created under torch.inference_mode() carry no version counter: for those only a new
        allocation is detected and in-place updates need an explicit `repack()`."""
        sig = []
        for _, p in self._named_tensors():
            try:
                version = p._version
            except RuntimeError:
                version = -1
            sig.append((p.data_ptr(), version))
        return tuple(sig)

    def _ensure_packed(self, device):
        lib = L.load()
        if self._engine is None:
            handle = C.c_void_p()
            L.check(lib.bsi_dit_create(C.byref(self._cfg), C.byref(handle)), "bsi_dit_create")
            self._engine = handle
        sig = self._signature()
        if self._arena is None or self._arena.device != device or sig != self._packed_sig:
            nbytes = lib.bsi_dit_param_bytes(self._engine)
            if self._arena is None or self._arena.device != device:
                self._arena = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
            base = (self._arena.data_ptr() + 255) // 256 * 256
            st = L.stream_ptr(device)
            L.check(lib.bsi_dit_bind_params(self._engine, base, nbytes), "bsi_dit_bind_params")
            for key, p in self._named_tensors():
                if key.startswith("fourier_features"):
                    continue
                src = p.detach()
                if src.device != device or src.dtype != torch.float32 or not src.is_contiguous():
                    src = src.to(device=device, dtype=torch.float32).contiguous()
                L.check(lib.bsi_dit_set_param(self._engine, key.encode(), src.data_ptr(), src.numel(), st), f"bsi_dit_set_param({key})")
            missing = lib.bsi_dit_missing_params(self._engine)
            if missing:
                raise L.BsiNativeError(f"{missing} DiT parameters were not packed")
            self._packed_sig = sig
        return self._engine

    def repack(self):
        """Force re-packing of the bf16 parameter arena on the next call."""
        self._packed_sig = None

    def _buffer(self, name: str, nbytes: int, device) -> Tensor:
        buf = self._scratch.get(name)
        if buf is None or buf.device != device or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            self._scratch[name] = buf
        return buf

    def _conditioning(self, eng, t: Tensor, name: str = "cond") -> Tensor:
        """cond[layer][row][6*dim] = adaLN_modulation(t_embedding(t)) for all layers (two GEMM launches)."""
        lib, dev, rows = L.load(), t.device, t.numel()
        cond = self._buffer(name, lib.bsi_dit_cond_bytes(eng, rows), dev)
        scratch_bytes = lib.bsi_dit_cond_scratch_bytes(eng, rows)
        scratch = self._buffer("cond_scratch", scratch_bytes, dev)
        t = t.detach().to(torch.float32).contiguous()
        L.check(lib.bsi_dit_conditioning(eng, cond.data_ptr(), t.data_ptr(), rows, scratch.data_ptr(), scratch_bytes, L.stream_ptr(dev)),
                "bsi_dit_conditioning")
        return cond

    # ---- forward --------------------------------------------------------------------------------
    def _check_input(self, mu: Tensor):
        if not mu.is_cuda:
            raise L.BsiNativeError("DenoisingDiT runs on CUDA only (no CPU fallback)")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "bsi_b200.DenoisingDiT has no backward kernels yet: call it under torch.no_grad()/inference_mode() "
                "or with parameters frozen (training through the native DiT is scheduled after the sampling path)"
            )
        if tuple(mu.shape[1:]) != self.data_shape:
            raise ValueError(f"expected input of shape [B, {self.data_shape}], got {tuple(mu.shape)}")

    def forward_scaled(self, mu: Tensor, t: Tensor, in_scale: Tensor | None) -> Tensor:
        """f(in_scale[b] * mu[b], t[b]); the scaling is fused into the operand builder."""
        self._check_input(mu)
        dev, lib, B = mu.device, L.load(), mu.shape[0]
        with torch.cuda.device(dev):
            eng = self._ensure_packed(dev)
            mu = mu.detach().to(torch.float32).contiguous()
            cond = self._conditioning(eng, t)
            ws_bytes = lib.bsi_dit_workspace_bytes(eng, B)
            ws = self._buffer("workspace", ws_bytes, dev)
            scale = torch.ones(1, dtype=torch.float32, device=dev) if in_scale is None else in_scale.detach().to(torch.float32).contiguous()
            out = torch.empty_like(mu)
            L.check(
                lib.bsi_dit_forward(eng, out.data_ptr(), mu.data_ptr(), L.rowref(scale, 0 if in_scale is None else 1), cond.data_ptr(), B, 0, 1, 0,
                                    None, B, ws.data_ptr(), ws_bytes, L.stream_ptr(dev)),
                "bsi_dit_forward",
            )
        return out

    def forward(self, mu: Tensor, t: Tensor) -> Tensor:
        return self.forward_scaled(mu, t, None)

    def residual_stream(self, B: int) -> Tensor:
        """Debug: fp32 token stream [B*T, dim] after the last block of the previous forward."""
        dev = self._scratch["workspace"].device
        T = (self.data_shape[1] // self._cfg.patch) * (self.data_shape[2] // self._cfg.patch)
        out = torch.empty((B * T, self._cfg.dim), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L.check(L.load().bsi_dit_peek(self._engine, 0, out.data_ptr(), B, self._scratch["workspace"].data_ptr(), L.stream_ptr(dev)), "bsi_dit_peek")
        return out

    # ---- k-step sampler with a CUDA graph per step (reference bsi/bsi.py:328-336) -------------------
    @torch.no_grad()
    def sample_loop(self, n: int, lam0_rsqrt: Tensor, coef: Tensor, c_in: Tensor, t_rows: Tensor, k: int, seed: int, sample_offset: int,
                    precond: int, use_graph: bool = True) -> Tensor:
        """Run mu_0 -> ... -> mu_k -> x_hat with in-kernel Philox noise.  coef[k+1,8], c_in[k+1], t_rows[k+1]."""
        dev, lib = coef.device, L.load()
        D = self.data_shape[0] * self.data_shape[1] * self.data_shape[2]
        with torch.cuda.device(dev):
            eng = self._ensure_packed(dev)
            cond = self._conditioning(eng, t_rows, "cond_sampler")
            ws_bytes = lib.bsi_dit_workspace_bytes(eng, n)
            ws = self._buffer("workspace", ws_bytes, dev)
            mu = torch.empty((n, *self.data_shape), dtype=torch.float32, device=dev)
            f = torch.empty_like(mu)
            step = torch.zeros(1, dtype=torch.int32, device=dev)
            scale_ref = L.rowref(c_in, 0, 1) if precond else L.rowref(torch.ones(1, dtype=torch.float32, device=dev), 0, 0)
            self._keepalive = (scale_ref, c_in)

            def enqueue_forward():
                L.check(
                    lib.bsi_dit_forward(eng, f.data_ptr(), mu.data_ptr(), scale_ref, cond.data_ptr(), k + 1, 0, 0, 1, step.data_ptr(), n,
                                        ws.data_ptr(), ws_bytes, L.stream_ptr(dev)),
                    "bsi_dit_forward",
                )

            def enqueue_step():
                enqueue_forward()
                L.check(
                    lib.bsi_step_fused(mu.data_ptr(), f.data_ptr(), coef.data_ptr(), step.data_ptr(), 0, precond,
                                       L.noise(seed=seed, sample_base=sample_offset, draw=1), None, None, n, D, L.stream_ptr(dev)),
                    "bsi_step_fused",
                )
                L.check(lib.bsi_step_advance(step.data_ptr(), L.stream_ptr(dev)), "bsi_step_advance")

            L.check(lib.bsi_sample_init(mu.data_ptr(), lam0_rsqrt.data_ptr(), L.noise(seed=seed, sample_base=sample_offset, draw=0), n, D,
                                        L.stream_ptr(dev)), "bsi_sample_init")
            if use_graph and k > 1:
                enqueue_forward()  # eager warm-up: first-launch attribute setup happens outside capture; idempotent
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    enqueue_step()
                for _ in range(k):
                    graph.replay()
            else:
                for _ in range(k):
                    enqueue_step()
            # final prediction at t = 1 (table row k; the step counter now equals k)
            enqueue_forward()
            if not precond:
                return f
            x_hat = torch.empty_like(mu)
            L.check(
                lib.bsi_edm_combine(x_hat.data_ptr(), mu.data_ptr(), f.data_ptr(), L.rowref(coef, 0, 8, 0), L.rowref(coef, 0, 8, 1),
                                    step.data_ptr(), n, D, L.stream_ptr(dev)),
                "bsi_edm_combine",
            )
            self.last_sampler_state = {"mu": mu, "step": step}
        return x_hat
