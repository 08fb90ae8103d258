"""Shared host machinery of the native denoisers (DiT, VDM U-Net): parameter packing into a device arena, conditioning
tables, the forward call and the CUDA-graph sampler loop.  Sub-classes hold the fp32 master parameters under the
reference's state_dict keys and name the C-ABI family (`bsi_dit_*` / `bsi_unet_*`) that executes them."""

from __future__ import annotations

import ctypes as C

import torch
from torch import Tensor, nn

from .. import _lib as L


class NativeDenoiser(nn.Module):
    bsi_native = True
    _api = ""  # "bsi_dit_" or "bsi_unet_"

    def _init_native(self):
        self._engine = None
        self._arena = None
        self._packed_sig = None
        self._scratch: dict = {}
        self._epoch = 0  # bumped whenever the engine or the arena is (re)created: part of the sampler-plan key

    def _fn(self, name: str):
        return getattr(L.load(), self._api + name)

    def __getstate__(self):
        """Copies (``copy.deepcopy`` for the EMA model, bsi/tasks/ema_pytorch.py:203-236) and pickles carry the parameters only:
        the engine handle, the packed arena, workspaces / CUDA graphs and an attached optimizer sink belong to this instance and
        are rebuilt lazily by the copy's first call."""
        state = self.__dict__.copy()
        state.update(_engine=None, _arena=None, _packed_sig=None, _scratch={}, _epoch=0)
        state.pop("_grad_sink", None)
        state.pop("_train_wcache", None)  # bf16 weight copies of the training path: rebuilt on first use
        return state

    def __del__(self):
        eng = getattr(self, "_engine", None)
        if eng:
            try:
                self._fn("destroy")(eng)
            except Exception:
                pass

    # ---- parameters ------------------------------------------------------------------------------------
    def _named_tensors(self):
        """(reference key, tensor) for every tensor the engine packs; sub-classes add non-persistent buffers."""
        yield from self.state_dict(keep_vars=True).items()

    def _signature(self):
        """(address, version, arena generation) of every packed tensor; a change triggers re-packing.

        The native optimizer / EMA kernels (bsi_b200/optim.py) write parameters through raw pointers, which torch's version
        counters do not see: parameters adopted by a ``FlatArena`` therefore also carry the arena's generation counter, which
        every native write bumps.  Tensors created under torch.inference_mode() carry no version counter: for those only a
        new allocation is detected and in-place updates need an explicit `repack()`."""
        sig = []
        for _, p in self._named_tensors():
            try:
                version = p._version
            except RuntimeError:
                version = -1
            arena = getattr(p, "_bsi_arena", None)
            sig.append((p.data_ptr(), version, arena[0].generation if arena is not None else 0))
        return tuple(sig)

    def _ensure_packed(self, device):
        if self._engine is None:
            handle = C.c_void_p()
            L.check(self._fn("create")(C.byref(self._cfg), C.byref(handle)), self._api + "create")
            self._engine = handle
            self._epoch = getattr(self, "_epoch", 0) + 1
        sig = self._signature()
        if self._arena is None or self._arena.device != device or sig != self._packed_sig:
            nbytes = self._fn("param_bytes")(self._engine)
            if self._arena is None or self._arena.device != device:
                self._arena = torch.zeros(nbytes + 256, dtype=torch.uint8, device=device)
                self._epoch = getattr(self, "_epoch", 0) + 1
            base = (self._arena.data_ptr() + 255) // 256 * 256
            st = L.stream_ptr(device)
            L.check(self._fn("bind_params")(self._engine, base, nbytes), self._api + "bind_params")
            for key, p in self._named_tensors():
                if key.startswith("fourier_features"):
                    continue
                src = p.detach()
                if src.device != device or src.dtype != torch.float32 or not src.is_contiguous():
                    src = src.to(device=device, dtype=torch.float32).contiguous()
                L.check(self._fn("set_param")(self._engine, key.encode(), src.data_ptr(), src.numel(), st), f"{self._api}set_param({key})")
            missing = self._fn("missing_params")(self._engine)
            if missing:
                raise L.BsiNativeError(f"{missing} denoiser parameters were not packed")
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
        """Conditioning table for every block and every row of `t` (a few GEMM launches)."""
        dev, rows = t.device, t.numel()
        cond = self._buffer(name, self._fn("cond_bytes")(eng, rows), dev)
        scratch_bytes = self._fn("cond_scratch_bytes")(eng, rows)
        scratch = self._buffer("cond_scratch", scratch_bytes, dev)
        t = t.detach().to(torch.float32).contiguous()
        L.check(self._fn("conditioning")(eng, cond.data_ptr(), t.data_ptr(), rows, scratch.data_ptr(), scratch_bytes, L.stream_ptr(dev)),
                self._api + "conditioning")
        return cond

    # ---- forward -----------------------------------------------------------------------------------------
    def _check_input(self, mu: Tensor):
        if not mu.is_cuda:
            raise L.BsiNativeError(f"{type(self).__name__} runs on CUDA only (no CPU fallback)")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                f"bsi_b200.{type(self).__name__} has no backward kernels yet (only DenoisingDiT is differentiable): call it under "
                "torch.no_grad()/inference_mode() or with parameters frozen, or train the reference's PyTorch module inside bsi_b200.BSI"
            )
        if tuple(mu.shape[1:]) != self.data_shape:
            raise ValueError(f"expected input of shape [B, {self.data_shape}], got {tuple(mu.shape)}")

    def forward_scaled(self, mu: Tensor, t: Tensor, in_scale: Tensor | None) -> Tensor:
        """f(in_scale[b] * mu[b], t[b]); the scaling is fused into the operand builder."""
        self._check_input(mu)
        dev, B = mu.device, mu.shape[0]
        t = t.reshape(-1)
        if t.numel() == 1 and B != 1:  # one time for the whole batch: the reference's modulate() broadcasts it (bsi/models/dit.py:50-55)
            t = t.expand(B)
        if t.numel() != B:
            raise ValueError(f"t has {t.numel()} elements for a batch of {B}")
        if in_scale is not None:
            in_scale = in_scale.reshape(-1)
            if in_scale.numel() == 1 and B != 1:
                in_scale = in_scale.expand(B)
            if in_scale.numel() != B:
                raise ValueError(f"in_scale has {in_scale.numel()} elements for a batch of {B}")
        with torch.cuda.device(dev):
            eng = self._ensure_packed(dev)
            mu = mu.detach().to(torch.float32).contiguous()
            cond = self._conditioning(eng, t)
            ws_bytes = self._fn("workspace_bytes")(eng, B)
            ws = self._buffer("workspace", ws_bytes, dev)
            scale = torch.ones(1, dtype=torch.float32, device=dev) if in_scale is None else in_scale.detach().to(torch.float32).contiguous()
            out = torch.empty_like(mu)
            L.check(
                self._fn("forward")(eng, out.data_ptr(), mu.data_ptr(), L.rowref(scale, 0 if in_scale is None else 1), cond.data_ptr(), B, 0, 1, 0,
                                    None, B, ws.data_ptr(), ws_bytes, L.stream_ptr(dev)),
                self._api + "forward",
            )
        return out

    def forward(self, mu: Tensor, t: Tensor) -> Tensor:
        return self.forward_scaled(mu, t, None)

    # ---- k-step sampler with a CUDA graph per step (reference bsi/bsi.py:328-336) ---------------------------
    _MAX_PLANS = 4

    @torch.no_grad()
    def sample_loop(self, n: int, lam0_rsqrt: Tensor, coef: Tensor, c_in: Tensor, t_rows: Tensor, k: int, seed: int, sample_offset: int,
                    precond: int, use_graph: bool = True, plans: dict | None = None) -> Tensor:
        """Run mu_0 -> ... -> mu_k -> x_hat with in-kernel Philox noise.  coef[k+1,8], c_in[k+1], t_rows[k+1].

        The captured graph of one step (denoiser forward + fused update + counter increment) only refers to buffers owned by
        its plan (belief state, step tables, Philox key, step counter) and to this denoiser's arena / conditioning table /
        workspace, so a plan is reused by every later call with the same (n, k): the call refreshes the tables, the key and
        the counter and replays.  `plans` is the caller's cache (``BSI._plans``, cleared by ``set_model``)."""
        dev, lib = coef.device, L.load()
        D = self.data_shape[0] * self.data_shape[1] * self.data_shape[2]
        forward = self._fn("forward")
        with torch.cuda.device(dev):
            eng = self._ensure_packed(dev)
            cond = self._conditioning(eng, t_rows, "cond_sampler")
            ws_bytes = self._fn("workspace_bytes")(eng, n)
            ws = self._buffer("workspace", ws_bytes, dev)
            # a plan is valid for these buffers only (they are re-allocated when a larger request comes along)
            key = (n, k, precond, dev.index, id(self), self._epoch, self._arena.data_ptr(), cond.data_ptr(), ws.data_ptr())
            plan = plans.get(key) if plans is not None else None
            if plan is None:
                plan = {
                    "mu": torch.empty((n, *self.data_shape), dtype=torch.float32, device=dev), "f": None,
                    "step": torch.zeros(1, dtype=torch.int32, device=dev), "coef": torch.empty_like(coef), "c_in": torch.empty_like(c_in),
                    "key": torch.zeros(2, dtype=torch.int64, device=dev), "ones": torch.ones(1, dtype=torch.float32, device=dev), "graph": None,
                }
                plan["f"] = torch.empty_like(plan["mu"])
                if plans is not None:
                    while len(plans) >= self._MAX_PLANS:
                        plans.pop(next(iter(plans)))
                    plans[key] = plan
            mu, f, step, key_buf = plan["mu"], plan["f"], plan["step"], plan["key"]
            plan["coef"].copy_(coef), plan["c_in"].copy_(c_in)
            step.zero_()
            # {seed, sample_base} as two's-complement int64 (the kernels read them as uint64)
            to_i64 = lambda v: (v & 0xFFFFFFFFFFFFFFFF) - (1 << 64) if (v & 0xFFFFFFFFFFFFFFFF) >= (1 << 63) else (v & 0xFFFFFFFFFFFFFFFF)
            key_buf.copy_(torch.tensor([to_i64(int(seed)), to_i64(int(sample_offset))], dtype=torch.int64), non_blocking=False)
            scale_ref = L.rowref(plan["c_in"], 0, 1) if precond else L.rowref(plan["ones"], 0, 0)

            def enqueue_forward():
                L.check(
                    forward(eng, f.data_ptr(), mu.data_ptr(), scale_ref, cond.data_ptr(), k + 1, 0, 0, 1, step.data_ptr(), n,
                            ws.data_ptr(), ws_bytes, L.stream_ptr(dev)),  # the current stream: the capture stream while capturing
                    self._api + "forward",
                )

            def enqueue_step():
                enqueue_forward()
                L.check(
                    lib.bsi_step_fused(mu.data_ptr(), f.data_ptr(), plan["coef"].data_ptr(), step.data_ptr(), 0, precond,
                                       L.noise(draw=1, key=key_buf), None, None, n, D, L.stream_ptr(dev)),
                    "bsi_step_fused",
                )
                L.check(lib.bsi_step_advance(step.data_ptr(), L.stream_ptr(dev)), "bsi_step_advance")

            L.check(lib.bsi_sample_init(mu.data_ptr(), lam0_rsqrt.data_ptr(), L.noise(draw=0, key=key_buf), n, D, L.stream_ptr(dev)),
                    "bsi_sample_init")
            if use_graph and k > 1:
                if plan["graph"] is None:
                    enqueue_forward()  # eager warm-up: first-launch attribute setup happens outside capture; idempotent
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        enqueue_step()
                    plan["graph"] = graph
                for _ in range(k):
                    plan["graph"].replay()
            else:
                for _ in range(k):
                    enqueue_step()
            # final prediction at t = 1 (table row k; the step counter now equals k)
            enqueue_forward()
            self.last_sampler_state = {"mu": mu, "step": step}
            if not precond:
                return f.clone()
            x_hat = torch.empty_like(mu)
            L.check(
                lib.bsi_edm_combine(x_hat.data_ptr(), mu.data_ptr(), f.data_ptr(), L.rowref(plan["coef"], 0, 8, 0), L.rowref(plan["coef"], 0, 8, 1),
                                    step.data_ptr(), n, D, L.stream_ptr(dev)),
                "bsi_edm_combine",
            )
        return x_hat
