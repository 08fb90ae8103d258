"""Optimizer side of the reference's training step on native kernels.

The reference runs, after ``loss.backward()``: Lightning's ``clip_grad_norm_(1.0)`` (config/train.yaml:40),
``torch.optim.AdamW(fused=True)`` (config/task/optimizer/adamw.yaml) and ``EMA.update()`` (bsi/tasks/bsi.py:196-198,
bsi/tasks/ema_pytorch.py:316-341) -- one norm pass, one multiply pass, the fused Adam kernel and a ``_foreach_lerp_``.
Here the parameters, gradients, both moments and the EMA weights live in flat fp32 arenas and one step is two launches:
``bsi_grad_sumsq`` (read g) and ``bsi_adamw_ema_step`` (clip, AdamW, EMA action, optional bf16 copy, gradient zeroing).

    opt = bsi_b200.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
    ema = bsi_b200.optim.create_ema(model, beta=0.9999, update_after_step=1000, update_every=1)
    opt.attach_ema(ema)                 # optional: the EMA action of the coming ema.update() runs inside opt.step()
    loss.backward(); opt.step(); ema.update(); opt.zero_grad()

Same constructor arguments, hyper-parameter defaults, ``state_dict`` layout and call order as the classes they replace.
There is no CPU path: parameters must be fp32 CUDA tensors and the shared library must load.
"""

from __future__ import annotations

import ctypes
from copy import deepcopy

import torch
from torch import Tensor, nn

from . import _lib as L

__all__ = ["AdamW", "EMA", "create_ema", "FlatArena"]

_ALIGN = 4  # elements: every tensor starts on a 16-byte boundary so the kernels can use float4


def unreduced_ranges(done: list[tuple[int, int]], numel: int) -> list[tuple[int, int]]:
    """Complement of the element ranges `done` (already all-reduced during the backward) inside [0, numel)."""
    out, start = [], 0
    for a, b in sorted(done) + [(numel, numel)]:
        if a > start:
            out.append((start, a))
        start = max(start, b)
    return out


class FlatArena:
    """One contiguous fp32 CUDA buffer holding a list of tensors back to back (each padded to 4 elements).

    ``adopt`` re-points ``p.data`` of every parameter at its slice, so the parameters *are* the arena from then on
    (``model.to(...)`` or ``p.data = ...`` afterwards would detach them; construct the optimizer after moving the model)."""

    def __init__(self, shapes: list[torch.Size], device: torch.device):
        self.shapes = list(shapes)
        self.offsets, off = [], 0
        for s in self.shapes:
            self.offsets.append(off)
            off += (s.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = max(off, _ALIGN)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=device)
        # bumped by every native kernel that writes the arena through raw pointers (torch's version counters do not see those
        # writes): NativeDenoiser._signature() includes it, so the packed bf16 weights are rebuilt before the next evaluation
        self.generation = 0

    def view(self, i: int) -> Tensor:
        n = self.shapes[i].numel()
        return self.flat[self.offsets[i] : self.offsets[i] + n].view(self.shapes[i])

    def views(self) -> list[Tensor]:
        return [self.view(i) for i in range(len(self.shapes))]

    @classmethod
    def like(cls, other: "FlatArena") -> "FlatArena":
        return cls(other.shapes, other.flat.device)

    @classmethod
    def adopt(cls, params: list[Tensor]) -> "FlatArena":
        """Arena over ``params`` (reused if they already are, in this order, the slices of one arena)."""
        if not params:
            raise ValueError("no parameters")
        first = getattr(params[0], "_bsi_arena", None)
        if first is not None and len(first[0].shapes) == len(params) and all(
            getattr(p, "_bsi_arena", (None, -1)) == (first[0], i) and p.data_ptr() == first[0].view(i).data_ptr() for i, p in enumerate(params)
        ):
            return first[0]
        for p in params:
            if not (p.is_cuda and p.dtype == torch.float32):
                raise RuntimeError(f"bsi_b200.optim needs fp32 CUDA parameters, got {p.dtype} on {p.device} (there is no CPU path)")
        arena = cls([p.shape for p in params], params[0].device)
        with torch.no_grad():
            for i, p in enumerate(params):
                v = arena.view(i)
                v.copy_(p.data)
                p.data = v
                p._bsi_arena = (arena, i)
        return arena


class EMA(nn.Module):
    """Exponential moving average of a model's weights: the subset of the reference's ``EMA`` that ``BSITraining`` uses
    (bsi/tasks/ema_pytorch.py:60-436 as configured by ``create_ema``, bsi/tasks/bsi.py:73-81): python-side step counters,
    copy until ``update_after_step``, then ``ema.lerp_(param, 1 - decay)`` with the warm-up decay
    ``1 - (1 + epoch/inv_gamma)^-power`` clamped to ``[min_value, beta]``, every ``update_every`` steps."""

    def __init__(self, model: nn.Module, ema_model: nn.Module | None = None, beta=0.9999, update_after_step=100, update_every=10,
                 inv_gamma=1.0, power=2 / 3, min_value=0.0, include_online_model=True, **kwargs):
        super().__init__()
        self.beta, self.update_after_step, self.update_every = beta, update_after_step, update_every
        self.inv_gamma, self.power, self.min_value = inv_gamma, power, min_value
        self.is_frozen = beta == 1.0
        self.include_online_model = include_online_model
        if include_online_model:
            self.online_model = model
        else:
            self.online_model = [model]  # not registered: managed (and saved) by the owner, like the reference
        self.ema_model = ema_model if ema_model is not None else deepcopy(model)
        for p in self.ema_model.parameters():
            p.detach_()
            p.requires_grad_(False)
        self.step, self.initted = 0, False
        self._arena: FlatArena | None = None  # EMA weights, same layout as the online parameters' arena
        self._online_arena: FlatArena | None = None
        self._fused_pending = False  # set by AdamW.step when it has already applied this update's action

    # --- state (python values, saved through extra_state like the reference: ema_pytorch.py:196-201)
    def get_extra_state(self):
        return {"initted": self.initted, "step": self.step}

    def set_extra_state(self, state):
        self.initted, self.step = state["initted"], state["step"]

    @property
    def model(self) -> nn.Module:
        return self.online_model if self.include_online_model else self.online_model[0]

    def forward(self, *args, **kwargs):
        return self.ema_model(*args, **kwargs)

    def get_current_decay(self) -> float:
        epoch = self.step - self.update_after_step - 1
        if epoch <= 0:
            return 0.0
        value = 1 - (1 + epoch / self.inv_gamma) ** -self.power
        return max(self.min_value, min(value, self.beta))

    def _arenas(self) -> tuple[FlatArena, FlatArena]:
        online = FlatArena.adopt(list(self.model.parameters()))
        if self._arena is None or self._online_arena is not online:
            ema_params = list(self.ema_model.parameters())
            if [p.shape for p in ema_params] != online.shapes:
                raise RuntimeError("EMA model and online model have different parameter lists")
            self._arena, self._online_arena = FlatArena.adopt(ema_params), online
        return self._arena, online

    def _next_action(self) -> tuple[int, float]:
        """(mode, weight) of the coming ``update()``: 0 nothing, 1 copy, 2 lerp -- bsi/tasks/ema_pytorch.py:316-341."""
        step = self.step
        if not self.initted:
            return 1, 1.0
        should_update = step % self.update_every == 0
        if should_update and step <= self.update_after_step:
            return 1, 1.0
        if should_update and not self.is_frozen:
            epoch = (step + 1) - self.update_after_step - 1
            decay = 0.0 if epoch <= 0 else max(self.min_value, min(1 - (1 + epoch / self.inv_gamma) ** -self.power, self.beta))
            return 2, 1.0 - decay
        return 0, 0.0

    @torch.no_grad()
    def update(self) -> None:
        mode, weight = self._next_action()
        self.step += 1
        self.initted = True
        if self._fused_pending:  # AdamW.step already ran this action inside its kernel
            self._fused_pending = False
        elif mode != 0:
            ema, online = self._arenas()
            L.check(L.load().bsi_ema_update(L.ptr(ema.flat), L.ptr(online.flat), online.numel, weight, mode, L.stream_ptr(online.flat.device)),
                    "bsi_ema_update")
            ema.generation += 1
        if mode == 1:  # copy_params_from_model_to_ema also copies the buffers (ema_pytorch.py:273-284)
            for b_ema, b in zip(self.ema_model.buffers(), self.model.buffers()):
                b_ema.copy_(b)


def create_ema(model, beta=0.9999, update_after_step=100, update_every=10, **kwargs) -> EMA:
    """``create_ema`` of bsi/tasks/bsi.py:73-81 (extra yaml keys such as ``power`` are swallowed there as well)."""
    return EMA(model, beta=beta, update_after_step=update_after_step, update_every=update_every, include_online_model=False)


class AdamW(torch.optim.Optimizer):
    """``torch.optim.AdamW`` with the global-norm clip, the EMA update and the bf16 weight copy fused into its step.

    Arguments and defaults of ``torch.optim.AdamW``; ``amsgrad``, ``maximize``, ``capturable`` and ``differentiable`` are not
    implemented (the reference trains with ``amsgrad: no``), ``foreach`` / ``fused`` are accepted and ignored.  All parameter
    groups must share their hyper-parameters (the reference has one group).
    ``max_grad_norm`` replaces Lightning's ``gradient_clip_val`` (set that to ``None`` when using it)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, *, maximize=False, foreach=None,
                 capturable=False, differentiable=False, fused=None, max_grad_norm: float | None = None, bf16_copy: bool = False):
        if amsgrad or maximize or capturable or differentiable:
            raise NotImplementedError("bsi_b200.optim.AdamW implements amsgrad=False, maximize=False, capturable=False, differentiable=False")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or not 0.0 <= weight_decay:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False))
        self.max_grad_norm = max_grad_norm
        self._lib = L.load()  # fails loudly when the CUDA library is missing
        plist = [p for g in self.param_groups for p in g["params"]]
        self._params = plist
        self._p = FlatArena.adopt(plist)
        self._g, self._m, self._v = FlatArena.like(self._p), FlatArena.like(self._p), FlatArena.like(self._p)
        self._bf16 = torch.empty(self._p.numel, dtype=torch.bfloat16, device=self._p.flat.device) if bf16_copy else None
        dev = self._p.flat.device
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._ws = torch.empty(int(self._lib.bsi_grad_sumsq_workspace_floats()), dtype=torch.float32, device=dev)
        self._t = 0
        self._pending, self._reduced, self._index = [], [], {}
        self._overlap, self._sync, self._group = False, True, None
        self._last_scale = 1.0
        self._ema: EMA | None = None
        self._grad_scale = 1.0  # 1/world_size once all_reduce_grads() has summed the arena over ranks
        self._clean_version = -1  # version counter of the gradient arena right after the kernel zeroed it
        for i, p in enumerate(plist):
            p.grad = self._g.view(i)
            self.state[p] = {"step": torch.tensor(0.0), "exp_avg": self._m.view(i), "exp_avg_sq": self._v.view(i)}

    # --- EMA fusion
    def attach_ema(self, ema: EMA) -> None:
        """Run the action of each coming ``ema.update()`` inside ``step()`` (call order stays ``opt.step(); ema.update()``)."""
        ema_arena, online = ema._arenas()
        if online is not self._p:
            raise RuntimeError("the EMA tracks a different parameter list than this optimizer")
        self._ema, self._ema_arena = ema, ema_arena

    def bf16_params(self) -> list[Tensor]:
        """bf16 copies of the parameters as of the last step (views of one buffer laid out like the fp32 arena)."""
        if self._bf16 is None:
            raise RuntimeError("construct the optimizer with bf16_copy=True")
        return [self._bf16[o : o + s.numel()].view(s) for o, s in zip(self._p.offsets, self._p.shapes)]

    def total_grad_norm(self) -> Tensor:
        """Device scalar: the gradient norm the last ``step()`` clipped against (``clip_grad_norm_``'s return value)."""
        return self._sumsq.sqrt() * self._last_scale

    def all_reduce_grads(self, group=None) -> None:
        """Data-parallel gradient exchange (``DistributedDataParallel`` of bsi/tasks/bsi.py:163-166) as sum all-reduces over
        the flat gradient arena (NCCL over NVLink); the division by the world size is folded into the next ``step()``.
        Call between ``backward()`` and ``step()`` with the model *not* wrapped in DDP.  Ranges that ``attach_model`` already
        reduced during the backward pass are skipped; everything else goes out as one collective per remaining range."""
        import torch.distributed as dist

        self._gather_grads()
        for a, b in unreduced_ranges(self._reduced, self._g.numel):
            dist.all_reduce(self._g.flat[a:b], op=dist.ReduceOp.SUM, group=group)
        for h in self._pending:
            h.wait()
        self._pending, self._reduced = [], []
        self._grad_scale = 1.0 / dist.get_world_size(group)

    # --- direct gradient sink for the native training path (bsi_b200/models/dit_train.py)
    def attach_model(self, model, overlap: bool = False, group=None) -> None:
        """Let the native ``DenoisingDiT`` backward write weight gradients straight into this optimizer's gradient arena (the
        weight-gradient GEMM accumulates in place: no temporaries, no autograd ``+=`` pass).  ``overlap=True`` additionally starts,
        under ``torch.distributed``, the all-reduce of each transformer block's gradient range as soon as the block's backward is
        done (what DDP's bucketed hooks do, bsi/tasks/bsi.py:163-166).  It is off by default because it does not pay on B200: the
        persistent GEMM kernels of the backward occupy every SM, NCCL's CTAs squeeze in between them and slow both down -- 8 GPUs,
        global batch 1024: 87.2 ms per step with per-block collectives against 84.6 ms with one all-reduce after the backward
        (profiles/train_8gpu_r02.jsonl)."""
        model._grad_sink = self
        self._overlap, self._group = overlap, group
        self._index = {id(p): i for i, p in enumerate(self._params)}

    def no_sync(self):
        """Context manager for gradient accumulation: backward passes inside it do not start all-reduces (like DDP.no_sync)."""
        opt = self

        class _NoSync:
            def __enter__(self):
                opt._sync = False

            def __exit__(self, *exc):
                opt._sync = True

        return _NoSync()

    def grad_view(self, p: Tensor) -> Tensor | None:
        """The arena slice that is ``p``'s gradient (None if ``p`` is not managed here or its ``.grad`` was replaced)."""
        i = self._index.get(id(p))
        if i is None:
            return None
        v = self._g.view(i)
        if p.grad is None or p.grad.data_ptr() != v.data_ptr():
            p.grad = v
        return v

    def grads_ready(self, params: list[Tensor]) -> None:
        """Called by the backward when the gradients of ``params`` (a contiguous run of the arena) are final for this step."""
        import torch.distributed as dist

        if not (self._overlap and self._sync and dist.is_available() and dist.is_initialized() and dist.get_world_size(self._group) > 1):
            return
        idx = sorted(self._index[id(p)] for p in params)
        if idx != list(range(idx[0], idx[-1] + 1)):
            return  # not contiguous in the arena: leave it to all_reduce_grads
        a = self._g.offsets[idx[0]]
        b = self._g.offsets[idx[-1]] + (self._g.shapes[idx[-1]].numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self._pending.append(dist.all_reduce(self._g.flat[a:b], op=dist.ReduceOp.SUM, group=self._group, async_op=True))
        self._reduced.append((a, b))

    def _gather_grads(self) -> None:
        """Gradients normally accumulate straight into the arena (``p.grad`` is a view of it).  If something replaced
        ``p.grad`` (``zero_grad(set_to_none=True)`` on the module, gradient-as-bucket-view DDP), copy it back in."""
        stray_dst, stray_src = [], []
        for i, p in enumerate(self._params):
            view = self._g.view(i)
            if p.grad is None:
                raise RuntimeError("bsi_b200.optim.AdamW needs a gradient for every parameter (parameter %d has none)" % i)
            if p.grad.data_ptr() != view.data_ptr():
                stray_dst.append(view)
                stray_src.append(p.grad)
                p.grad = view
        if stray_dst:
            torch._foreach_copy_(stray_dst, stray_src)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g0 = self.param_groups[0]
        for g in self.param_groups[1:]:
            if any(g[k] != g0[k] for k in ("lr", "betas", "eps", "weight_decay")):
                raise NotImplementedError("parameter groups with different hyper-parameters")
        self._gather_grads()
        st = L.stream_ptr(self._p.flat.device)
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        if clip:
            L.check(self._lib.bsi_grad_sumsq(L.ptr(self._sumsq), L.ptr(self._ws), L.ptr(self._g.flat), self._g.numel, st), "bsi_grad_sumsq")
        self._t += 1
        mode, weight = (0, 0.0)
        if self._ema is not None:
            mode, weight = self._ema._next_action()
            self._ema._fused_pending = True
        a = L.AdamWArgs(
            param=L.ptr(self._p.flat), grad=L.ptr(self._g.flat), exp_avg=L.ptr(self._m.flat), exp_avg_sq=L.ptr(self._v.flat),
            ema=L.ptr(self._ema_arena.flat) if self._ema is not None else None, param_bf16=L.ptr(self._bf16) if self._bf16 is not None else None,
            grad_sumsq=L.ptr(self._sumsq) if clip else None, numel=self._p.numel, step=self._t, lr=float(g0["lr"]), beta1=float(g0["betas"][0]),
            beta2=float(g0["betas"][1]), eps=float(g0["eps"]), weight_decay=float(g0["weight_decay"]),
            max_norm=float(self.max_grad_norm) if clip else 0.0, grad_scale=float(self._grad_scale), ema_weight=float(weight), ema_mode=int(mode), zero_grad=1,
        )
        L.check(self._lib.bsi_adamw_ema_step(ctypes.byref(a), st), "bsi_adamw_ema_step")
        self._p.generation += 1  # the kernel rewrote the parameters (and the EMA weights) behind torch's back
        if self._ema is not None and mode != 0:
            self._ema_arena.generation += 1
        self._last_scale, self._grad_scale = self._grad_scale, 1.0
        self._clean_version = self._g.flat._version  # the kernel left the gradients zeroed; autograd accumulation bumps the counter
        for p in self._params:
            self.state[p]["step"] += 1
        return loss

    def zero_grad(self, set_to_none: bool = True) -> None:
        """The gradients stay views of the arena (``set_to_none`` is ignored); ``step()`` has already zeroed them."""
        for i, p in enumerate(self._params):
            if p.grad is None or p.grad.data_ptr() != self._g.view(i).data_ptr():
                p.grad = self._g.view(i)
        if self._g.flat._version != self._clean_version:
            self._g.flat.zero_()
            self._clean_version = self._g.flat._version

    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)
        with torch.no_grad():
            for i, p in enumerate(self._params):
                s = self.state.get(p, {})
                for key, arena in (("exp_avg", self._m), ("exp_avg_sq", self._v)):
                    if key in s and s[key].data_ptr() != arena.view(i).data_ptr():
                        arena.view(i).copy_(s[key])
                        s[key] = arena.view(i)
                if "step" in s:
                    self._t = int(s["step"])
