"""CPU restatement of the optimizer side of the reference's training step.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import this module; the product
(``bsi_b200``) never does.  Pinned against golden vectors produced by the real thing (``tests/golden/make_golden.py optim``:
``torch.nn.utils.clip_grad_norm_`` + ``torch.optim.AdamW`` + the reference's own ``EMA`` class) in ``tests/test_oracle_golden.py``.

What the reference does after ``loss.backward()`` (one Lightning optimisation step):
  1. ``clip_grad_norm_(params, 1.0)``                      config/train.yaml:40 (Lightning ``gradient_clip_val``)
  2. ``torch.optim.AdamW(lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01).step()``   config/task/optimizer/adamw.yaml
  3. ``EMA.update()``                                      bsi/tasks/bsi.py:196-198 -> bsi/tasks/ema_pytorch.py:316-341
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import Tensor


@dataclass
class EMASchedule:
    """Hyper-parameters as ``create_ema`` passes them (bsi/tasks/bsi.py:73-81): ``power`` / ``inv_gamma`` of the yaml are
    swallowed by ``**kwargs`` and never reach ``EMA``, so the class defaults apply (ema_pytorch.py:86-92)."""

    beta: float = 0.9999
    update_after_step: int = 100
    update_every: int = 10
    inv_gamma: float = 1.0
    power: float = 2 / 3
    min_value: float = 0.0


def ema_current_decay(step: int, s: EMASchedule) -> float:
    """``EMA.get_current_decay`` (ema_pytorch.py:308-314); ``step`` is the counter *after* ``update()`` incremented it."""
    epoch = step - s.update_after_step - 1
    if epoch <= 0:
        return 0.0
    value = 1 - (1 + epoch / s.inv_gamma) ** -s.power
    return max(s.min_value, min(value, s.beta))


def ema_action(step: int, initted: bool, s: EMASchedule) -> tuple[str, float, int, bool]:
    """One ``EMA.update()`` (ema_pytorch.py:316-341) as data: returns (action, lerp weight, new step, new initted) with
    action in {"copy", "lerp", "none"}.  ``step`` is the counter before the call."""
    new_step = step + 1
    if not initted:
        return "copy", 1.0, new_step, True
    should_update = step % s.update_every == 0
    if should_update and step <= s.update_after_step:
        return "copy", 1.0, new_step, True
    if should_update:
        return "lerp", 1.0 - ema_current_decay(new_step, s), new_step, True
    return "none", 0.0, new_step, True


def clip_coef(grads: list[Tensor], max_norm: float) -> Tensor:
    """``torch.nn.utils.clip_grad_norm_`` (norm_type 2): norm of the per-tensor norms, coefficient clamped to 1."""
    norms = torch.stack([torch.linalg.vector_norm(g, 2) for g in grads])
    total = torch.linalg.vector_norm(norms, 2)
    return torch.clamp(max_norm / (total + 1e-6), max=1.0)


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, beta1: float, beta2: float, eps: float, wd: float) -> None:
    """``torch.optim.adamw._single_tensor_adamw`` (non-capturable, no amsgrad) on one tensor, in place."""
    p.mul_(1 - lr * wd)
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bias_correction1 = 1 - beta1**step
    bias_correction2 = 1 - beta2**step
    step_size = lr / bias_correction1
    denom = (v.sqrt() / bias_correction2**0.5).add_(eps)
    p.addcdiv_(m, denom, value=-step_size)


class OptimizerSide:
    """clip -> AdamW -> EMA over a list of tensors, state held like the reference holds it."""

    def __init__(self, params: list[Tensor], *, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_norm: float | None = 1.0,
                 ema: EMASchedule | None = None):
        self.params = params
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.t = 0
        self.ema_schedule = ema
        self.ema = [p.clone() for p in params] if ema is not None else None  # EMA.__init__ deep-copies the model
        self.ema_step, self.ema_initted = 0, False

    def step(self, grads: list[Tensor], coef: Tensor | None = None) -> None:
        """``coef`` overrides the clip coefficient (teacher forcing: torch's CPU fp32 norm of a million-element tensor is
        only good to ~1e-5, so large-arena parity tests inject the coefficient computed from the fp64 norm)."""
        grads = [g.clone() for g in grads]
        if self.max_norm is not None:
            c = clip_coef(grads, self.max_norm) if coef is None else coef
            for g in grads:
                g.mul_(c)
        self.t += 1
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            adamw_step(p, g, m, v, self.t, self.lr, *self.betas, self.eps, self.wd)
        if self.ema is not None:
            action, w, self.ema_step, self.ema_initted = ema_action(self.ema_step, self.ema_initted, self.ema_schedule)
            for e, p in zip(self.ema, self.params):
                if action == "copy":
                    e.copy_(p)
                elif action == "lerp":
                    e.lerp_(p, w)
