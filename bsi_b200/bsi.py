"""Host-side mirror of the reference's `bsi/bsi.py` API on top of the sm_100a kernels.

Same names, arguments, return structures and error behaviour as the reference
(`BSI`, `Discretization`, `LogUniform`, `broadcast_right`; reference bsi/bsi.py:12-445), but
every full-tensor operation is one call into libbsi_b200.so (include/bsi_b200.h):

    sample / sample_history   -> bsi_sample_init, bsi_step_fused (+ CUDA graph of the step)
    _predict_x                -> bsi_scale_rows, bsi_edm_combine
    _sample_q_mu_lambda       -> bsi_q_sample
    reconstruction_loss       -> bsi_recon_reduce  (x_hat never materialised)
    *_measurement_loss, train_loss -> bsi_sqerr_reduce (+ bsi_sqerr_backward for autograd)
    Discretization.bucketize  -> bsi_bucketize

Only O(batch) scalar bookkeeping (lambda grid, EDM coefficients) stays in torch ops, in the
reference's op order so that the per-sample coefficients are bit-identical.
There is no CPU path: tensors must live on a CUDA device.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Literal

import torch
from torch import Tensor, nn

from . import _lib as L

__all__ = ["BSI", "Discretization", "LogUniform", "ReplayNoise", "broadcast_right"]


def _cuda_f32(x: Tensor, what: str) -> Tensor:
    if not x.is_cuda:
        raise L.BsiNativeError(f"{what} must be a CUDA tensor: bsi_b200 has no CPU fallback")
    if x.dtype != torch.float32:
        raise L.BsiNativeError(f"{what} must be float32 (got {x.dtype})")
    return x.contiguous()


# ----------------------------------------------------------------------------- discretisation
@dataclass
class Discretization:
    """k right-open bins on [min, max] centred on the k grid values (reference bsi/bsi.py:12-58)."""

    min: float
    max: float
    k: int

    @classmethod
    def image_8bit(cls):
        return cls(-1.0, 1.0, 256)

    @property
    def dx(self) -> float:
        return (self.max - self.min) / (self.k - 1)

    @property
    def range(self) -> tuple[float, float]:
        half = self.dx / 2
        return (self.min - half, self.max + half)

    def bin_boundaries(self, device: torch.device, dtype: torch.dtype) -> Tensor:
        lo, hi = self.range
        return torch.linspace(lo, hi, self.k + 1, device=device, dtype=dtype)

    def bucketize(self, x: Tensor) -> Tensor:
        """int64 bin index of every element, bit-exact with the reference's fp32 op order."""
        x = _cuda_f32(x, "bucketize input")
        out = torch.empty(x.shape, dtype=torch.int64, device=x.device)
        with torch.cuda.device(x.device):
            L.check(
                L.load().bsi_bucketize(L.ptr(x), L.ptr(out), None, self.range[0], self.dx, self.k, x.numel(), L.stream_ptr(x.device)),
                "bsi_bucketize",
            )
        return out

    def bucketize_u8(self, x: Tensor) -> Tensor:
        """Same index as `bucketize`, stored in 8 bits (k <= 256)."""
        x = _cuda_f32(x, "bucketize input")
        out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            L.check(
                L.load().bsi_bucketize(L.ptr(x), None, L.ptr(out), self.range[0], self.dx, self.k, x.numel(), L.stream_ptr(x.device)),
                "bsi_bucketize",
            )
        return out

    def to_unit_interval(self, x: Tensor) -> Tensor:
        return (x - self.min) / (self.max - self.min)

    def to_8bit_image(self, data: Tensor) -> Tensor:
        if data.is_cuda and data.dtype == torch.float32:  # one fused pass instead of four elementwise kernels
            data = data.contiguous()
            out = torch.empty(data.shape, dtype=torch.uint8, device=data.device)
            with torch.cuda.device(data.device):
                L.check(L.load().bsi_to_uint8(L.ptr(out), L.ptr(data), self.min, self.max, data.numel(), L.stream_ptr(data.device)), "bsi_to_uint8")
            return out
        return (self.to_unit_interval(data) * 255).clamp(0, 255).to(torch.uint8)


def broadcast_right(x: Tensor, other: Tensor) -> Tensor:
    """View `x` with trailing singleton dims so it broadcasts against `other` (reference bsi/bsi.py:61-64)."""
    assert other.ndim >= x.ndim
    return x.reshape(x.shape + (1,) * (other.ndim - x.ndim))


class LogUniform:
    """Log-uniform law of the belief precision lambda on [low, high] (reference bsi/bsi.py:67-84)."""

    def __init__(self, low, high):
        self.low, self.high = low, high
        # math.log of the (fp32-rounded) 0-d tensors, exactly like the reference
        self.ln_low, self.ln_high = math.log(low), math.log(high)
        self.diff_ln_high_ln_low = self.ln_high - self.ln_low

    def reciprocal_pdf(self, value: Tensor) -> Tensor:
        return value * self.diff_ln_high_ln_low

    def cdf(self, value: Tensor) -> Tensor:
        return (torch.log(value) - self.ln_low) / self.diff_ln_high_ln_low

    def icdf(self, quantile: Tensor) -> Tensor:
        return torch.exp(self.diff_ln_high_ln_low * quantile + self.ln_low)


class ReplayNoise:
    """Stands in for the ``generator`` argument and hands out PRE-DRAWN tensors in the order the reference consumes its
    generator (``randn`` / ``rand`` / ``randperm`` / ``randint`` calls of bsi/bsi.py:224-226,264,283-284,325,332,415,430-434).

    Parity hook: the stored draws of a reference run (tests/golden/full.pt) are injected into the CUDA path, so losses
    and bits-per-dim are compared on exactly the noise and lambda grid the reference used."""

    def __init__(self, draws):
        self._draws = list(draws)

    def take(self, shape, device) -> Tensor:
        if not self._draws:
            raise RuntimeError("ReplayNoise exhausted: the call consumed more random draws than were recorded")
        d = self._draws.pop(0)
        if tuple(d.shape) != tuple(shape):
            raise RuntimeError(f"ReplayNoise: next recorded draw has shape {tuple(d.shape)}, the call asked for {tuple(shape)}")
        return d.to(device)

    @property
    def remaining(self) -> int:
        return len(self._draws)


def _randn(shape, generator, **tensor_args) -> Tensor:
    if isinstance(generator, ReplayNoise):
        return generator.take(shape, tensor_args["device"]).to(tensor_args["dtype"]).contiguous()
    return torch.randn(shape, **tensor_args, generator=generator)


# ----------------------------------------------------------------------------- autograd bridge
class _SquaredError(torch.autograd.Function):
    """sum_d (x - (c_skip*mu + c_out*f))^2 per row with a native backward w.r.t. f."""

    @staticmethod
    def forward(ctx, f, x, mu, c_skip, c_out, B):
        R, D = mu.shape[0], mu[0].numel()
        out = torch.empty(R, dtype=torch.float32, device=f.device)
        with torch.cuda.device(f.device):
            L.check(
                L.load().bsi_sqerr_reduce(L.ptr(out), L.ptr(x), L.ptr(mu), L.ptr(f), L.ptr(c_skip), L.ptr(c_out), R, B, D, L.stream_ptr(f.device)),
                "bsi_sqerr_reduce",
            )
        ctx.save_for_backward(f, x, mu, c_skip, c_out)
        ctx.B = B
        return out

    @staticmethod
    def backward(ctx, grad_out):
        f, x, mu, c_skip, c_out = ctx.saved_tensors
        R, D = mu.shape[0], mu[0].numel()
        grad_f = torch.empty_like(f)
        w = grad_out.contiguous().float()
        with torch.cuda.device(f.device):
            L.check(
                L.load().bsi_sqerr_backward(
                    L.ptr(grad_f), L.ptr(w), L.ptr(x), L.ptr(mu), L.ptr(f), L.ptr(c_skip), L.ptr(c_out), R, ctx.B, D, L.stream_ptr(f.device)
                ),
                "bsi_sqerr_backward",
            )
        return grad_f, None, None, None, None, None


# ----------------------------------------------------------------------------- the algorithm
class BSI(nn.Module):
    """Bayesian Sample Inference (arXiv 2502.07580) with the reference's public surface.

    Extra, optional attributes (not in the reference):
      noise_source: "philox" (default; in-kernel counter-based noise keyed by a seed drawn from
          `generator`) or "torch" (draw torch.randn from `generator` in the reference's exact
          order and inject it — the parity mode).
    """

    def __init__(
        self,
        model: nn.Module,
        *,
        data_shape: tuple[int, ...],
        lambda_0: float,
        alpha_M: float,
        alpha_R: float,
        k: int,
        preconditioning: Literal["edm"] | None,
        low_discrepancy_sampling: bool = True,
        discretization: Discretization | None = None,
    ):
        super().__init__()
        self._model = [model]  # kept out of the module tree: empty state_dict, like the reference
        self.data_shape = tuple(data_shape)
        for name, value in (("lambda_0", lambda_0), ("alpha_R", alpha_R), ("alpha_M", alpha_M)):
            self.register_buffer(name, torch.as_tensor(value), persistent=False)
        self.k = k
        self.preconditioning = preconditioning
        self.low_discrepancy_sampling = low_discrepancy_sampling
        self.discretization = discretization
        self.p_lambda = LogUniform(self.lambda_0, self.lambda_0 + self.alpha_M)
        self.register_buffer("default_schedule", torch.linspace(0.0, 1.0, self.k + 1), persistent=False)
        self.noise_source: str = "philox"
        self._plans: dict = {}

    # ---- plumbing -----------------------------------------------------------------------
    @property
    def model(self):
        return self._model[0]

    def set_model(self, model):
        self._model[0] = model
        self._plans.clear()  # cached graphs / packed state refer to the old denoiser

    @property
    def tensor_args(self):
        return {"device": self.lambda_0.device, "dtype": self.lambda_0.dtype}

    @property
    def _numel(self) -> int:
        return math.prod(self.data_shape)

    def _require_cuda(self):
        dev = self.lambda_0.device
        if dev.type != "cuda":
            raise L.BsiNativeError("BSI buffers are on %s: move the module to a CUDA device (no CPU fallback)" % dev)
        if self.lambda_0.dtype != torch.float32:
            raise L.BsiNativeError("bsi_b200 keeps the belief state in float32; got %s" % self.lambda_0.dtype)
        return dev

    def _check_precond(self):
        if self.preconditioning not in (None, "edm"):
            raise RuntimeError(f"Unknown preconditioning {self.preconditioning}")

    def _noise(self, shape, generator, draw: int, sample_base: int = 0, seed: int | None = None):
        """bsi_noise for one draw: injected torch.randn (parity) or Philox keyed by `seed`."""
        if self.noise_source == "torch" or isinstance(generator, ReplayNoise):
            eps = _randn(shape, generator, **self.tensor_args)
            return L.noise(eps=eps), eps
        return L.noise(seed=seed, sample_base=sample_base, draw=draw), None

    def _philox(self, generator) -> bool:
        return self.noise_source != "torch" and not isinstance(generator, ReplayNoise)

    def _draw_seed(self, generator) -> int:
        dev = generator.device if generator is not None else "cpu"
        return int(torch.randint(0, 2**62, (), generator=generator, device=dev, dtype=torch.int64).item())

    def _native(self):
        """The native denoiser behind ``self.model``: the model itself, or the ``ema_model`` of an EMA wrapper whose call just
        delegates to it (bsi/tasks/ema_pytorch.py:436-437; ``BSITraining`` builds its evaluation ``BSI`` around that wrapper,
        bsi/tasks/bsi.py:115-118) -- so sampling / ELBO with EMA weights also take the fused native path."""
        m = self.model
        for _ in range(3):  # EMA(ema_model=...) and DistributedDataParallel(module=...) (bsi/tasks/bsi.py:163-166) only delegate
            if getattr(m, "bsi_native", False):
                return m
            m = getattr(m, "ema_model", None) or getattr(m, "module", None)
            if m is None:
                return None
        return None

    def _is_native_denoiser(self) -> bool:
        return self._native() is not None

    # ---- EDM preconditioning (reference bsi/bsi.py:375-403) ---------------------------------
    def _edm_preconditioning(self, t: Tensor):
        lam = self.p_lambda.icdf(t)
        a = lam - self.lambda_0
        kappa = 1 + a * (a / lam)  # avoids squaring alpha
        return a / kappa, torch.rsqrt(kappa), torch.sqrt(lam / kappa)

    def _denoise(self, mu: Tensor, t: Tensor, c_in: Tensor | None) -> Tensor:
        """f = model(c_in * mu, t); the scaling is fused into the native DiT's operand builder when possible."""
        dev = mu.device
        if self._is_native_denoiser() and not torch.is_grad_enabled():
            return self._native().forward_scaled(mu, t, c_in)
        if c_in is None:
            return self.model(mu, t)
        scaled = torch.empty_like(mu)
        c_in = c_in.contiguous()  # held in a local until the launch is enqueued
        with torch.cuda.device(dev):
            L.check(
                L.load().bsi_scale_rows(L.ptr(scaled), L.ptr(mu), L.rowref(c_in, 1), None, mu.shape[0], mu[0].numel(), L.stream_ptr(dev)),
                "bsi_scale_rows",
            )
        return self.model(scaled, t)

    def _predict_x(self, mu: Tensor, t: Tensor) -> Tensor:
        self._check_precond()
        self._require_cuda()
        mu = _cuda_f32(mu, "mu")
        if self.preconditioning is None:
            return self.model(mu, t)
        c_skip, c_out, c_in = self._edm_preconditioning(t)
        f = self._denoise(mu, t, c_in)
        f = _cuda_f32(f.detach() if not f.requires_grad else f, "denoiser output")
        if f.requires_grad:  # differentiable combine for callers that backprop through _predict_x
            return torch.addcmul(broadcast_right(c_skip, mu) * mu, broadcast_right(c_out, mu), f)
        x_hat = torch.empty_like(mu)
        c_skip, c_out = c_skip.contiguous(), c_out.contiguous()
        with torch.cuda.device(mu.device):
            L.check(
                L.load().bsi_edm_combine(
                    L.ptr(x_hat), L.ptr(mu), L.ptr(f), L.rowref(c_skip, 1), L.rowref(c_out, 1), None,
                    mu.shape[0], mu[0].numel(), L.stream_ptr(mu.device),
                ),
                "bsi_edm_combine",
            )  # fmt: skip
        return x_hat

    # ---- forward process (reference bsi/bsi.py:405-445) --------------------------------------
    def _sample_q_mu_lambda(self, x: Tensor, lambda_: Tensor, generator=None, *, _c_in: Tensor | None = None, _seed=None, _draw=0,
                            _shard: tuple[int, int] | None = None):
        """mu ~ q(mu | x, lambda) for lambda of shape [..., batch]; returns [..., batch, *data_shape].

        _shard = (start, batch_total): `x` / `lambda_` are the columns [start, start + len(x)) of a batch of `batch_total` data
        points evaluated on another rank layout; noise is keyed by the row index of the FULL [..., batch_total] problem (or the
        full tensor is drawn and sliced in the injected-noise modes), so every shard reproduces the single-GPU values."""
        dev = self._require_cuda()
        x = _cuda_f32(x, "x")
        B, D = x.shape[0], self._numel
        R = lambda_.numel()
        lam = lambda_.reshape(-1)
        gamma = ((lam - self.lambda_0) / lam).contiguous()
        sigma = torch.rsqrt(lam).contiguous()
        if x.ndim < 1 or lambda_.ndim < 1 or lambda_.shape[-1] != B:
            # the reference would broadcast a mismatched lambda inside torch (or fail there); the kernels index x[r % B] and the
            # per-row coefficient vectors directly, so a shape that does not end in the batch axis is rejected here
            raise ValueError(f"lambda_ must have shape [..., batch={B}], got {tuple(lambda_.shape)}")
        if _c_in is not None and _c_in.numel() != R:
            raise ValueError(f"c_in has {_c_in.numel()} elements for {R} (sample, data point) rows")
        if self._philox(generator) and _seed is None:
            _seed = self._draw_seed(generator)
        if R == 0:  # empty batch: nothing to draw (torch.randn of an empty shape consumes no random numbers either)
            mu = torch.empty((*lambda_.shape, *self.data_shape), **self.tensor_args)
            return (mu, torch.empty_like(mu)) if _c_in is not None else mu
        mu = torch.empty((*lambda_.shape, *self.data_shape), **self.tensor_args)
        model_in = torch.empty_like(mu) if _c_in is not None else None
        c_in = _c_in.contiguous() if _c_in is not None else None
        lib, st = L.load(), L.stream_ptr(dev)
        if _shard is None or (_shard[0] == 0 and _shard[1] == B):
            nz, _keep = self._noise((*lambda_.shape, *self.data_shape), generator, _draw, 0, _seed)
            with torch.cuda.device(dev):
                L.check(lib.bsi_q_sample(L.ptr(mu), L.ptr(model_in), L.ptr(x), L.ptr(gamma), L.ptr(sigma), L.ptr(c_in), nz, R, B, D, st), "bsi_q_sample")
            return (mu, model_in) if _c_in is not None else mu
        # a column shard of the [..., batch_total] problem: one launch per leading row, keyed by the full problem's row index
        start, total = _shard
        lead = R // B
        eps_full = None
        if not self._philox(generator):
            eps_full = _randn((*lambda_.shape[:-1], total, *self.data_shape), generator, **self.tensor_args).reshape(lead, total, D)
        off4 = lambda t, i: None if t is None else t.data_ptr() + 4 * i
        with torch.cuda.device(dev):
            for i in range(lead):
                if eps_full is not None:
                    eps_i = eps_full[i, start : start + B].contiguous()
                    nz = L.noise(eps=eps_i)
                else:
                    nz = L.noise(seed=_seed, sample_base=i * total + start, draw=_draw)
                L.check(lib.bsi_q_sample(off4(mu, i * B * D), off4(model_in, i * B * D), L.ptr(x), off4(gamma, i * B), off4(sigma, i * B), off4(c_in, i * B),
                                         nz, B, B, D, st), "bsi_q_sample")
        return (mu, model_in) if _c_in is not None else mu

    def _sample_lambda(self, n_samples: int, batch_size: int, generator=None) -> Tensor:
        if self.low_discrepancy_sampling:
            # low-discrepancy grid of the VDM paper: one random offset, a permuted regular grid
            total = n_samples * batch_size
            if isinstance(generator, ReplayNoise):
                dev = self.tensor_args["device"]
                offset, perm = generator.take((), dev).to(self.tensor_args["dtype"]), generator.take((total,), dev)
            else:
                offset = torch.rand((), **self.tensor_args, generator=generator)
                perm = torch.randperm(total, device=self.tensor_args["device"], generator=generator)
            grid = perm / (1 + total)
            t = torch.remainder(grid.reshape(n_samples, batch_size) + offset, 1)
        else:
            # the reference draws the transposed shape here (bsi/bsi.py:441-445); mirrored as is.  Only n_samples == batch_size
            # survives the shape checks of the loss methods below (the reference broadcasts or fails inside torch instead)
            t = torch.rand((batch_size, n_samples), **self.tensor_args, generator=generator)
        return self.p_lambda.icdf(t)

    # ---- losses (reference bsi/bsi.py:152-310) ------------------------------------------------
    def _errors(self, x: Tensor, lambda_: Tensor, t_flat: Tensor, generator, draw: int, _shard=None) -> Tensor:
        """sum_d (x - x_hat)^2 for mu ~ q(.|x, lambda_[n,B]) and the model evaluated at t_flat -> [n,B]."""
        self._check_precond()
        if lambda_.ndim != 2 or lambda_.shape[1] != len(x) or t_flat.numel() != lambda_.numel():
            raise ValueError(
                f"lambda_ must be [n_samples, batch={len(x)}] with one time per entry, got lambda_ {tuple(lambda_.shape)} and "
                f"{t_flat.numel()} times (low_discrepancy_sampling=False draws the reference's transposed [batch, n_samples] grid, "
                "bsi/bsi.py:441-445, which only fits when n_samples == batch)"
            )
        n, B = lambda_.shape
        if B == 0:
            return _cuda_f32(x, "x").new_zeros((n, 0))
        if self.preconditioning == "edm":
            c_skip, c_out, c_in = self._edm_preconditioning(t_flat)
            mu, model_in = self._sample_q_mu_lambda(x, lambda_, generator, _c_in=c_in, _draw=draw, _shard=_shard)
        else:
            mu = self._sample_q_mu_lambda(x, lambda_, generator, _draw=draw, _shard=_shard)
            model_in, c_skip, c_out = mu, torch.zeros_like(t_flat), torch.ones_like(t_flat)
        mu_f, in_f = mu.flatten(end_dim=1), model_in.flatten(end_dim=1)
        f = _cuda_f32(self.model(in_f, t_flat), "denoiser output")
        err = _SquaredError.apply(f, _cuda_f32(x, "x"), mu_f, c_skip.contiguous(), c_out.contiguous(), B)
        return err.reshape(n, B)

    def reconstruction_loss(self, x: Tensor, n_samples: int, generator=None, *, _shard=None) -> Tensor:
        self._check_precond()
        dev = self._require_cuda()
        x = _cuda_f32(x, "x")
        B, D = x.shape[0], self._numel
        if B == 0:
            return x.new_zeros((n_samples, 0))
        lam_M = x.new_full((n_samples, B), float(self.lambda_0 + self.alpha_M))
        t_one = x.new_ones(n_samples * B)
        if self.preconditioning == "edm":
            c_skip, c_out, c_in = self._edm_preconditioning(t_one)
            mu, model_in = self._sample_q_mu_lambda(x, lam_M, generator, _c_in=c_in, _draw=0, _shard=_shard)
        else:
            mu = self._sample_q_mu_lambda(x, lam_M, generator, _draw=0, _shard=_shard)
            model_in, c_skip, c_out = mu, torch.zeros_like(t_one), torch.ones_like(t_one)
        f = _cuda_f32(self.model(model_in.flatten(end_dim=1), t_one).detach(), "denoiser output")
        mu_f = mu.flatten(end_dim=1)
        if self.discretization is None:
            # continuous log-likelihood (reference bsi/bsi.py:234-235): -log N(x; x_hat, 1/alpha_R) summed over dims
            err = _SquaredError.apply(f, x, mu_f, c_skip.contiguous(), c_out.contiguous(), B).reshape(n_samples, B)
            log_norm = 0.5 * (math.log(2 * math.pi) - torch.log(self.alpha_R))
            return 0.5 * self.alpha_R * err + D * log_norm
        disc = self.discretization
        edges = disc.bin_boundaries(dev, torch.float32)
        inv_scale = float(torch.rsqrt(self.alpha_R).reciprocal())
        out = torch.empty(n_samples * B, **self.tensor_args)
        c_skip, c_out = c_skip.contiguous(), c_out.contiguous()
        with torch.cuda.device(dev):
            L.check(
                L.load().bsi_recon_reduce(
                    L.ptr(out), L.ptr(x), L.ptr(mu_f), L.ptr(f), L.ptr(c_skip), L.ptr(c_out), L.ptr(edges),
                    disc.k, disc.range[0], disc.dx, inv_scale, n_samples * B, B, D, L.stream_ptr(dev),
                ),
                "bsi_recon_reduce",
            )  # fmt: skip
        return out.reshape(n_samples, B)

    def _shard_columns(self, grid: Tensor, x: Tensor, _shard) -> Tensor:
        """Columns of this shard out of a [n, batch_total] grid drawn for the whole batch (draw once, slice)."""
        if _shard is None:
            return grid
        start, total = _shard
        assert grid.shape[1] == total and start + len(x) <= total
        return grid[:, start : start + len(x)].contiguous()

    def inf_measurement_loss(self, x: Tensor, n_samples: int, generator=None, *, _shard=None) -> Tensor:
        lambda_ = self._shard_columns(self._sample_lambda(n_samples, _shard[1] if _shard else len(x), generator), x, _shard)
        t = self.p_lambda.cdf(lambda_).flatten(end_dim=1)
        err = self._errors(x, lambda_, t, generator, draw=1, _shard=_shard)
        return 0.5 * self.p_lambda.reciprocal_pdf(lambda_) * err

    def finite_measurement_loss(self, x: Tensor, n_samples: int, generator=None, *, t: Tensor | None = None, _shard=None) -> Tensor:
        if t is None:
            t = self.default_schedule
        lambda_ = self.p_lambda.icdf(t)
        alpha = lambda_.diff()
        k = len(alpha)
        total = _shard[1] if _shard else len(x)
        if isinstance(generator, ReplayNoise):
            i = generator.take((n_samples, total), x.device)
        else:
            i = torch.randint(0, k, (n_samples, total), device=x.device, generator=generator)
        i = self._shard_columns(i, x, _shard)
        err = self._errors(x, lambda_[i], t[i].flatten(end_dim=1), generator, draw=1, _shard=_shard)
        return (0.5 * k) * alpha[i] * err

    def _assemble_elbo(self, l_recon: Tensor, l_measure: Tensor, estimate_var: bool):
        elbo = -(l_recon.mean(dim=0) + l_measure.mean(dim=0))
        to_bpd = -1 / (math.log(2) * self._numel)
        extra = {"l_recon": l_recon, "l_measure": l_measure}
        if estimate_var:
            assert l_recon.shape[0] > 1 and l_measure.shape[0] > 1, "Need at least two samples of each to estimate variance"
            var = l_recon.var(dim=0, unbiased=True) / l_recon.shape[0] + l_measure.var(dim=0, unbiased=True) / l_measure.shape[0]
            extra["bpd_var"] = (to_bpd**2) * var
        return elbo, to_bpd * elbo, extra

    def elbo(self, x: Tensor, n_recon_samples: int, n_measure_samples: int, generator=None, *, estimate_var: bool = False, _shard=None):
        """Monte-Carlo estimate of the infinite-step ELBO: returns (elbo[B], bpd[B], extra).

        _shard = (start, batch_total) (extension, used by bsi_b200.distributed.sharded_elbo): `x` is a slice of a larger batch;
        all random draws are made for the whole batch from `generator` and sliced, so the slice's result equals the corresponding
        entries of the single-GPU call with the same generator state."""
        l_recon = self.reconstruction_loss(x, n_recon_samples, generator, _shard=_shard)
        l_measure = self.inf_measurement_loss(x, n_measure_samples, generator, _shard=_shard)
        return self._assemble_elbo(l_recon, l_measure, estimate_var)

    def finite_elbo(self, x: Tensor, n_recon_samples: int, n_measure_samples: int, generator=None, *, t: Tensor | None = None, estimate_var: bool = False,
                    _shard=None):
        l_recon = self.reconstruction_loss(x, n_recon_samples, generator, _shard=_shard)
        l_measure = self.finite_measurement_loss(x, n_measure_samples, generator, t=t, _shard=_shard)
        return self._assemble_elbo(l_recon, l_measure, estimate_var)

    def train_loss(self, x: Tensor, generator=None) -> Tensor:
        """lambda * dln * mean_d (x - x_hat)^2 per data point; differentiable w.r.t. the denoiser."""
        lambda_ = self._sample_lambda(1, len(x), generator)
        t = self.p_lambda.cdf(lambda_[0])
        err = self._errors(x, lambda_, t, generator, draw=0)[0]
        return self.p_lambda.reciprocal_pdf(lambda_[0]) * (err / self._numel)

    # ---- sampler (reference bsi/bsi.py:312-373) -------------------------------------------------
    def _step_table(self, t: Tensor):
        """Per-step scalars, tabulated with the reference's torch ops: coef[k+1, 8], c_in[k+1], t_rows[k+1].

        Rows 0..k-1 belong to the k updates; row k is the final prediction at t = 1."""
        lam = self.p_lambda.icdf(t)
        alpha = lam.diff()
        k = len(alpha)
        t_rows = torch.cat((t[:k], t.new_ones(1)))
        if self.preconditioning == "edm":
            c_skip, c_out, c_in = self._edm_preconditioning(t_rows)
        else:
            c_skip, c_out, c_in = torch.zeros_like(t_rows), torch.ones_like(t_rows), torch.ones_like(t_rows)
        coef = torch.zeros((k + 1, 8), **self.tensor_args)
        coef[:, 0], coef[:, 1] = c_skip, c_out
        coef[:k, 2], coef[:k, 3], coef[:k, 4], coef[:k, 5] = torch.rsqrt(alpha), alpha, lam[:k], lam[1:]
        return k, lam, coef.contiguous(), c_in.contiguous(), t_rows.contiguous()

    @torch.no_grad()
    def _run_sampler(self, n: int, generator, t: Tensor | None, history: bool, sample_offset: int = 0, seed: int | None = None):
        self._check_precond()
        dev = self._require_cuda()
        t = self.default_schedule if t is None else t.to(**self.tensor_args)
        k, lam, coef, c_in, t_rows = self._step_table(t)
        D = self._numel
        lib, st = L.load(), L.stream_ptr(dev)
        philox = self._philox(generator)
        if philox and seed is None:
            seed = self._draw_seed(generator)
        shape = (n, *self.data_shape)
        if n == 0:  # the reference returns empty tensors (every op of its loop accepts an empty batch)
            empty = torch.empty(shape, **self.tensor_args)
            if history:
                return empty.new_empty((k + 1, *shape)), empty.new_empty((k + 1, *shape)), empty.new_empty((k, *shape))
            return empty
        precond = 1 if self.preconditioning == "edm" else 0
        sigma0 = torch.rsqrt(lam[:1]).contiguous()
        if self._is_native_denoiser() and philox and not history:
            # whole loop on the device: CUDA graph of (denoiser forward + fused step), replayed k times
            return self._native().sample_loop(n, sigma0, coef, c_in, t_rows, k, seed, sample_offset, precond, plans=self._plans)
        mu = torch.empty(shape, **self.tensor_args)
        with torch.cuda.device(dev):
            nz, _keep = self._noise(shape, generator, 0, sample_offset, seed)
            L.check(lib.bsi_sample_init(L.ptr(mu), L.ptr(sigma0), nz, n, D, st), "bsi_sample_init")
            if history:
                mus = torch.empty((k + 1, *shape), **self.tensor_args)
                x_hats = torch.empty((k + 1, *shape), **self.tensor_args)
                ys = torch.empty((k, *shape), **self.tensor_args)
                mus[0].copy_(mu)
            for i in range(k):
                ti = t_rows[i].expand(n)
                f = _cuda_f32(self._denoise(mu, ti, c_in[i].expand(n) if precond else None), "denoiser output")
                # Philox counter of step i is draw + step = 1 + i, the same as in the captured graph (draw = 1, device step counter)
                nz, _keep = self._noise(shape, generator, 1, sample_offset, seed)
                L.check(
                    lib.bsi_step_fused(
                        L.ptr(mu), L.ptr(f), L.ptr(coef), None, i, precond, nz,
                        L.ptr(x_hats[i]) if history else None, L.ptr(ys[i]) if history else None, n, D, st,
                    ),
                    "bsi_step_fused",
                )  # fmt: skip
                if history:
                    mus[i + 1].copy_(mu)
            final = self._predict_x(mu, mu.new_ones(n))
        if history:
            x_hats[k].copy_(final)
            return mus, x_hats, ys
        return final

    def sample(self, n_samples: int, generator=None, *, t: Tensor | None = None, sample_offset: int = 0, seed: int | None = None) -> Tensor:
        """Draw `n_samples` samples with the k-step sampler; `t` overrides the default schedule.

        sample_offset / seed (extensions): global index of the first sample and Philox key, so a
        batch sharded over several GPUs reproduces the single-GPU result."""
        return self._run_sampler(n_samples, generator, t, False, sample_offset, seed)

    def sample_history(self, n_samples: int, generator=None, *, t: Tensor | None = None):
        """Like `sample` but returns (mus[k+1], x_hats[k+1], ys[k]) of every step."""
        return self._run_sampler(n_samples, generator, t, True)
