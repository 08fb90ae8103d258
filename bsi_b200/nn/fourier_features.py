"""Fourier-feature configuration object (reference bsi/nn/fourier_features.py:5-36).

Inside the native denoisers the features are generated on the fly by the patch-operand
kernel (bsi_dit_patch_operand) and never materialised; this module only carries the
exponent range (and the reference's non-persistent buffers).  Its own `forward` is a small
stand-alone utility with the reference's channel order (channel, frequency, {sin, cos}).
"""

import torch
from torch import Tensor, nn


class FourierFeatures(nn.Module):
    def __init__(self, *, n_min: int, n_max: int, **kwargs):
        super().__init__()
        self.n_min, self.n_max = n_min, n_max
        exponents = torch.arange(n_min, n_max + 1)
        self.register_buffer("coefs", 2 * torch.pi * 2**exponents, persistent=False)
        self.register_buffer("offsets", torch.tensor([0, torch.pi / 2]), persistent=False)

    def n_features(self) -> int:
        return len(self.coefs) * len(self.offsets)

    def forward(self, x: Tensor, *, dim: int) -> Tensor:
        assert dim >= 0, "Implementation expects a non-negative dimension index"
        trailing = x.dim() - dim - 1
        coefs = self.coefs.reshape(-1, 1, *([1] * trailing))
        offsets = self.offsets.reshape(-1, *([1] * trailing))
        phase = torch.addcmul(offsets, coefs, x.unsqueeze(dim + 1).unsqueeze(dim + 1))
        return phase.sin().flatten(start_dim=dim, end_dim=dim + 2)
