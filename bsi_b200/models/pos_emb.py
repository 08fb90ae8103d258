"""Sine/cosine embedding between 1/8 and a fraction of the Nyquist frequency
(reference bsi/models/pos_emb.py:42-84).  The native engine consumes the `scale`/`bias`
tables through bsi_time_embed; `forward` is the stand-alone utility."""

import numpy as np
import torch
from torch import Tensor, nn


class NyquistPositionalEmbedding(nn.Module):
    @classmethod
    def from_config(cls, size, expected_rate, **kwargs):
        return cls(size, expected_rate)

    def __init__(self, size: int, expected_rate: int):
        super().__init__()
        assert size % 2 == 0
        self.size = size
        pairs = size // 2
        top = (expected_rate / 2) / (1 + np.sqrt(5))  # Nyquist / (2 * golden ratio)
        freqs = np.geomspace(1 / 8, top, num=pairs)
        self.register_buffer("scale", torch.tensor(np.repeat(2 * np.pi * freqs, 2), dtype=torch.float32), persistent=False)
        self.register_buffer("bias", torch.tensor(np.tile([0.0, np.pi / 2], pairs), dtype=torch.float32), persistent=False)

    def forward(self, t: Tensor) -> Tensor:
        return torch.addcmul(self.bias, self.scale, t[..., None]).sin()
