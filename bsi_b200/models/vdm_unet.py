"""DenoisingVDMUNet with the reference's constructor, module tree and state_dict layout
(reference bsi/models/vdm_unet.py:20-100), executed by the sm_100a U-Net engine of libbsi_b200.so:
NHWC activations, every 3x3 / 1x1 convolution an implicit GEMM on tcgen05 fed by shifted TMA boxes, GroupNorm+SiLU
operand kernels, scale/shift modulation and residual adds fused into the GEMM epilogues, flash attention for the
centre block.  Implemented configuration: dim = 128, one attention head, SiLU, zero padding, no downsampling attention
(config/experiment/cifar10-vdm.yaml:32-39)."""

from __future__ import annotations

from functools import partial

from torch import nn

from .. import _lib as L
from ..nn import Attention2D, FourierFeatures, KwargsSequential, Residual, ResidualBlock, SimplifiedUNet
from ._native import NativeDenoiser
from .pos_emb import NyquistPositionalEmbedding
from .utils import actfn_from_str


class DenoisingVDMUNet(NativeDenoiser):
    """U-Net structure as in the VDM paper without downsampling, native on B200."""

    _api = "bsi_unet_"

    def __init__(self, data_shape, pos_emb: NyquistPositionalEmbedding, actfn: str, dim: int, levels: int, pos_emb_mult: int,
                 n_attention_heads: int = 1, dropout: float | None = None, downsampling_attention: bool = False,
                 fourier_features: FourierFeatures | None = None, padding_mode: str = "zeros", **kwargs):
        super().__init__()
        self.data_shape = tuple(data_shape)
        assert len(self.data_shape) == 3, "Only works for 2D images"
        self.pos_emb = pos_emb
        self.fourier_features = fourier_features
        if actfn != "silu" or padding_mode != "zeros" or downsampling_attention:
            raise NotImplementedError("the native U-Net implements actfn='silu', padding_mode='zeros', downsampling_attention=False")
        channels = self.data_shape[0]
        in_features = channels + (channels * fourier_features.n_features() if fourier_features is not None else 0)
        ActFn = actfn_from_str(actfn)
        Norm = partial(nn.GroupNorm, 32)
        block = partial(ResidualBlock, ActFn=ActFn, Norm=Norm, dropout=dropout, attention=downsampling_attention, padding_mode=padding_mode)
        c_dim = pos_emb.size * pos_emb_mult
        self.pos_map = KwargsSequential(self.pos_emb, nn.Linear(pos_emb.size, c_dim), ActFn(), nn.Linear(c_dim, c_dim), ActFn())
        self.encode = nn.Conv2d(in_features, dim, 3, padding=1, padding_mode=padding_mode)
        self.decode = nn.Conv2d(dim, channels, 1)
        down = [block(dim, dim, c_dim=c_dim) for _ in range(levels)]
        up = [block(2 * dim, dim, c_dim=c_dim) for _ in range(levels)]
        center = KwargsSequential(
            block(dim, dim, c_dim=c_dim),
            Residual(KwargsSequential(Norm(dim), Attention2D(dim, heads=n_attention_heads, padding_mode=padding_mode))),
            block(dim, dim, c_dim=c_dim),
        )
        self.u_net = SimplifiedUNet(down, up, center)
        self._cfg = L.UnetConfig(channels, self.data_shape[1], self.data_shape[2], dim, levels, n_attention_heads, pos_emb.size, pos_emb_mult,
                                 fourier_features.n_min if fourier_features is not None else 0,
                                 fourier_features.n_max if fourier_features is not None else -1)
        self._init_native()

    def _named_tensors(self):
        yield from self.state_dict(keep_vars=True).items()
        yield "pos_emb.scale", self.pos_emb.scale
        yield "pos_emb.bias", self.pos_emb.bias
