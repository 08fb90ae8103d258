"""Parameter containers with the reference's U-Net module tree and state_dict keys
(reference bsi/nn/residual_block.py, bsi/nn/attention.py, bsi/nn/simplified_unet.py).

The native DenoisingVDMUNet never calls their `forward`: convolutions, GroupNorm, modulation and attention run as
sm_100a kernels driven by the U-Net engine of libbsi_b200.so.  The classes exist so that reference checkpoints load
unchanged and default initialisation consumes the RNG exactly like the reference.
"""

from torch import nn

from .sequential import KwargsSequential

_NATIVE = "executed by the native U-Net engine (bsi_b200.models.DenoisingVDMUNet); this module only holds parameters"


class FeatureModulation(nn.Module):
    def forward(self, x, *, scale_shift):
        raise NotImplementedError(_NATIVE)


class Attention2D(nn.Module):
    def __init__(self, dim: int, *, heads: int = 4, padding_mode: str = "zeros"):
        super().__init__()
        self.heads = heads
        self.to_qkv = nn.Conv2d(dim, dim * 3, 3, padding=1, padding_mode=padding_mode)
        self.to_out = nn.Conv2d(dim, dim, 3, padding=1, padding_mode=padding_mode)

    def forward(self, x):
        raise NotImplementedError(_NATIVE)


class Residual(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x, *args, **kwargs):
        raise NotImplementedError(_NATIVE)


class ResidualBlock(nn.Module):
    def __init__(self, dim_in, dim_out, *, c_dim: int, ActFn, Norm, dropout, attention: bool = True, padding_mode: str = "zeros"):
        super().__init__()
        self.project_onto_scale_shift = nn.Linear(c_dim, dim_out * 2, 1)
        self.skip = nn.Conv2d(dim_in, dim_out, 1) if dim_in != dim_out else nn.Identity()
        self.layers = KwargsSequential(
            Norm(dim_in),
            ActFn(),
            nn.Conv2d(dim_in, dim_out, 3, padding=1, padding_mode=padding_mode),
            FeatureModulation(),
            ActFn(),
            *([nn.Dropout(dropout)] if dropout is not None else []),
            nn.Conv2d(dim_out, dim_out, 3, padding=1, padding_mode=padding_mode),
        )
        self.attention = attention
        self.res_attention = Residual(KwargsSequential(Norm(dim_out), Attention2D(dim_out, padding_mode=padding_mode))) if attention else nn.Identity()

    def forward(self, x, c):
        raise NotImplementedError(_NATIVE)


class SimplifiedUNet(nn.Module):
    def __init__(self, downsampling_blocks, upsampling_blocks, center_block):
        super().__init__()
        assert len(downsampling_blocks) == len(upsampling_blocks)
        as_list = lambda b: nn.ModuleList(b if isinstance(b, list) else [b])
        self.downsampling_blocks = nn.ModuleList([as_list(b) for b in downsampling_blocks])
        self.upsampling_blocks = nn.ModuleList([as_list(b) for b in upsampling_blocks])
        self.center_block = center_block

    def forward(self, x, *args, **kwargs):
        raise NotImplementedError(_NATIVE)
