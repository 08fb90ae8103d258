from .fourier_features import FourierFeatures  # noqa: F401
from .mlp import MLP  # noqa: F401
from .sequential import KwargsSequential  # noqa: F401
from .unet_blocks import Attention2D, Residual, ResidualBlock, SimplifiedUNet  # noqa: F401
