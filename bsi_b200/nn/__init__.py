from .fourier_features import FourierFeatures  # noqa: F401
from .mlp import MLP  # noqa: F401
