from .dit import DenoisingDiT  # noqa: F401
from .pos_emb import NyquistPositionalEmbedding  # noqa: F401
from .vdm_unet import DenoisingVDMUNet  # noqa: F401
