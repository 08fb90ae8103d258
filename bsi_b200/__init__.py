"""bsi_b200 — B200-native (sm_100a) implementation of the BSI sample / train_loss / elbo hot path.

Drop-in for the reference's `bsi.bsi` / `bsi.models.dit` import surface:

    from bsi_b200 import BSI, Discretization
    from bsi_b200.models import DenoisingDiT
    from bsi_b200.nn import FourierFeatures
"""

from .bsi import BSI, Discretization, LogUniform, broadcast_right  # noqa: F401

__version__ = "0.1.0"
