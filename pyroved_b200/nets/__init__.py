"""Network parameter containers (names follow reference pyroved/nets)."""
from .fc import (fcClassifierNet, fcDecoderNet, fcEncoderNet, jfcEncoderNet,
                 sDecoderNet, fcRegressorNet, coord_latent, make_fc_layers)

__all__ = ["fcEncoderNet", "fcDecoderNet", "sDecoderNet", "fcRegressorNet",
           "fcClassifierNet", "jfcEncoderNet"]
