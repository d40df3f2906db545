"""Network parameter containers (names follow reference pyroved/nets)."""
from .fc import (fcClassifierNet, fcDecoderNet, fcEncoderNet, jfcEncoderNet,
                 sDecoderNet, fcRegressorNet, coord_latent, make_fc_layers)

from .conv import (convEncoderNet, convDecoderNet, FeatureExtractor, Upsampler, UpsampleBlock,
                   features_to_latent, latent_to_features)

__all__ = ["convEncoderNet", "convDecoderNet", "fcEncoderNet", "fcDecoderNet", "sDecoderNet", "fcRegressorNet",
           "fcClassifierNet", "jfcEncoderNet"]
