"""Variational autoencoder models (names follow reference pyroved/models)."""
from .ivae import iVAE
from .jivae import jiVAE
from .ssivae import ssiVAE
from .ved import VED

__all__ = ['iVAE', 'jiVAE', 'ssiVAE', 'VED']
