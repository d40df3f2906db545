"""Variational autoencoder models (names follow reference pyroved/models)."""
from .ivae import iVAE
from .jivae import jiVAE
from .ssivae import ssiVAE

__all__ = ['iVAE', 'jiVAE', 'ssiVAE']
