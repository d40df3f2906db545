"""Variational autoencoder models (names follow reference pyroved/models)."""
from .ivae import iVAE

__all__ = ['iVAE']
