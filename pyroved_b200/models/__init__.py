"""Variational autoencoder models (names follow reference pyroved/models)."""
from .ivae import iVAE
from .jivae import jiVAE
from .ssivae import ssiVAE
from .ss_reg_ivae import ss_reg_iVAE
from .ved import VED

__all__ = ['iVAE', 'jiVAE', 'ssiVAE', 'ss_reg_iVAE', 'VED']
