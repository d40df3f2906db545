"""pyroved_b200: B200-native (sm_100a) implementation of pyroVED's SVI hot path.

Same Python surface as the reference for that path (`models.iVAE` ...,
`trainers.SVItrainer`, `.encode/.decode/.manifold2d`); the compute is
hand-written CUDA behind the C ABI in include/pvb.h.
"""
from . import models, trainers, nets, utils  # noqa: F401

__version__ = "0.1.0"
__all__ = ['models', 'trainers', 'nets', 'utils', '__version__']
