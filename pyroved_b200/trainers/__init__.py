"""Training loops (names follow reference pyroved/trainers)."""
from .svi import SVItrainer

__all__ = ['SVItrainer']
