"""Training loops (names follow reference pyroved/trainers)."""
from .svi import SVItrainer
from .auxsvi import auxSVItrainer

__all__ = ['SVItrainer', 'auxSVItrainer']
