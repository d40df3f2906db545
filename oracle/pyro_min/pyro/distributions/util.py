"""broadcast_shape as used by reference utils/nn.py:8,71 (oracle-only)."""
import torch


def broadcast_shape(*shapes, **kwargs):
    return tuple(torch.broadcast_shapes(*[tuple(s) for s in shapes]))
