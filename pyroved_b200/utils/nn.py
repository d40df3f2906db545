"""Small host-side helpers mirroring reference utils/nn.py."""
from copy import deepcopy

import torch
import torch.nn as nn


def set_deterministic_mode(seed: int) -> None:
    """reference utils/nn.py:87-100"""
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
        torch.cuda.manual_seed_all(seed)
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False


_ACTS = {"lrelu": nn.LeakyReLU, "tanh": nn.Tanh, "softplus": nn.Softplus, "relu": nn.ReLU,
         "gelu": nn.GELU}


def get_activation(activation):
    """Activation MODULE class; used as a structural placeholder so that
    state_dict keys match the reference's nn.Sequential (Linear at even
    indices).  The arithmetic runs in csrc/ kernels."""
    if activation is None:
        return None
    return _ACTS[activation]


def get_bnorm(dim):
    return {1: nn.BatchNorm1d, 2: nn.BatchNorm2d, 3: nn.BatchNorm3d}[dim]


def get_conv(dim):
    return {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}[dim]


def get_maxpool(dim):
    return {1: nn.MaxPool1d, 2: nn.MaxPool2d, 3: nn.MaxPool3d}[dim]


def to_onehot(idx: torch.Tensor, n: int) -> torch.Tensor:
    """reference utils/nn.py:37-48"""
    if torch.max(idx).item() >= n:
        raise AssertionError(
            "Labelling must start from 0 and "
            "maximum label value must be less than total number of classes")
    if idx.dim() == 1:
        idx = idx.unsqueeze(1)
    onehot = torch.zeros(idx.size(0), n)
    return onehot.scatter_(1, idx, 1)


def average_weights(ensemble):
    """Average state_dicts of same-architecture snapshots (utils/nn.py:11-34)."""
    lo = min(ensemble.keys())
    ensemble = {k - lo: v for k, v in ensemble.items()}
    avg = deepcopy(ensemble[0])
    for name in avg:
        if name.split("_")[-1] in ("mean", "var", "tracked"):
            continue
        stack = [m[name].detach().clone() for m in ensemble.values() if name in m]
        avg[name].copy_(sum(stack) / float(len(stack)))
    return avg


def broadcast_concat(input_args):
    """Concat with broadcasting over leading dims (reference Concat,
    utils/nn.py:51-74).  Shape bookkeeping only (views + one cat)."""
    if torch.is_tensor(input_args):
        return input_args
    args = [a.flatten(1) if a.ndim >= 4 else a for a in input_args]
    shape = torch.broadcast_shapes(*[a.shape[:-1] for a in args]) + (-1,)
    return torch.cat([a.expand(shape) for a in args], dim=-1)


class Concat(nn.Module):
    def __init__(self, allow_broadcast: bool = True):
        self.allow_broadcast = allow_broadcast
        super().__init__()

    def forward(self, input_args):
        return broadcast_concat(input_args)


def _to_device(input_data, **kwargs):
    device = kwargs.get("device", "cuda" if torch.cuda.is_available() else "cpu")
    if len(input_data) == 1:
        return input_data[0].to(device)
    return [t.to(device) for t in input_data]
