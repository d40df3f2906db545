"""Fully-connected encoder / decoder parameter containers.

Same constructor signatures, attribute names and state_dict keys as reference
nets/fc.py (so `.pt` checkpoints interchange), but the modules only OWN the
parameters: `forward` runs hand-written CUDA kernels through the C ABI
(inference, no autograd); training runs in `pyroved_b200.engine`.
"""
from typing import List, Tuple

import torch
import torch.nn as nn

from .. import ops
from ..utils.nn import get_activation, Concat, broadcast_concat


def _prod(t):
    n = 1
    for v in t:
        n *= int(v)
    return n


def _check_dim(d):
    if len(d) not in [1, 2, 3]:
        raise ValueError("in_dim must be (h, w), (h, w, c), or (l,)")


def make_fc_layers(in_dim: int, hidden_dim: List[int], activation: str = "tanh") -> nn.Sequential:
    """[Linear, act, Linear, act, ...] -- Linear modules sit at even indices as
    in the reference (nets/fc.py:307-324)."""
    hidden_dim = list(hidden_dim)
    dims = [in_dim] + hidden_dim
    mods = []
    for i in range(len(hidden_dim)):
        mods += [nn.Linear(dims[i], dims[i + 1]), get_activation(activation)()]
    return nn.Sequential(*mods)


def fc_stack_forward(fc_layers: nn.Sequential, x: torch.Tensor, activation: str) -> torch.Tensor:
    for m in fc_layers:
        if isinstance(m, nn.Linear):
            x = ops.linear_fwd(x, m.weight.data, m.bias.data, activation)
    return x


def linear_layers(fc_layers):
    return [m for m in fc_layers if isinstance(m, nn.Linear)]


class fcEncoderNet(nn.Module):
    """x -> (mu, sigma=softplus(.))  (reference nets/fc.py:19-61)"""

    def __init__(self, in_dim: Tuple[int], latent_dim: int = 2, c_dim: int = 0,
                 hidden_dim: List[int] = None, activation: str = 'tanh',
                 softplus_out: bool = True, flat: bool = True) -> None:
        super().__init__()
        _check_dim(in_dim)
        self.in_dim = _prod(in_dim) + c_dim
        if hidden_dim is None:
            hidden_dim = [128, 128]
        self.flat = flat
        self.activation = activation
        self.softplus_out = softplus_out
        self.concat = Concat()
        self.fc_layers = make_fc_layers(self.in_dim, hidden_dim, activation)
        self.fc11 = nn.Linear(hidden_dim[-1], latent_dim)
        self.fc12 = nn.Linear(hidden_dim[-1], latent_dim)

    def forward(self, x):
        x = broadcast_concat(x)
        if self.flat:
            x = x.reshape(-1, self.in_dim)
        x = x.contiguous().float()
        h = fc_stack_forward(self.fc_layers, x, self.activation)
        mu = ops.linear_fwd(h, self.fc11.weight.data, self.fc11.bias.data, None)
        sigma = ops.linear_fwd(h, self.fc12.weight.data, self.fc12.bias.data,
                               "softplus" if self.softplus_out else None)
        return mu, sigma


class jfcEncoderNet(nn.Module):
    """x -> (mu, sigma, class probabilities)  (reference nets/fc.py:64-108)"""

    def __init__(self, in_dim: Tuple[int], latent_dim: int = 2, discrete_dim: int = 0,
                 hidden_dim: List[int] = None, activation: str = 'tanh',
                 softplus_out: bool = True, flat: bool = True) -> None:
        super().__init__()
        _check_dim(in_dim)
        self.in_dim = _prod(in_dim)
        if hidden_dim is None:
            hidden_dim = [128, 128]
        self.flat = flat
        self.activation = activation
        self.softplus_out = softplus_out
        self.concat = Concat()
        self.fc_layers = make_fc_layers(self.in_dim, hidden_dim, activation)
        self.fc11 = nn.Linear(hidden_dim[-1], latent_dim)
        self.fc12 = nn.Linear(hidden_dim[-1], latent_dim)
        self.fc13 = nn.Linear(hidden_dim[-1], discrete_dim)

    def forward(self, x):
        x = broadcast_concat(x)
        if self.flat:
            x = x.reshape(-1, self.in_dim)
        x = x.contiguous().float()
        h = fc_stack_forward(self.fc_layers, x, self.activation)
        mu = ops.linear_fwd(h, self.fc11.weight.data, self.fc11.bias.data, None)
        sigma = ops.linear_fwd(h, self.fc12.weight.data, self.fc12.bias.data,
                               "softplus" if self.softplus_out else None)
        logits = ops.linear_fwd(h, self.fc13.weight.data, self.fc13.bias.data, None)
        alpha = torch.empty_like(logits)
        ops.enum_head_fwd(logits, alpha, None)
        return mu, sigma, alpha


class fcDecoderNet(nn.Module):
    """z -> x  (reference nets/fc.py:111-152)"""

    def __init__(self, out_dim: Tuple[int], latent_dim: int, c_dim: int = 0,
                 hidden_dim: List[int] = None, activation: str = 'tanh',
                 sigmoid_out: bool = True, unflat: bool = True) -> None:
        super().__init__()
        _check_dim(out_dim)
        self.unflat = unflat
        if self.unflat:
            self.reshape = out_dim
        n_out = _prod(out_dim)
        if hidden_dim is None:
            hidden_dim = [128, 128]
        self.activation = activation
        self.sigmoid_out = sigmoid_out
        self.concat = Concat()
        self.fc_layers = make_fc_layers(latent_dim + c_dim, hidden_dim, activation)
        self.out = nn.Linear(hidden_dim[-1], n_out)

    def forward(self, z):
        z = broadcast_concat(z).contiguous().float()
        h = fc_stack_forward(self.fc_layers, z, self.activation)
        x = ops.linear_fwd(h, self.out.weight.data, self.out.bias.data,
                           "sigmoid" if self.sigmoid_out else None)
        if self.unflat:
            return x.view(-1, *self.reshape)
        return x


class coord_latent(nn.Module):
    """First ("spatial") decoder layer: tanh(fc_coord(x') + fc_latent(z))
    (reference nets/fc.py:202-237; always tanh, fc_latent has no bias)."""

    def __init__(self, latent_dim: int, out_dim: int, ndim: int = 2,
                 activation_out: bool = True) -> None:
        super().__init__()
        self.fc_coord = nn.Linear(ndim, out_dim)
        self.fc_latent = nn.Linear(latent_dim, out_dim, bias=False)
        self.activation = nn.Tanh() if activation_out else None


class sDecoderNet(nn.Module):
    """Spatial decoder: per-pixel MLP over transformed coordinates + latent
    code (reference nets/fc.py:155-199)."""

    def __init__(self, out_dim: Tuple[int], latent_dim: int, c_dim: int = 0,
                 hidden_dim: List[int] = None, activation: str = 'tanh',
                 sigmoid_out: bool = True, unflat: bool = True) -> None:
        super().__init__()
        _check_dim(out_dim)
        self.unflat = unflat
        if self.unflat:
            self.reshape = out_dim
        self.out_dim = tuple(out_dim)
        if hidden_dim is None:
            hidden_dim = [128, 128]
        coord_dim = 1 if len(out_dim) < 2 else 2
        self.activation = activation
        self.sigmoid_out = sigmoid_out
        self.concat = Concat()
        self.coord_latent = coord_latent(latent_dim + c_dim, hidden_dim[0], coord_dim)
        self.fc_layers = make_fc_layers(hidden_dim[0], hidden_dim, activation)
        self.out = nn.Linear(hidden_dim[-1], 1)

    def forward(self, x_coord: torch.Tensor, z) -> torch.Tensor:
        """Generic call with explicit coordinates [B, N, ndim] (user code /
        reference `_decode`).  The coordinate layer is two small GEMMs plus a
        broadcast-add-tanh; the model classes never take this route -- they
        fold the affine map into the first layer instead."""
        z = broadcast_concat(z).contiguous().float()
        b, n = x_coord.shape[:2]
        cl = self.coord_latent
        hx = ops.linear_fwd(x_coord.reshape(b * n, -1).contiguous().float(),
                            cl.fc_coord.weight.data, cl.fc_coord.bias.data, None)
        hz = ops.linear_fwd(z.reshape(-1, z.shape[-1]), cl.fc_latent.weight.data, None, None)
        # broadcast add + tanh through the fold kernel's generic form:
        # pre0 = hx + hz[b]  ->  Uv = (0, 0, hz) is not expressible with explicit
        # coordinates, so use the identity layer: tanh(I * (hx + hz))
        h = (hx.view(b, n, -1) + hz.unsqueeze(1)).reshape(b * n, -1)
        eye = torch.eye(h.shape[-1], device=h.device)
        h = ops.linear_fwd(h, eye, None, "tanh")
        h = fc_stack_forward(self.fc_layers, h, self.activation)
        x = ops.linear_fwd(h, self.out.weight.data, self.out.bias.data,
                           "sigmoid" if self.sigmoid_out else None)
        if self.unflat:
            return x.view(-1, *self.reshape)
        return x


class fcClassifierNet(nn.Module):
    """x -> class probabilities (reference nets/fc.py:240-271)"""

    def __init__(self, in_dim: Tuple[int], num_classes: int, hidden_dim: List[int] = None,
                 activation: str = 'tanh') -> None:
        super().__init__()
        _check_dim(in_dim)
        self.in_dim = _prod(in_dim)
        if hidden_dim is None:
            hidden_dim = [128, 128]
        self.activation = activation
        self.fc_layers = make_fc_layers(self.in_dim, hidden_dim, activation)
        self.out = nn.Linear(hidden_dim[-1], num_classes)

    def forward(self, x):
        x = x.reshape(-1, self.in_dim).contiguous().float()
        h = fc_stack_forward(self.fc_layers, x, self.activation)
        logits = ops.linear_fwd(h, self.out.weight.data, self.out.bias.data, None)
        alpha = torch.empty_like(logits)
        ops.enum_head_fwd(logits, alpha, None)
        return alpha


class fcRegressorNet(nn.Module):
    """x -> continuous label (reference nets/fc.py:274-304)"""

    def __init__(self, in_dim: Tuple[int], c_dim: int, hidden_dim: List[int] = None,
                 activation: str = 'tanh') -> None:
        super().__init__()
        _check_dim(in_dim)
        self.in_dim = _prod(in_dim)
        if hidden_dim is None:
            hidden_dim = [128, 128]
        self.activation = activation
        self.fc_layers = make_fc_layers(self.in_dim, hidden_dim, activation)
        self.out = nn.Linear(hidden_dim[-1], c_dim)

    def forward(self, x):
        x = x.reshape(-1, self.in_dim).contiguous().float()
        h = fc_stack_forward(self.fc_layers, x, self.activation)
        return ops.linear_fwd(h, self.out.weight.data, self.out.bias.data, None)
