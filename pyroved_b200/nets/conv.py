"""Convolutional encoder / decoder parameter containers for VED.

Constructor signatures, attribute names and state_dict keys follow reference
nets/conv.py:24-277 (`encoder_z.feature_extractor.layers.{0,3,5,8,10}.*`,
`decoder.upsampler.layers.{4,9,12}.conv.*`, ...), so `.pt` checkpoints interchange and the same
seed gives the same initial weights.  The modules only OWN parameters and describe the layer
sequence; `forward` (inference) and training (conv_engine.VEDProgram) run the hand-written CUDA
kernels of csrc/pvb_conv.cu / pvb_conv3d.cu / pvb_norm.cu through the C ABI (1-D, 2-D and 3-D).
"""
import os
from typing import List, Tuple
from warnings import warn

import torch
import torch.nn as nn

from .. import ops
from ..utils.nn import get_activation, get_bnorm, get_conv, get_maxpool

DEFAULT_ENC = [(32,), (64, 64), (128, 128)]
DEFAULT_DEC = [(128, 128), (64, 64), (32,)]


def _prod(t):
    n = 1
    for v in t:
        n *= int(v)
    return n


def _check_ndim(ndim):
    if not 0 < ndim < 4:
        raise AssertionError("ndim must be equal to 1, 2 or 3")


class UpsampleBlock(nn.Module):
    """x2 interpolation ('bilinear' in 2-D, 'nearest' otherwise) followed by a 1x1 convolution
    (reference nets/conv.py:105-143)."""

    def __init__(self, ndim: int, input_channels: int, output_channels: int,
                 scale_factor: int = 2, mode: str = "bilinear") -> None:
        super().__init__()
        if mode not in ("bilinear", "nearest"):
            raise NotImplementedError("Use 'bilinear' or 'nearest' for upsampling mode")
        _check_ndim(ndim)
        if mode == "bilinear" and ndim in (3, 1):
            warn("'bilinear' mode is not supported for 1D and 3D; switching to 'nearest' mode",
                 category=UserWarning)
            mode = "nearest"
        if scale_factor != 2:
            raise NotImplementedError("pyroved_b200: only scale_factor=2 is implemented")
        self.mode = mode
        self.scale_factor = scale_factor
        self.conv = get_conv(ndim)(input_channels, output_channels, kernel_size=1, stride=1,
                                   padding=0)

    def forward(self, x):
        return run_layers([self], x, None)


def _conv_blocks(ndim, input_channels, conv_filters, activation, batchnorm, closer):
    """conv (+activation) (+batch norm, reference nets/conv.py:186-187) per filter count,
    `closer(block_index, convs_so_far, channels)` appended after every block (pool / upsample);
    the modules land at the reference's indices."""
    mods = []
    ch_in = input_channels
    n_convs = 0
    for i, block in enumerate(conv_filters):
        for ch in block:
            mods.append(get_conv(ndim)(ch_in, ch, 3, 1, 1))
            if activation is not None:
                mods.append(get_activation(activation)())
            if batchnorm:
                mods.append(get_bnorm(ndim)(ch))
            ch_in = ch
            n_convs += 1
        mods.extend(closer(i, n_convs, ch_in))
    return mods, ch_in


class FeatureExtractor(nn.Sequential):
    """3x3 conv blocks separated by 2x max-pools (reference nets/conv.py:146-196: a pool follows a
    block while more convolutions remain, plus one after the last block if `pool_last`)."""

    def __init__(self, ndim: int, input_channels: int = 1, conv_filters: List[int] = None,
                 kernel_size=3, stride=1, padding=1, batchnorm: bool = False,
                 activation: str = "lrelu", pool_last: bool = True) -> None:
        super().__init__()
        _check_ndim(ndim)
        if (kernel_size, stride, padding) != (3, 1, 1):
            raise NotImplementedError("pyroved_b200: kernel_size=3, stride=1, padding=1 only")
        if conv_filters is None:
            conv_filters = DEFAULT_ENC
        total = sum(len(c) for c in conv_filters)

        def closer(i, n_convs, ch):
            if n_convs + 1 < total or (n_convs + 1 >= total and pool_last):
                return [get_maxpool(ndim)(2, 2)]
            return []

        mods, _ = _conv_blocks(ndim, input_channels, conv_filters, activation, batchnorm, closer)
        self.activation = activation
        self.layers = nn.Sequential(*mods)

    def forward(self, x):
        return run_layers(self.layers, x, self.activation)


class Upsampler(nn.Sequential):
    """3x3 conv blocks, each followed by an UpsampleBlock, and a final 1x1 convolution to the
    output channels (reference nets/conv.py:199-249)."""

    def __init__(self, ndim: int, input_channels: int = 128, conv_filters: List[int] = None,
                 output_channels: int = 1, kernel_size=3, stride=1, padding=1,
                 batchnorm: bool = False, activation: str = "lrelu",
                 upsampling_mode: str = "bilinear") -> None:
        super().__init__()
        _check_ndim(ndim)
        if (kernel_size, stride, padding) != (3, 1, 1):
            raise NotImplementedError("pyroved_b200: kernel_size=3, stride=1, padding=1 only")
        if conv_filters is None:
            conv_filters = DEFAULT_DEC

        def closer(i, n_convs, ch):
            return [UpsampleBlock(ndim, ch, ch, mode=upsampling_mode)]

        mods, ch = _conv_blocks(ndim, input_channels, conv_filters, activation, batchnorm, closer)
        mods.append(get_conv(ndim)(ch, output_channels, 1, 1, 0))
        self.activation = activation
        self.layers = nn.Sequential(*mods)

    def forward(self, x):
        return run_layers(self.layers, x, self.activation)


class features_to_latent(nn.Module):
    """flatten -> Linear (reference nets/conv.py:252-263)"""

    def __init__(self, input_dim: Tuple[int], latent_dim: int = 2) -> None:
        super().__init__()
        self.reshape_ = _prod(input_dim)
        self.fc_latent = nn.Linear(self.reshape_, latent_dim)

    def forward(self, x):
        x = x.reshape(-1, self.reshape_).contiguous()
        return ops.linear_fwd(x, self.fc_latent.weight.data, self.fc_latent.bias.data, None)


class latent_to_features(nn.Module):
    """Linear -> [C, *dims] feature map (reference nets/conv.py:266-277)"""

    def __init__(self, latent_dim: int, out_dim: Tuple[int]) -> None:
        super().__init__()
        self.reshape_ = [int(d) for d in out_dim]
        self.fc = nn.Linear(latent_dim, _prod(out_dim))

    def forward(self, x):
        y = ops.linear_fwd(x.contiguous().float(), self.fc.weight.data, self.fc.bias.data, None)
        return y.view(-1, *self.reshape_)


class convEncoderNet(nn.Module):
    """x -> (mu, sigma)  (reference nets/conv.py:24-64)"""

    def __init__(self, input_dim: Tuple[int], latent_dim: int = 2, input_channels: int = 1,
                 hidden_dim: List[int] = None, batchnorm: bool = False, activation: str = "lrelu",
                 softplus_out: bool = True, pool_last: bool = False) -> None:
        super().__init__()
        if hidden_dim is None:
            hidden_dim = DEFAULT_ENC
        n_pool = len(hidden_dim) if pool_last else len(hidden_dim) - 1
        feat_dim = [int(d) // 2 ** n_pool for d in input_dim]
        self.input_dim = tuple(int(d) for d in input_dim)
        self.input_channels = input_channels
        self.latent_dim = latent_dim
        self.softplus_out = softplus_out
        self.feature_extractor = FeatureExtractor(len(input_dim), input_channels, hidden_dim,
                                                  batchnorm=batchnorm, activation=activation,
                                                  pool_last=pool_last)
        self.features2latent = features_to_latent([hidden_dim[-1][-1], *feat_dim], 2 * latent_dim)

    def forward(self, x):
        x = x.reshape(-1, self.input_channels, *self.input_dim).contiguous().float()
        enc = self.features2latent(self.feature_extractor(x))
        mu, s = enc[:, :self.latent_dim], enc[:, self.latent_dim:]
        if self.softplus_out:
            s = torch.nn.functional.softplus(s)   # [B, L] epilogue of an inference call
        return mu.contiguous(), s.contiguous()


class convDecoderNet(nn.Module):
    """z -> reconstruction [B, C, *output_dim]  (reference nets/conv.py:67-102)"""

    def __init__(self, latent_dim: int, output_dim: int, output_channels: int = 1,
                 hidden_dim: List[int] = None, batchnorm: bool = False, activation: str = "lrelu",
                 sigmoid_out: bool = True, upsampling_mode: str = "bilinear") -> None:
        super().__init__()
        if hidden_dim is None:
            hidden_dim = DEFAULT_DEC
        in_dim = [int(d) // 2 ** len(hidden_dim) for d in output_dim]
        self.output_dim = tuple(int(d) for d in output_dim)
        self.sigmoid_out = sigmoid_out
        self.latent2features = latent_to_features(latent_dim, [hidden_dim[0][0], *in_dim])
        self.upsampler = Upsampler(len(output_dim), hidden_dim[0][0], hidden_dim, output_channels,
                                   batchnorm=batchnorm, activation=activation,
                                   upsampling_mode=upsampling_mode)

    def forward(self, x):
        y = self.upsampler(self.latent2features(x))
        return torch.sigmoid(y) if self.sigmoid_out else y


# ---- layer-sequence description shared by inference (here) and training (conv_engine) --------
_CONVS = (nn.Conv1d, nn.Conv2d, nn.Conv3d)
_POOLS = (nn.MaxPool1d, nn.MaxPool2d, nn.MaxPool3d)
_BNORMS = (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)


def layer_plan(layers, activation):
    """[(kind, module, fused_activation)] with kind in 'conv' | 'bn' | 'pool' | 'up'.  An
    activation module directly after a convolution is fused into that convolution's epilogue."""
    mods = list(layers)
    plan, i = [], 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, _CONVS):
            k = m.kernel_size[0]
            if (any(s != 1 for s in m.stride) or any(p != k // 2 for p in m.padding)
                    or k not in (1, 3) or any(kk != k for kk in m.kernel_size)):
                raise NotImplementedError("pyroved_b200: conv layers must be k=1|3, stride 1, same padding")
            act = None
            if i + 1 < len(mods) and not isinstance(
                    mods[i + 1], _CONVS + _POOLS + _BNORMS + (UpsampleBlock,)):
                act = activation
                i += 1
            plan.append(("conv", m, act))
        elif isinstance(m, _BNORMS):
            plan.append(("bn", m, None))
        elif isinstance(m, _POOLS):
            plan.append(("pool", m, None))
        elif isinstance(m, UpsampleBlock):
            plan.append(("up", m, None))
            plan.append(("conv", m.conv, None))
        else:
            raise NotImplementedError("pyroved_b200: unsupported layer {}".format(type(m).__name__))
        i += 1
    return plan


def out_shape(kind, mod, shape):
    """shape = (C, *spatial) -> shape after the step"""
    c, sp = shape[0], list(shape[1:])
    if kind == "conv":
        return (mod.out_channels, *sp)
    if kind == "bn":
        return tuple(shape)
    if kind == "pool":
        return (c, *[s // 2 for s in sp])
    return (c, *[2 * s for s in sp])


def run_layers(layers, x, activation):
    """Forward-only execution of a layer sequence (inference; allocates its outputs)."""
    x = x.contiguous().float()
    for kind, mod, act in layer_plan(layers, activation):
        shp = out_shape(kind, mod, tuple(x.shape[1:]))
        y = torch.empty((x.shape[0], *shp), device=x.device, dtype=torch.float32)
        if kind == "conv":
            bias = mod.bias.data if mod.bias is not None else None
            if os.environ.get("PVB_FORCE_GENERIC", "0") != "1" and ops.conv_tc_supported(mod.weight):
                ops.conv_tc_fwd(x, mod.weight.data, bias, act, y, ops.conv_tc_workspace(mod.weight))
            else:
                ops.conv_fwd(x, mod.weight.data, bias, act, y)
        elif kind == "bn":
            # the module's own mode: running statistics in eval mode (VED.encode / decode /
            # manifold2d switch to it like the reference), batch statistics + a running-statistics
            # update otherwise (e.g. VED.predict before any eval() call, as in the reference)
            C = x.shape[1]
            stats = torch.empty(2, C, device=x.device, dtype=torch.float32)
            ops.bn_fwd(x, mod, y, stats[0], stats[1], ops.bn_workspace(C, x.device))
        elif kind == "pool":
            ops.maxpool2_fwd(x, y)
        else:
            ops.upsample2_fwd(x, y, mod.mode == "bilinear")
        x = y
    return x
