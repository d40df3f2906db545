"""iVAE: VAE with rotational / translational / scale invariances
(reference models/ivae.py:21-310).  Same constructor, attributes, state_dict
keys and encode/decode/manifold2d API; the SVI step itself runs as fused CUDA
kernels (engine.SpatialVAEProgram)."""
from typing import List, Tuple, Union

import torch

from .base import baseVAE
from ..nets import fcDecoderNet, fcEncoderNet, sDecoderNet
from ..utils import (generate_latent_grid, get_sampler, plot_img_grid, plot_spect_grid,
                     set_deterministic_mode)


class iVAE(baseVAE):
    """
    Args:
        data_dim: (h, w) or (length,)
        latent_dim: number of "content" latent dimensions
        invariances: subset of ['r', 't', 's'] (1-D: ['t']) or None (vanilla VAE)
        c_dim: size of the conditioning vector (0 = unconditional)
        hidden_dim_e / hidden_dim_d: hidden sizes (default [128, 128])
        activation: 'tanh' (default), 'relu', 'lrelu', 'softplus', 'gelu'
        sampler_d: 'bernoulli' (default), 'gaussian', 'continuous_bernoulli'
        sigmoid_d: sigmoid on the decoder output (default True)
        seed: torch seed used for weight init (default 1)
    Keyword Args: device, dx_prior, dy_prior, sc_prior, decoder_sig
    """

    def __init__(self, data_dim: Tuple[int], latent_dim: int = 2, invariances: List[str] = None,
                 c_dim: int = 0, hidden_dim_e: List[int] = None, hidden_dim_d: List[int] = None,
                 activation: str = "tanh", sampler_d: str = "bernoulli", sigmoid_d: bool = True,
                 seed: int = 1, **kwargs: Union[str, float]) -> None:
        super().__init__(data_dim, invariances, **kwargs)
        set_deterministic_mode(seed)
        self.data_dim = tuple(data_dim)
        self.encoder_z = fcEncoderNet(data_dim, latent_dim + self.coord, c_dim, hidden_dim_e,
                                      activation, softplus_out=True)
        dnet = sDecoderNet if 0 < self.coord < 5 else fcDecoderNet
        self.decoder = dnet(data_dim, latent_dim, c_dim, hidden_dim_d, activation,
                            sigmoid_out=sigmoid_d)
        self.sampler_d = get_sampler(sampler_d, **kwargs)
        self.z_dim = latent_dim + self.coord
        self.c_dim = c_dim
        self._latent_dim = latent_dim
        self.to(self.device)

    def split_latent(self, z: torch.Tensor):
        return self._split_latent(z)

    def encode(self, x_new: torch.Tensor, y: torch.Tensor = None, **kwargs):
        """(z_mean, z_sd) with the transform latents first (rotation, dx, dy,
        scale) followed by the `latent_dim` content latents."""
        enc_args = [x_new, y] if y is not None else [x_new]
        z = self._encode(*enc_args, **kwargs)
        z_loc, z_scale = z.split(self.z_dim, 1)
        return z_loc, z_scale

    def decode(self, z: torch.Tensor, y: torch.Tensor = None, **kwargs) -> torch.Tensor:
        z = z.to(self.device)
        if y is not None:
            z = torch.cat([z, y.to(self.device)], -1)
        return self._decode(z, **kwargs)

    def manifold2d(self, d: int, y: torch.Tensor = None, plot: bool = True, **kwargs):
        z, (grid_x, grid_y) = generate_latent_grid(d, **kwargs)
        z = [z]
        if self.c_dim > 0:
            if y is None:
                raise ValueError("To generate a manifold pass a conditional vector y")
            y = y.unsqueeze(1) if 0 < y.ndim < 2 else y
            z = z + [y.expand(z[0].shape[0], *y.shape[1:])]
        loc = self.decode(*z, **kwargs)
        if plot:
            if self.ndim == 2:
                plot_img_grid(loc, d, extent=[grid_x.min(), grid_x.max(), grid_y.min(),
                                              grid_y.max()], **kwargs)
            elif self.ndim == 1:
                plot_spect_grid(loc, d, **kwargs)
        return loc

    def predict_on_latent(self, *args, **kwargs):
        """Reference models/ivae.py:312-349 fits a Pyro Gaussian process (`pyro.contrib.gp`,
        utils/gp.py) on the encoded data: outside the SVI hot path this package implements
        (DESIGN.md 8).  Use `encode()` / `decode()` / `manifold2d()` and any GP library on the
        codes."""
        raise NotImplementedError(
            "pyroved_b200 implements the SVI training / inference path, not the Pyro GP regression "
            "of predict_on_latent; encode() the data and fit a GP on the latent codes instead")

    def _make_program(self, engine, B, has_y, mode="main"):
        from ..engine import SpatialVAEProgram
        return SpatialVAEProgram(engine, B, has_y)
