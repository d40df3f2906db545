"""jiVAE: joint continuous + discrete latent, exact enumeration of the
discrete one (reference models/jivae.py:21-329; TraceEnum_ELBO semantics in
SURVEY.md 3.2).  Train with SVItrainer(model, enumerate_parallel=True)."""
from typing import List, Tuple, Union

import torch

from .base import baseVAE
from ..nets import fcDecoderNet, jfcEncoderNet, sDecoderNet
from ..utils import (generate_latent_grid, generate_latent_grid_traversal, get_sampler,
                     plot_grid_traversal, plot_img_grid, plot_spect_grid,
                     set_deterministic_mode, to_onehot)


class jiVAE(baseVAE):
    """
    Args:
        data_dim, latent_dim, discrete_dim (number of classes), invariances,
        hidden_dim_e, hidden_dim_d, activation, sampler_d, sigmoid_d, seed
    Keyword Args: device, dx_prior, dy_prior, sc_prior, decoder_sig
    `scale_factor` at step time may be a number or [continuous, discrete].
    """

    def __init__(self, data_dim: Tuple[int], latent_dim: int, discrete_dim: int,
                 invariances: List[str] = None, hidden_dim_e: List[int] = None,
                 hidden_dim_d: List[int] = None, activation: str = "tanh",
                 sampler_d: str = "bernoulli", sigmoid_d: bool = True, seed: int = 1,
                 **kwargs: Union[str, float]) -> None:
        super().__init__(data_dim, invariances, **kwargs)
        set_deterministic_mode(seed)
        self.data_dim = data_dim
        self.encoder_z = jfcEncoderNet(data_dim, latent_dim + self.coord, discrete_dim,
                                       hidden_dim_e, activation, softplus_out=True)
        dnet = sDecoderNet if 0 < self.coord < 5 else fcDecoderNet
        self.decoder = dnet(data_dim, latent_dim, discrete_dim, hidden_dim_d, activation,
                            sigmoid_out=sigmoid_d, unflat=False)
        self.sampler_d = get_sampler(sampler_d, **kwargs)
        self.z_dim = latent_dim + self.coord
        self.discrete_dim = discrete_dim
        self._latent_dim = latent_dim
        self.to(self.device)

    def split_latent(self, z: torch.Tensor):
        return self._split_latent(z)

    def encode(self, x_new: torch.Tensor, logits: bool = False, **kwargs):
        """(z_mean, z_sd, class) -- class index, or probabilities if logits=True."""
        z = self._encode(x_new, **kwargs)
        z_loc = z[:, :self.z_dim]
        z_scale = z[:, self.z_dim:2 * self.z_dim]
        classes = z[:, 2 * self.z_dim:]
        if not logits:
            _, classes = torch.max(classes, 1)
        return z_loc, z_scale, classes

    def decode(self, z: torch.Tensor, y: torch.Tensor, **kwargs) -> torch.Tensor:
        z = torch.cat([z.to(self.device), y.to(self.device)], -1)
        loc = self._decode(z, **kwargs)
        return loc.view(-1, *self.data_dim)

    def manifold2d(self, d: int, disc_idx: int = 0, plot: bool = True, **kwargs):
        z, (grid_x, grid_y) = generate_latent_grid(d, **kwargs)
        z_disc = to_onehot(torch.tensor(disc_idx).unsqueeze(0), self.discrete_dim)
        z_disc = z_disc.repeat(z.shape[0], 1)
        loc = self.decode(z, z_disc, **kwargs)
        if plot:
            if self.ndim == 2:
                plot_img_grid(loc, d, extent=[grid_x.min(), grid_x.max(), grid_y.min(),
                                              grid_y.max()], **kwargs)
            elif self.ndim == 1:
                plot_spect_grid(loc, d, **kwargs)
        return loc

    def manifold_traversal(self, d: int, cont_idx: int, cont_idx_fixed: int = 0,
                           plot: bool = True, **kwargs):
        samples_cont, samples_disc = generate_latent_grid_traversal(
            d, self.z_dim - self.coord, self.discrete_dim, cont_idx, cont_idx_fixed, d ** 2)
        decoded = self.decode(samples_cont, samples_disc, **kwargs)
        if plot:
            plot_grid_traversal(decoded, d, **kwargs)
        return decoded

    # ---- engine hooks ---------------------------------------------------------
    def _beta(self, kwargs):
        """scale_factor -> (continuous, discrete) as in jivae.py:161-165."""
        b = kwargs.get("scale_factor", [1., 1.])
        b = torch.as_tensor(b, dtype=torch.float32)
        if b.ndim == 0:
            return (float(b), float(b))
        return (float(b[0]), float(b[1]))

    def _make_program(self, engine, B, has_y, mode="main"):
        from ..engine import EnumVAEProgram
        if not engine.enumerate_parallel:
            raise ValueError("jiVAE has a discrete latent: use "
                             "SVItrainer(model, enumerate_parallel=True)")
        return EnumVAEProgram(engine, B, "jivae")
