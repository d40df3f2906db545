"""ss_reg_iVAE: semi-supervised invariant VAE for regression, with a regressor q(y|x) = N(c(x), s)
(reference models/ss_reg_ivae.py:24-346).  Train with trainers.auxSVItrainer(task="regression")."""
from typing import List, Optional, Tuple, Union

import torch

from .base import baseVAE
from ..nets import fcDecoderNet, fcEncoderNet, fcRegressorNet, sDecoderNet
from ..utils import (generate_latent_grid, get_sampler, init_dataloader, plot_img_grid,
                     plot_spect_grid, set_deterministic_mode)


class ss_reg_iVAE(baseVAE):
    """
    Args:
        data_dim, latent_dim, reg_dim (number of regression outputs), invariances,
        hidden_dim_e, hidden_dim_d, hidden_dim_reg, activation, sampler_d, sigmoid_d, seed
    Keyword Args: device, dx_prior, dy_prior, sc_prior, decoder_sig, regressor_sig (0.5)
    """

    def __init__(self, data_dim: Tuple[int], latent_dim: int, reg_dim: int,
                 invariances: List[str] = None, hidden_dim_e: List[int] = None,
                 hidden_dim_d: List[int] = None, hidden_dim_reg: List[int] = None,
                 activation: str = "tanh", sampler_d: str = "bernoulli", sigmoid_d: bool = True,
                 seed: int = 1, **kwargs: Union[str, float]) -> None:
        super().__init__(data_dim, invariances, **kwargs)
        set_deterministic_mode(seed)
        self.data_dim = data_dim
        self.encoder_z = fcEncoderNet(data_dim, latent_dim + self.coord, reg_dim, hidden_dim_e,
                                      activation, flat=False)
        self.encoder_y = fcRegressorNet(data_dim, reg_dim, hidden_dim_reg, activation)
        dnet = sDecoderNet if 0 < self.coord < 5 else fcDecoderNet
        self.decoder = dnet(data_dim, latent_dim, reg_dim, hidden_dim_d, activation,
                            sigmoid_out=sigmoid_d, unflat=False)
        self.sampler_d = get_sampler(sampler_d, **kwargs)
        self.reg_sig = kwargs.get("regressor_sig", 0.5)
        self.z_dim = latent_dim + self.coord
        self.reg_dim = reg_dim
        self._latent_dim = latent_dim
        self.to(self.device)

    def split_latent(self, zs: torch.Tensor):
        zdims = list(zs.shape)
        zdims[-1] = zdims[-1] - self.coord
        zs = zs.view(-1, zs.size(-1))
        phi, dx, sc, zs = self._split_latent(zs)
        return phi, dx, sc, zs.view(*zdims)

    def set_regressor(self, reg_net: torch.nn.Module) -> None:
        self.encoder_y = reg_net

    def regressor(self, x_new: torch.Tensor, **kwargs) -> torch.Tensor:
        """Regressor predictions, batch by batch (reference ss_reg_ivae.py:254-278)."""
        loader = init_dataloader(x_new, shuffle=False, **kwargs)
        return torch.cat([self.encoder_y(x_i.to(self.device)).cpu() for (x_i,) in loader])

    def encode(self, x_new: torch.Tensor, y: Optional[torch.Tensor] = None, **kwargs):
        if y is None:
            y = self.regressor(x_new, **kwargs)
        z = self._encode(x_new.flatten(1), y, **kwargs)
        z_loc, z_scale = z.split(self.z_dim, 1)
        return z_loc, z_scale, y

    def decode(self, z: torch.Tensor, y: torch.Tensor, **kwargs) -> torch.Tensor:
        z = torch.cat([z.to(self.device), y.to(self.device)], -1)
        return self._decode(z, **kwargs).view(-1, *self.data_dim)

    def manifold2d(self, d: int, y: torch.Tensor, plot: bool = True, **kwargs):
        z, (grid_x, grid_y) = generate_latent_grid(d, **kwargs)
        y = y.unsqueeze(1) if 0 < y.ndim < 2 else y
        y = y.expand(z.shape[0], *y.shape[1:])
        loc = self.decode(z, y, **kwargs)
        if plot:
            if self.ndim == 2:
                plot_img_grid(loc, d, extent=[grid_x.min(), grid_x.max(), grid_y.min(),
                                              grid_y.max()], **kwargs)
            elif self.ndim == 1:
                plot_spect_grid(loc, d, **kwargs)
        return loc

    # ---- engine hooks ---------------------------------------------------------
    def _aux_scale(self, kwargs):
        return float(kwargs.get("aux_loss_multiplier", 20))

    def _make_program(self, engine, B, has_y, mode="main"):
        from ..engine import RegressorAuxProgram, SsRegProgram
        if mode == "aux":
            return RegressorAuxProgram(engine, B, has_y)
        return SsRegProgram(engine, B, has_y)
