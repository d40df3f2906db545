"""ssiVAE: semi-supervised invariant VAE with a classifier q(y|x)
(reference models/ssivae.py:21-384).  Train with trainers.auxSVItrainer."""
import math
from typing import List, Optional, Tuple, Union

import torch

from .base import baseVAE
from ..nets import fcClassifierNet, fcDecoderNet, fcEncoderNet, sDecoderNet
from ..utils import (generate_latent_grid, generate_latent_grid_traversal, get_sampler,
                     init_dataloader, plot_grid_traversal, plot_img_grid, plot_spect_grid,
                     set_deterministic_mode, to_onehot)


class ssiVAE(baseVAE):
    """
    Args:
        data_dim, latent_dim, num_classes, invariances, hidden_dim_e,
        hidden_dim_d, hidden_dim_cls, activation, sampler_d, sigmoid_d, seed
    Keyword Args: device, dx_prior, dy_prior, sc_prior, decoder_sig
    """

    def __init__(self, data_dim: Tuple[int], latent_dim: int, num_classes: int,
                 invariances: List[str] = None, hidden_dim_e: List[int] = None,
                 hidden_dim_d: List[int] = None, hidden_dim_cls: List[int] = None,
                 activation: str = "tanh", sampler_d: str = "bernoulli", sigmoid_d: bool = True,
                 seed: int = 1, **kwargs: Union[str, float]) -> None:
        super().__init__(data_dim, invariances, **kwargs)
        set_deterministic_mode(seed)
        self.data_dim = data_dim
        self.encoder_z = fcEncoderNet(data_dim, latent_dim + self.coord, num_classes,
                                      hidden_dim_e, activation, flat=False)
        self.encoder_y = fcClassifierNet(data_dim, num_classes, hidden_dim_cls, activation)
        dnet = sDecoderNet if 0 < self.coord < 5 else fcDecoderNet
        self.decoder = dnet(data_dim, latent_dim, num_classes, hidden_dim_d, activation,
                            sigmoid_out=sigmoid_d, unflat=False)
        self.sampler_d = get_sampler(sampler_d, **kwargs)
        self.z_dim = latent_dim + self.coord
        self.num_classes = num_classes
        self._latent_dim = latent_dim
        self.to(self.device)

    def split_latent(self, zs: torch.Tensor):
        """Flattening variant (reference ssivae.py:217-227)."""
        zdims = list(zs.shape)
        zdims[-1] = zdims[-1] - self.coord
        zs = zs.view(-1, zs.size(-1))
        phi, dx, sc, zs = self._split_latent(zs)
        return phi, dx, sc, zs.view(*zdims)

    def set_classifier(self, cls_net: torch.nn.Module) -> None:
        self.encoder_y = cls_net

    def classifier(self, x_new: torch.Tensor, **kwargs) -> torch.Tensor:
        """Predicted class indices, batch by batch (reference ssivae.py:256-280)."""
        loader = init_dataloader(x_new, shuffle=False, batch_size=kwargs.get("batch_size", 100))
        out = []
        for (x_i,) in loader:
            alpha = self.encoder_y(x_i.to(self.device))
            out.append(torch.max(alpha, 1)[1].cpu())
        return torch.cat(out)

    def encode(self, x_new: torch.Tensor, y: Optional[torch.Tensor] = None, **kwargs):
        if y is None:
            y = self.classifier(x_new, **kwargs)
        if y.ndim < 2:
            y = to_onehot(y, self.num_classes)
        z = self._encode(x_new.flatten(1), y, **kwargs)
        z_loc, z_scale = z.split(self.z_dim, 1)
        _, y_pred = torch.max(y, 1)
        return z_loc, z_scale, y_pred

    def decode(self, z: torch.Tensor, y: torch.Tensor, **kwargs) -> torch.Tensor:
        z = torch.cat([z.to(self.device), y.to(self.device)], -1)
        loc = self._decode(z, **kwargs)
        return loc.view(-1, *self.data_dim)

    def manifold2d(self, d: int, plot: bool = True, **kwargs):
        z, (grid_x, grid_y) = generate_latent_grid(d, **kwargs)
        cls = torch.as_tensor(kwargs.get("label", 0))
        if cls.ndim < 2:
            cls = to_onehot(cls.reshape(1), self.num_classes)
        cls = cls.repeat(z.shape[0], 1)
        loc = self.decode(z, cls, **kwargs)
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
            d, self.z_dim - self.coord, self.num_classes, cont_idx, cont_idx_fixed, d ** 2)
        decoded = self.decode(samples_cont, samples_disc, **kwargs)
        if plot:
            plot_grid_traversal(decoded, d, **kwargs)
        return decoded

    # ---- engine hooks ---------------------------------------------------------
    def _aux_scale(self, kwargs):
        return float(kwargs.get("aux_loss_multiplier", 20))

    def _make_program(self, engine, B, has_y, mode="main"):
        from ..engine import ClassifierAuxProgram, EnumVAEProgram, SpatialVAEProgram
        if mode == "aux":
            return ClassifierAuxProgram(engine, B, has_y)
        if has_y:   # supervised: y observed, constant log p(y) = log(1/K) per sample
            return SpatialVAEProgram(engine, B, True, cond_dim=self.num_classes,
                                     loss_const=B * math.log(self.num_classes))
        return EnumVAEProgram(engine, B, "ssivae")
