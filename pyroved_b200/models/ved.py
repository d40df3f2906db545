"""VED: variational encoder-decoder whose input and output differ (im2spec / spec2im)
(reference models/ved.py:22-243).  Same constructor, attributes, state_dict keys and
encode / decode / predict / manifold2d API; the SVI step runs as hand-written CUDA kernels
(conv_engine.VEDProgram)."""
from typing import List, Tuple, Union

import torch

from .base import baseVAE
from ..nets.conv import convDecoderNet, convEncoderNet
from ..utils import (generate_latent_grid, get_sampler, init_dataloader, plot_img_grid,
                     plot_spect_grid, set_deterministic_mode)


class VED(baseVAE):
    """
    Args:
        input_dim: (h, w) for images or (length,) for spectra
        output_dim: (h, w) or (length,); need not match the input
        input_channels / output_channels: default 1
        latent_dim: number of latent dimensions
        hidden_dim_e: conv filters per encoder block, default [(32,), (64, 64), (128, 128)]
        hidden_dim_d: conv filters per decoder block, default [(128, 128), (64, 64), (32,)]
        activation: 'lrelu' (default), 'relu', 'tanh', 'softplus', 'gelu'
        batchnorm: BatchNorm after every conv + activation.  Training uses batch statistics;
            `encode`, `decode` and `manifold2d` switch the model to eval mode like the reference
            (models/ved.py:178,193,230), so they normalise with the running statistics -- and,
            as in the reference, the model STAYS in eval mode afterwards (call `.train()` before
            training further)
        sampler_d: 'bernoulli' (default) or 'gaussian'
        sigmoid_d: sigmoid on the decoder output (default True)
        seed: torch seed used for weight init
    Keyword Args: decoder_sig (sigma of the gaussian sampler), device
    """

    def __init__(self, input_dim: Tuple[int], output_dim: Tuple[int], input_channels: int = 1,
                 output_channels: int = 1, latent_dim: int = 2, hidden_dim_e: List[int] = None,
                 hidden_dim_d: List[int] = None, activation: str = "lrelu",
                 batchnorm: bool = False, sampler_d: str = "bernoulli", sigmoid_d: bool = True,
                 seed: int = 1, **kwargs: float) -> None:
        super().__init__(output_dim, None, **kwargs)
        set_deterministic_mode(seed)
        self.ndim = len(output_dim)
        self.input_dim = tuple(int(d) for d in input_dim)
        self.output_dim = tuple(int(d) for d in output_dim)
        self.input_channels, self.output_channels = input_channels, output_channels
        self.activation = activation
        self.encoder_z = convEncoderNet(input_dim, latent_dim, input_channels, hidden_dim_e,
                                        batchnorm, activation)
        self.decoder = convDecoderNet(latent_dim, output_dim, output_channels, hidden_dim_d,
                                      batchnorm, activation, sigmoid_d)
        self.sampler_d = get_sampler(sampler_d, **kwargs)
        self.z_dim = latent_dim
        self.to(self.device)

    def encode(self, x_new: torch.Tensor, **kwargs: int):
        """(z_mean, z_sd) of q(z|x), batch by batch (eval mode, reference ved.py:178)."""
        self.eval()
        z = self._encode(x_new, **kwargs)
        return z.split(self.z_dim, 1)

    def decode(self, z: torch.Tensor, **kwargs: int) -> torch.Tensor:
        """Decoded latent codes (eval mode, reference ved.py:193)."""
        self.eval()
        return self._decode(z.to(self.device), **kwargs)

    def predict(self, x_new: torch.Tensor, **kwargs: int):
        """encode -> 30 samples of z -> decode; mean and std of the decoded samples
        (reference models/ved.py:198-216)."""
        loader = init_dataloader(x_new, shuffle=False, **kwargs)
        mus, sds = [], []
        for (x_i,) in loader:
            z_mu, z_sig = self.encoder_z(x_i.to(self.device))
            eps = torch.randn((30,) + tuple(z_mu.shape), device=z_mu.device)
            zs = z_mu[None] + z_sig[None] * eps
            y = torch.stack([self.decoder(z) for z in zs])
            mus.append(y.mean(0).cpu())
            sds.append(y.std(0).cpu())
        return torch.cat(mus), torch.cat(sds)

    def manifold2d(self, d: int, plot: bool = True, **kwargs: Union[str, int]) -> torch.Tensor:
        self.eval()                     # reference ved.py:230
        z, (grid_x, grid_y) = generate_latent_grid(d, **kwargs)
        loc = self.decoder(z.to(self.device)).cpu()
        if plot:
            if self.ndim == 2:
                plot_img_grid(loc, d, extent=[grid_x.min(), grid_x.max(), grid_y.min(),
                                              grid_y.max()], **kwargs)
            elif self.ndim == 1:
                plot_spect_grid(loc, d, **kwargs)
        return loc

    def _make_program(self, engine, B, has_y, mode="main"):
        from ..conv_engine import VEDProgram
        if not has_y:
            raise ValueError("VED needs (x, y) pairs: the decoder output is scored against y")
        return VEDProgram(engine, B)
