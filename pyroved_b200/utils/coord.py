"""Coordinate grids and latent-space grids (host-side helpers).

Mirrors the public helpers of reference utils/coord.py.  The training hot path
never materialises a transformed grid: the per-sample affine map is folded into
the first decoder layer (csrc/pvb_latent.cu: fold_fwd) and grid coordinates are
regenerated from the pixel index inside the kernels.  These functions exist for
API compatibility (model.grid attribute, manifold2d grids, user code).
"""
from typing import Tuple, Union

import torch


def generate_grid(data_dim: Tuple[int]) -> torch.Tensor:
    """[N, 1] (1-D) or [N, 2] (2-D) coordinates in [-1, 1]
    (reference utils/coord.py:21-44: 1-D linspace(1,-1,L); 2-D rows
    p = i*W + j -> (linspace(-1,1,H)[i], linspace(1,-1,W)[j]))."""
    if len(data_dim) not in (1, 2):
        raise NotImplementedError("Currently supports only 1D and 2D data")
    if len(data_dim) == 1:
        return torch.linspace(1, -1, data_dim[0]).unsqueeze(1)
    rows = torch.linspace(-1, 1, data_dim[0])
    cols = torch.linspace(1, -1, data_dim[1])
    gx = rows.unsqueeze(1).expand(data_dim[0], data_dim[1]).reshape(-1)
    gy = cols.unsqueeze(0).expand(data_dim[0], data_dim[1]).reshape(-1)
    return torch.stack([gx, gy], dim=1)


def transform_coordinates(coord: torch.Tensor,
                          phi: Union[torch.Tensor, float] = 0,
                          coord_dx: Union[torch.Tensor, float] = 0,
                          scale: Union[torch.Tensor, float] = 1.) -> torch.Tensor:
    """Rotate -> scale -> translate a batch of grids [B, N, 2]; 1-D grids are
    only translated (reference utils/coord.py:47-88).  Utility for user code;
    not used by the CUDA training path."""
    if coord.shape[-1] == 1:
        return coord + coord_dx
    b = coord.shape[0]
    phi = torch.as_tensor(phi, dtype=coord.dtype, device=coord.device).expand(b)
    scale = torch.as_tensor(scale, dtype=coord.dtype, device=coord.device).expand(b)
    c, s = torch.cos(phi)[:, None], torch.sin(phi)[:, None]
    gx, gy = coord[..., 0], coord[..., 1]
    out = torch.stack([gx * c - gy * s, gx * s + gy * c], dim=-1) * scale[:, None, None]
    return out + coord_dx


def generate_latent_grid(d, **kwargs):
    """d x d grid of 2-D latent coordinates (reference utils/coord.py:91-109):
    either linear between `z_coord` bounds or N(0,1) quantiles 0.95..0.05."""
    if isinstance(d, int):
        d = [d, d]
    z_coord = kwargs.get("z_coord")
    if z_coord:
        z1, z2, z3, z4 = z_coord
        grid_x = torch.linspace(z2, z1, d[0])
        grid_y = torch.linspace(z3, z4, d[1])
    else:
        nrm = torch.distributions.Normal(0., 1.)
        grid_x = nrm.icdf(torch.linspace(0.95, 0.05, d[0]))
        grid_y = nrm.icdf(torch.linspace(0.05, 0.95, d[1]))
    z = torch.stack([grid_x.repeat_interleave(d[1]), grid_y.repeat(d[0])], dim=1).float()
    return z, (grid_x, grid_y)


def generate_latent_grid_traversal(d, cont_dim, disc_dim, cont_idx, cont_idx_fixed, num_samples):
    """Continuous x discrete traversal grids (reference utils/coord.py:112-133)."""
    samples_cont = torch.zeros(num_samples, cont_dim) + cont_idx_fixed
    trav = torch.distributions.Normal(0., 1.).icdf(torch.linspace(0.95, 0.05, d))
    samples_cont[:d * d, cont_idx] = trav.repeat(d)
    cls = torch.arange(disc_dim).tile(d // disc_dim + 1)[:d]
    samples_disc = torch.zeros(d * d, disc_dim)
    samples_disc[torch.arange(d * d), cls.repeat_interleave(d)] = 1
    return samples_cont, samples_disc
