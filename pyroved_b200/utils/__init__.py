"""Utility functions (names follow reference pyroved/utils/__init__.py)."""
from .coord import (generate_grid, generate_latent_grid,
                    generate_latent_grid_traversal, transform_coordinates)
from .data import (DeviceBatchLoader, TensorBatchLoader, init_dataloader,
                   init_ssvae_dataloaders)
from .nn import (get_activation, get_bnorm, get_conv, get_maxpool,
                 set_deterministic_mode, to_onehot, average_weights,
                 Concat, _to_device, broadcast_concat)
from .prob import get_sampler
from .viz import plot_grid_traversal, plot_img_grid, plot_spect_grid

__all__ = ['generate_grid', 'transform_coordinates', 'generate_latent_grid',
           'get_sampler', 'init_dataloader', 'init_ssvae_dataloaders',
           'get_activation', 'get_bnorm', 'get_conv', 'get_maxpool',
           'to_onehot', 'set_deterministic_mode', 'plot_img_grid',
           'plot_spect_grid', 'plot_grid_traversal',
           'generate_latent_grid_traversal', 'average_weights', 'TensorBatchLoader', 'DeviceBatchLoader']
