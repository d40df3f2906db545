"""Thin mixins over torch.distributions (Pyro's own are the same: its
Normal/Bernoulli/OneHotCategorical subclass the torch classes and add
to_event/has_enumerate_support).  Oracle-only."""
import torch
import torch.distributions as td
from torch.distributions import constraints  # noqa: F401

from . import util  # noqa: F401


class TorchDistributionMixin:
    def to_event(self, reinterpreted_batch_ndims=None):
        if reinterpreted_batch_ndims is None:
            reinterpreted_batch_ndims = len(self.batch_shape)
        if reinterpreted_batch_ndims == 0:
            return self
        return Independent(self, reinterpreted_batch_ndims)

    def __call__(self, sample_shape=torch.Size()):
        return self.rsample(sample_shape) if self.has_rsample \
            else self.sample(sample_shape)


class Distribution(TorchDistributionMixin):
    pass


class Independent(td.Independent, TorchDistributionMixin):
    pass


class Normal(td.Normal, TorchDistributionMixin):
    pass


class Bernoulli(td.Bernoulli, TorchDistributionMixin):
    pass


class ContinuousBernoulli(td.ContinuousBernoulli, TorchDistributionMixin):
    pass


class OneHotCategorical(td.OneHotCategorical, TorchDistributionMixin):
    pass


class Categorical(td.Categorical, TorchDistributionMixin):
    pass
