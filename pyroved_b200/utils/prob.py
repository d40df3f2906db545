"""Decoder samplers (reference utils/prob.py:5-37)."""
import torch.distributions as td

SAMPLERS = ("bernoulli", "continuous_bernoulli", "gaussian")


class Sampler:
    """Callable `loc -> distribution` like the reference's lambda, carrying the
    name and sigma the CUDA log-likelihood kernel needs."""

    def __init__(self, name, decoder_sig=0.5):
        self.name = name
        self.decoder_sig = float(decoder_sig)

    def __call__(self, loc):
        if self.name == "bernoulli":
            return td.Bernoulli(probs=loc, validate_args=False)
        if self.name == "continuous_bernoulli":
            return td.ContinuousBernoulli(probs=loc)
        return td.Normal(loc, self.decoder_sig)


def get_sampler(sampler: str, **kwargs) -> Sampler:
    if sampler not in SAMPLERS:
        raise KeyError(
            "Select between the following decoder "
            "samplers: {}".format(list(SAMPLERS)))
    return Sampler(sampler, kwargs.get("decoder_sig", 0.5))
