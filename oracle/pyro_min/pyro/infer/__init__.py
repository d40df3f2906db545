"""SVI / Trace_ELBO / TraceEnum_ELBO / config_enumerate restatement
(oracle-only; see pyro/__init__.py for the semantics relied upon)."""
import functools

import torch

from .. import poutine
from ..poutine.enum_messenger import EnumMessenger


class ELBO:
    def __init__(self, num_particles=1, max_plate_nesting=float("inf"),
                 strict_enumeration_warning=True, **kw):
        self.num_particles = num_particles
        self.max_plate_nesting = max_plate_nesting

    def _traces(self, model, guide, args, kwargs):
        guide_trace = poutine.trace(guide).get_trace(*args, **kwargs)
        model_trace = poutine.trace(
            poutine.replay(model, trace=guide_trace)).get_trace(*args, **kwargs)
        return model_trace, guide_trace

    def differentiable_loss(self, model, guide, *args, **kwargs):
        raise NotImplementedError

    def loss_and_grads(self, model, guide, *args, **kwargs):
        loss = self.differentiable_loss(model, guide, *args, **kwargs)
        if torch.is_tensor(loss) and loss.requires_grad:
            loss.backward()
        return float(loss)

    def loss(self, model, guide, *args, **kwargs):
        with torch.no_grad():
            return float(self.differentiable_loss(model, guide, *args, **kwargs))


class Trace_ELBO(ELBO):
    """loss = -(sum_model scale*log p - sum_guide scale*log q); all guide
    sites of the reference models are reparameterised Normals, so the
    surrogate loss equals the loss."""

    def differentiable_loss(self, model, guide, *args, **kwargs):
        model_trace, guide_trace = self._traces(model, guide, args, kwargs)
        elbo = 0.0
        for site in model_trace.nodes.values():
            if site["type"] == "sample":
                elbo = elbo + (site["fn"].log_prob(site["value"]) * site["scale"]).sum()
        for site in guide_trace.nodes.values():
            if site["type"] == "sample":
                elbo = elbo - (site["fn"].log_prob(site["value"]) * site["scale"]).sum()
        if not torch.is_tensor(elbo):
            elbo = torch.tensor(float(elbo))
        return -elbo


class TraceEnum_ELBO(ELBO):
    """Guide-side parallel enumeration.  Every cost term (scaled log p of a
    model site, minus scaled log q of a guide site) is weighted by the product
    of the UNSCALED probabilities of the enumerated guide sites and summed over
    the enumeration dim (exact expectation); costs that do not carry the
    enumeration dim broadcast (weights sum to one)."""

    def _traces(self, model, guide, args, kwargs):
        first = -1 - int(self.max_plate_nesting)
        guide_enum = EnumMessenger(first_available_dim=first)
        guide_trace = poutine.trace(guide_enum(guide)).get_trace(*args, **kwargs)
        model_trace = poutine.trace(
            poutine.replay(model, trace=guide_trace)).get_trace(*args, **kwargs)
        return model_trace, guide_trace

    def differentiable_loss(self, model, guide, *args, **kwargs):
        model_trace, guide_trace = self._traces(model, guide, args, kwargs)
        log_w = None
        for site in guide_trace.nodes.values():
            if site["type"] == "sample" and "_enumerate_dim" in site["infer"]:
                lq = site["fn"].log_prob(site["value"])
                log_w = lq if log_w is None else log_w + lq
        costs = []
        for site in model_trace.nodes.values():
            if site["type"] == "sample":
                costs.append(site["fn"].log_prob(site["value"]) * site["scale"])
        for site in guide_trace.nodes.values():
            if site["type"] == "sample":
                costs.append(-site["fn"].log_prob(site["value"]) * site["scale"])
        if not costs:
            return torch.tensor(0.0)
        if log_w is None:
            elbo = sum(c.sum() for c in costs)
        else:
            w = log_w.exp()
            elbo = 0.0
            for c in costs:
                elbo = elbo + (w * c).sum()
        return -elbo


def config_enumerate(guide=None, default="parallel", expand=False,
                     num_samples=None, tmc="diagonal"):
    """Mark every discrete (enumerable) unobserved site of `guide`."""
    if guide is None:
        return functools.partial(config_enumerate, default=default, expand=expand)

    class _Cfg(poutine.Messenger):
        def _process_message(self, msg):
            if msg["type"] != "sample" or msg["is_observed"]:
                return
            if getattr(msg["fn"], "has_enumerate_support", False):
                msg["infer"].setdefault("enumerate", default)
                msg["infer"].setdefault("expand", expand)

    @functools.wraps(guide)
    def wrapped(*args, **kwargs):
        with _Cfg():
            return guide(*args, **kwargs)
    return wrapped


class SVI:
    def __init__(self, model, guide, optim, loss, **kw):
        self.model, self.guide, self.optim, self.loss = model, guide, optim, loss

    def step(self, *args, **kwargs):
        with poutine.trace(param_only=True) as param_capture:
            loss = self.loss.loss_and_grads(self.model, self.guide, *args, **kwargs)
        params = []
        seen = set()
        for site in param_capture.trace.nodes.values():
            p = site["value"]
            if id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        self.optim(params)
        for p in params:  # pyro.infer.util.zero_grads
            if p.grad is not None:
                p.grad = torch.zeros_like(p.grad)
        return float(loss)

    def evaluate_loss(self, *args, **kwargs):
        return self.loss.loss(self.model, self.guide, *args, **kwargs)
