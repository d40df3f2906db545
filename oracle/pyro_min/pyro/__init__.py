"""pyro_min -- ORACLE-ONLY restatement of the slice of Pyro (pyro-ppl >= 1.6.0,
not vendored under /root/reference and not installable offline) that the
unmodified reference package touches.  TEST INFRASTRUCTURE: never imported by
the product package `pyroved_b200`.

Surface follows SURVEY.md Appendix B.  Semantics restated from Pyro's
documented behaviour:
  * guide is traced first, the model is replayed against the guide trace;
  * each sample site contributes  scale * fn.log_prob(value)  (to_event(1)
    sums the event dim, plates keep the batch dim);
  * ELBO = sum(model sites) - sum(guide sites); loss = -ELBO;
  * an enumerated guide site takes its whole support along a new leftmost
    dim (first_available_dim = -1 - max_plate_nesting) and downstream costs
    are weighted by the (unscaled) guide probabilities;
  * SVI.step = loss_and_grads -> per-parameter torch.optim.Adam ->
    zero grads -> python float.

Reference call sites: pyroved/trainers/svi.py:75-91,107;
trainers/auxsvi.py:67-81; models/ivae.py:173-221; models/jivae.py:159-220;
models/ssivae.py:160-242; models/ved.py:130-163; utils/prob.py:26-28;
utils/nn.py:8.
"""
from collections import OrderedDict

from . import poutine  # noqa: F401
from .poutine.runtime import _PARAM_STORE, apply_stack
from . import distributions  # noqa: F401
from . import infer  # noqa: F401
from . import optim  # noqa: F401

__version__ = "1.6.0-min"


def clear_param_store():
    _PARAM_STORE.clear()


def get_param_store():
    return _PARAM_STORE


def module(name, nn_module, update_module_params=False):
    """Register every parameter of `nn_module` under '<name>$$$<param>'."""
    for pname, p in nn_module.named_parameters():
        full = "{}$$${}".format(name, pname)
        _PARAM_STORE[full] = p
        apply_stack({"type": "param", "name": full, "value": p, "fn": None,
                     "is_observed": False, "scale": 1.0, "infer": {},
                     "done": True})
    return nn_module


def sample(name, fn, obs=None, infer=None, **kwargs):
    msg = {"type": "sample", "name": name, "fn": fn, "value": obs,
           "is_observed": obs is not None, "scale": 1.0,
           "infer": dict(infer) if infer else {}, "done": obs is not None,
           "cond_indep_stack": ()}
    apply_stack(msg)
    return msg["value"]


class plate(poutine.Messenger):
    """Independence context.  Only bookkeeping is needed for the reference
    models (no subsampling is used anywhere: models/*.py)."""

    def __init__(self, name, size=None, subsample_size=None, dim=None, **kw):
        super().__init__()
        self.name, self.size, self.dim = name, size, dim

    def _process_message(self, msg):
        if msg["type"] == "sample":
            msg["cond_indep_stack"] = (self.name,) + tuple(
                msg.get("cond_indep_stack", ()))
