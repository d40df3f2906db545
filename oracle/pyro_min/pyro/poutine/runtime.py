"""Effect-handler stack (oracle-only; see pyro/__init__.py)."""
from collections import OrderedDict

_STACK = []
_PARAM_STORE = OrderedDict()
# Optional noise injection for parity runs: callable(name, fn) -> eps tensor
# (or None).  When set, a reparameterised Normal site takes the value
# loc + scale * eps instead of drawing from torch's global generator, so the
# reference and the CUDA path can be fed the SAME epsilon ("RNG
# parity: parity tests must pass eps in").
EPS_HOOK = [None]


def apply_stack(msg):
    # innermost handler first
    for h in reversed(_STACK):
        h._process_message(msg)
        if msg.get("stop"):
            break
    if msg["type"] == "sample" and not msg["done"]:
        fn = msg["fn"]
        eps = EPS_HOOK[0](msg["name"], fn) if EPS_HOOK[0] is not None else None
        if eps is not None:
            base = getattr(fn, "base_dist", fn)
            msg["value"] = base.loc + base.scale * eps
        elif getattr(fn, "has_rsample", False):
            msg["value"] = fn.rsample()
        else:
            msg["value"] = fn.sample()
        msg["done"] = True
    for h in _STACK:
        h._postprocess_message(msg)
    return msg
