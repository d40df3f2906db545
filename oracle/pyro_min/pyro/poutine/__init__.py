"""poutine subset: Messenger, trace, replay, scale, enum (oracle-only)."""
import functools
from collections import OrderedDict

from .runtime import _STACK


class Messenger:
    def __init__(self, fn=None):
        self.fn = fn

    def __enter__(self):
        _STACK.append(self)
        return self

    def __exit__(self, *exc):
        assert _STACK[-1] is self
        _STACK.pop()
        return False

    def _process_message(self, msg):
        pass

    def _postprocess_message(self, msg):
        pass

    def __call__(self, fn):
        # decorator form: EnumMessenger(...)(guide)
        outer = self

        @functools.wraps(fn)
        def wrapped(*args, **kwargs):
            with outer:
                return fn(*args, **kwargs)
        return wrapped


class Trace:
    def __init__(self):
        self.nodes = OrderedDict()

    def add_node(self, name, site):
        self.nodes[name] = site

    def compute_log_prob(self):
        for name, site in self.nodes.items():
            if site["type"] == "sample" and "log_prob" not in site:
                lp = site["fn"].log_prob(site["value"])
                site["unscaled_log_prob"] = lp
                site["log_prob"] = lp * site["scale"]
                site["log_prob_sum"] = site["log_prob"].sum()

    def log_prob_sum(self):
        self.compute_log_prob()
        return sum(s["log_prob_sum"] for s in self.nodes.values()
                   if s["type"] == "sample")

    def stochastic_nodes(self):
        return [n for n, s in self.nodes.items()
                if s["type"] == "sample" and not s["is_observed"]]


class TraceMessenger(Messenger):
    def __init__(self, fn=None, graph_type="flat", param_only=False):
        super().__init__(fn)
        self.param_only = param_only
        self.trace = Trace()

    def __enter__(self):
        self.trace = Trace()
        return super().__enter__()

    def _postprocess_message(self, msg):
        if self.param_only and msg["type"] != "param":
            return
        self.trace.add_node(msg["name"], dict(msg))

    def get_trace(self, *args, **kwargs):
        with self:
            self.fn(*args, **kwargs)
        return self.trace


def trace(fn=None, graph_type="flat", param_only=False):
    return TraceMessenger(fn, graph_type=graph_type, param_only=param_only)


class ReplayMessenger(Messenger):
    def __init__(self, fn=None, trace=None):
        super().__init__(fn)
        self.guide_trace = trace

    def _process_message(self, msg):
        if msg["type"] != "sample" or msg["is_observed"]:
            return
        site = self.guide_trace.nodes.get(msg["name"])
        if site is not None and site["type"] == "sample":
            msg["value"] = site["value"]
            msg["infer"] = dict(site.get("infer", {}))
            msg["done"] = True


def replay(fn=None, trace=None):
    m = ReplayMessenger(fn, trace=trace)
    if fn is None:
        return m

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        with m:
            return fn(*args, **kwargs)
    return wrapped


class ScaleMessenger(Messenger):
    def __init__(self, fn=None, scale=1.0):
        super().__init__(fn)
        self.scale = scale

    def _process_message(self, msg):
        if msg["type"] == "sample":
            msg["scale"] = self.scale * msg["scale"]


def scale(fn=None, scale=1.0):
    return ScaleMessenger(fn, scale=scale)


from . import enum_messenger  # noqa: E402,F401
from .enum_messenger import EnumMessenger  # noqa: E402,F401
