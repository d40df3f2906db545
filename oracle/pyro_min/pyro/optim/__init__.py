"""PyroOptim restatement: one torch optimizer PER PARAMETER, created lazily on
first sight (Pyro semantics; call site reference trainers/svi.py:79-81).
Oracle-only."""
import torch


class PyroOptim:
    def __init__(self, optim_constructor, optim_args):
        self.ctor = optim_constructor
        self.args = optim_args
        self.optim_objs = {}

    def __call__(self, params):
        for p in params:
            if p not in self.optim_objs:
                self.optim_objs[p] = self.ctor([p], **self.args)
            self.optim_objs[p].step()


def Adam(optim_args):
    return PyroOptim(torch.optim.Adam, optim_args)
