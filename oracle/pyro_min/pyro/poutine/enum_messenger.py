"""Parallel enumeration of discrete guide sites (oracle-only restatement).

Pyro places the enumerated support of a site marked
infer={'enumerate': 'parallel'} on a fresh tensor dim to the LEFT of all
plate dims, starting at `first_available_dim` (-1 - max_plate_nesting; the
reference uses max_plate_nesting=1 -> -2: trainers/svi.py:84-86).  With
expand=True the value is expanded over the batch shape, e.g. a
OneHotCategorical with batch [B] and K classes yields a value of shape
[K, B, K] (reference tests/test_models.py:193-207).
"""
from . import Messenger


class EnumMessenger(Messenger):
    def __init__(self, fn=None, first_available_dim=None):
        super().__init__(fn)
        self.first_available_dim = first_available_dim
        self._next = first_available_dim

    def __enter__(self):
        self._next = self.first_available_dim
        return super().__enter__()

    def _process_message(self, msg):
        if msg["type"] != "sample" or msg["done"] or msg["is_observed"]:
            return
        if msg["infer"].get("enumerate") != "parallel":
            return
        fn = msg["fn"]
        if self._next is None:
            raise ValueError("EnumMessenger needs first_available_dim")
        dim = self._next
        self._next -= 1
        expand = msg["infer"].get("expand", False)
        value = fn.enumerate_support(expand=expand)  # [K, *batch, *event]
        # move the leading enumeration dim to position `dim` counted from the
        # right of the BATCH shape (event dims stay rightmost)
        batch_ndim = len(fn.batch_shape)
        event_ndim = len(fn.event_shape)
        pad = (-dim) - batch_ndim - 1
        if pad > 0:
            k = value.shape[0]
            value = value.reshape((k,) + (1,) * pad + value.shape[1:])
        msg["value"] = value
        msg["infer"]["_enumerate_dim"] = dim
        msg["infer"]["_event_ndim"] = event_ndim
        msg["done"] = True
