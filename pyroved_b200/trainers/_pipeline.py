"""Software pipeline shared by the epoch loops of SVItrainer and auxSVItrainer.

Per batch the reference does: host batch -> device, one or two SVI steps, each step's loss back to
the host as a Python float (trainers/svi.py:104-113, trainers/auxsvi.py:88-100).  Same work here,
overlapped:
  * the H2D copy of batch i+1 runs on a copy stream under the kernels of batch i, into one of a
    few staging slots whose addresses recur, so the copy from the slot into the step program's
    input buffer is a node of the step's CUDA graph (graphs are keyed by the slot address);
  * batches that already live in HBM (utils.DeviceBatchLoader) skip the staging copy;
  * the optimizer kernel writes every step's loss into a ring in pinned host memory
    (slot = optimizer step count & (LOSS_RING - 1)); the loop reads a slot a few steps later,
    after an event, so no step waits for a device -> host round trip.
"""
from collections import deque

import torch

from .. import ops

SLOTS = 4            # staging slots (two batches may be in flight: unlabelled + labelled)
MAX_LAG = 8          # ring entries kept pending before the oldest is read (< LOSS_RING)


class StepPipeline:
    def __init__(self, engine):
        self.eng = engine
        self.dev = engine.device
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.stage = {}
        self.slot_free = [None] * SLOTS
        self.next_slot = 0
        self.pending = deque()        # (ring index, event, host constant, weight)
        self.total = 0.0

    # ---- input side -----------------------------------------------------------------------------
    def upload(self, data):
        """Start moving one batch (tuple of tensors) towards the device.  Returns (bufs, event,
        slot); event is None for device-resident batches."""
        if all(t.is_cuda for t in data):
            return list(data), None, None
        slot = self.next_slot
        self.next_slot = (slot + 1) % SLOTS
        cs = self.copy_stream
        bufs = []
        with torch.cuda.stream(cs):
            if self.slot_free[slot] is not None:
                cs.wait_event(self.slot_free[slot])     # the step that read this slot is done
            for j, t in enumerate(data):
                key = (slot, j, tuple(t.shape), t.dtype)
                b = self.stage.get(key)
                if b is None:
                    b = self.stage[key] = torch.empty(t.shape, dtype=t.dtype, device=self.dev)
                b.copy_(t, non_blocking=True)
                bufs.append(b)
            ev = torch.cuda.Event()
            ev.record(cs)
        return bufs, ev, slot

    # ---- step side ------------------------------------------------------------------------------
    def run(self, staged, steps, weight=1.0, **kwargs):
        """Issue `steps` (engine methods, e.g. [eng.step] or [eng.step, eng.step_aux]) on a staged
        batch; their losses are collected (times `weight`) as they arrive."""
        bufs, ev, slot = staged
        main = torch.cuda.current_stream(self.dev)
        if ev is not None:
            main.wait_event(ev)
        for fn in steps:
            fn(*bufs, _sync=False, _static=True, **kwargs)
            done = torch.cuda.Event()
            done.record(main)
            self.pending.append((self.eng.updates_done & (ops.LOSS_RING - 1), done,
                                 self.eng.last_loss_const, weight))
            while len(self.pending) > MAX_LAG:
                self._pop()
        if slot is not None:
            self.slot_free[slot] = done

    def _pop(self):
        k, e, c, w = self.pending.popleft()
        e.synchronize()
        if w:
            self.total += w * (float(self.eng.loss_ring[k]) + c)

    def drain(self):
        """Wait for the outstanding steps; returns the accumulated (weighted) loss and resets it."""
        while self.pending:
            self._pop()
        t, self.total = self.total, 0.0
        return t


def run_epoch(pipe, items, **kwargs):
    """items: iterable of (data tuple, [engine step methods], weight).  One batch is uploaded
    ahead of the one being computed.  Returns the weighted sum of the step losses."""
    it = iter(items)
    cur = next(it, None)
    staged = pipe.upload(cur[0]) if cur is not None else None
    while cur is not None:
        nxt = next(it, None)
        staged_next = pipe.upload(nxt[0]) if nxt is not None else None
        pipe.run(staged, cur[1], cur[2], **kwargs)
        cur, staged = nxt, staged_next
    return pipe.drain()
