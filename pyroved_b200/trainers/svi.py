"""SVItrainer: epoch loop around the fused CUDA SVI step
(reference trainers/svi.py:11-175)."""
from collections import deque
from typing import Optional

import torch

from ..engine import SVIEngine
from ..utils import set_deterministic_mode


class SVItrainer:
    """
    Args:
        model: an initialised pyroved_b200 model (iVAE, jiVAE, VED, ...)
        optimizer: None (Adam) -- or a dict like {"lr": 1e-3}
        loss: None (Trace_ELBO, or TraceEnum_ELBO when enumerate_parallel)
        enumerate_parallel: exact enumeration of discrete latents (jiVAE)
        seed: reproducibility seed
    Keyword Args:
        lr (1e-3), device
    """

    def __init__(self, model, optimizer=None, loss=None, enumerate_parallel: bool = False,
                 seed: int = 1, **kwargs) -> None:
        set_deterministic_mode(seed)
        self.device = kwargs.get("device", 'cuda' if torch.cuda.is_available() else 'cpu')
        lr = kwargs.get("lr", 1e-3)
        if isinstance(optimizer, dict):
            lr = optimizer.get("lr", lr)
        elif optimizer is not None:
            raise TypeError(
                "pyroved_b200 fuses Adam into the CUDA step; pass optimizer=None or {'lr': ...}")
        if loss is not None:
            raise TypeError("pyroved_b200 implements Trace_ELBO / TraceEnum_ELBO natively; "
                            "pass loss=None (use enumerate_parallel for discrete latents)")
        self.svi = SVIEngine(model, lr=lr, enumerate_parallel=enumerate_parallel, seed=seed,
                             device=self.device)
        self.loss_history = {"training_loss": [], "test_loss": []}
        self.current_epoch = 0

    def train(self, train_loader, **kwargs) -> float:
        """One epoch; returns loss / number of samples (reference svi.py:95-115).

        Same per-batch work as the reference loop (host batch -> device, one SVI step, the
        step's loss back to the host), software-pipelined: the H2D copy of batch i+1 runs on a
        copy stream under the kernels of batch i, and the 4-byte loss of step i is read from a
        pinned ring two steps later, so neither transfer stalls the launch thread."""
        eng = self.svi
        dev = eng.device
        main = torch.cuda.current_stream(dev)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(dev)
            self._stage = {}
            self._stage_free = [None, None]
        cs = self._copy_stream

        def upload(data, slot):
            bufs = []
            with torch.cuda.stream(cs):
                if self._stage_free[slot] is not None:
                    cs.wait_event(self._stage_free[slot])    # step that read this slot is done
                for j, t in enumerate(data):
                    key = (slot, j, tuple(t.shape), t.dtype)
                    b = self._stage.get(key)
                    if b is None:
                        b = self._stage[key] = torch.empty(t.shape, dtype=t.dtype, device=dev)
                    b.copy_(t, non_blocking=True)
                    bufs.append(b)
                ev = torch.cuda.Event()
                ev.record(cs)
            return bufs, ev

        epoch_loss = 0.
        pending = deque()     # (ring index, event, host constant)
        it = iter(train_loader)
        nxt = next(it, None)
        staged = upload(nxt, 0) if nxt is not None else None
        i = 0
        while staged is not None:
            bufs, ev = staged
            nxt = next(it, None)
            staged = upload(nxt, (i + 1) % 2) if nxt is not None else None
            main.wait_event(ev)
            # the staging slot's address recurs: its copy into the program input is part of the
            # step's CUDA graph; the optimizer kernel writes the loss into eng.loss_ring (pinned
            # host memory, slot = optimizer step count & 3) -- no copy call on either side
            eng.step(*bufs, _sync=False, _static=True, **kwargs)
            free = torch.cuda.Event()
            free.record(main)
            self._stage_free[i % 2] = free
            if len(pending) >= 3:                       # ring slot about to be reused
                k, e, c = pending.popleft()
                e.synchronize()
                epoch_loss += float(eng.loss_ring[k]) + c
            pending.append((eng.updates_done & 3, free, eng.last_loss_const))
            i += 1
        while pending:
            k, e, c = pending.popleft()
            e.synchronize()
            epoch_loss += float(eng.loss_ring[k]) + c
        return epoch_loss / len(train_loader.dataset)

    def evaluate(self, test_loader, **kwargs) -> float:
        """Reference behaviour (svi.py:117-137): `svi.step` under no_grad, i.e.
        no backward pass, but Pyro still runs the optimizer on zeroed
        gradients, so Adam's momentum keeps moving the weights.  Reproduced
        here (forward-only kernels + Adam with g = 0)."""
        test_loss = 0.
        for data in test_loader:
            args = [t.to(self.device, non_blocking=True) for t in data]
            test_loss += self.svi._step(tuple(args), dict(kwargs), train=False, update=True)
        return test_loss / len(test_loader.dataset)

    def step(self, train_loader, test_loader: Optional[object] = None, **kwargs) -> None:
        train_loss = self.train(train_loader, **kwargs)
        self.loss_history["training_loss"].append(train_loss)
        if test_loader is not None:
            test_loss = self.evaluate(test_loader, **kwargs)
            self.loss_history["test_loss"].append(test_loss)
        self.current_epoch += 1

    def print_statistics(self) -> None:
        e = self.current_epoch
        if len(self.loss_history["test_loss"]) > 0:
            template = 'Epoch: {} Training loss: {:.4f}, Test loss: {:.4f}'
            print(template.format(e, self.loss_history["training_loss"][-1],
                                  self.loss_history["test_loss"][-1]))
        else:
            template = 'Epoch: {} Training loss: {:.4f}'
            print(template.format(e, self.loss_history["training_loss"][-1]))
