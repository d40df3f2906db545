"""SVItrainer: epoch loop around the fused CUDA SVI step
(reference trainers/svi.py:11-175)."""
from typing import Optional

import torch

from ..engine import SVIEngine
from ._pipeline import StepPipeline, run_epoch
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
                             device=self.device, force_generic=kwargs.get("force_generic"),
                             data_parallel=kwargs.get("data_parallel"))
        self.loss_history = {"training_loss": [], "test_loss": []}
        self.current_epoch = 0

    def train(self, train_loader, **kwargs) -> float:
        """One epoch; returns loss / number of samples (reference svi.py:95-115).

        Same per-batch work as the reference loop (host batch -> device, one SVI step, the
        step's loss back to the host), software-pipelined (trainers/_pipeline.py): the H2D copy
        of batch i+1 runs on a copy stream under the kernels of batch i, and the 4-byte loss of
        step i is read from a pinned ring a few steps later, so neither transfer stalls the
        launch thread.  Batches of a GPU-resident loader (utils.DeviceBatchLoader) are used in
        place."""
        eng = self.svi
        with torch.cuda.device(eng.device):
            if not hasattr(self, "_pipe"):
                self._pipe = StepPipeline(eng)
            steps = [eng.step]
            epoch_loss = run_epoch(self._pipe, ((data, steps, 1.0) for data in train_loader),
                                   **kwargs)
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
