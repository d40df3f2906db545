"""auxSVItrainer: semi-supervised training loop with an auxiliary
classification loss (reference trainers/auxsvi.py:19-225).  Each batch runs
TWO fused CUDA optimisation steps sharing one Adam state: the (enumerated)
ELBO step and the auxiliary classifier step -- for unlabeled batches the
auxiliary step has no loss terms but Pyro still runs the optimizer, which is
reproduced (Adam on zero gradients)."""
from collections import OrderedDict
from copy import deepcopy as dc
from typing import Optional

import torch

from ..engine import SVIEngine
from ._pipeline import StepPipeline, run_epoch
from ..utils import average_weights, set_deterministic_mode


class auxSVItrainer:
    """
    Args:
        model: initialised ssiVAE
        task: "classification" (ssiVAE, enumerated labels) or "regression" (ss_reg_iVAE)
        optimizer: None or {"lr": ...} (Adam, default lr 5e-4)
        seed: reproducibility seed
    Keyword Args: lr (5e-4), device
    """

    def __init__(self, model, task: str = "classification", optimizer=None, seed: int = 1,
                 **kwargs) -> None:
        set_deterministic_mode(seed)
        if task not in ["classification", "regression"]:
            raise ValueError("Choose between 'classification' and 'regression' tasks")
        self.task = task
        self.device = kwargs.get("device", 'cuda' if torch.cuda.is_available() else 'cpu')
        lr = kwargs.get("lr", 5e-4)
        if isinstance(optimizer, dict):
            lr = optimizer.get("lr", lr)
        elif optimizer is not None:
            raise TypeError("pass optimizer=None or {'lr': ...}: Adam is fused into the CUDA step")
        self.svi = SVIEngine(model, lr=lr, enumerate_parallel=(task == "classification"), seed=seed,
                             device=self.device, force_generic=kwargs.get("force_generic"),
                             data_parallel=kwargs.get("data_parallel"))
        self.model = model
        self.history = {"training_loss": [], "test": []}
        self.current_epoch = 0
        self.running_weights = {}

    def compute_loss(self, xs: torch.Tensor, ys: Optional[torch.Tensor] = None, **kwargs) -> float:
        """basic (ELBO) step + auxiliary step (reference auxsvi.py:88-100)."""
        xs = xs.to(self.device, non_blocking=True)
        args = (xs,)
        if ys is not None:
            args = (xs, ys.to(self.device, non_blocking=True))
        loss = self.svi.step(*args, **kwargs)
        loss_aux = self.svi.step_aux(*args, **kwargs)
        return loss + loss_aux

    def train(self, loader_unsup, loader_sup, **kwargs) -> float:
        """One epoch; a labeled batch every p-th iteration (auxsvi.py:102-128).  Same schedule and
        the same two optimisation steps per batch as `compute_loss`, software-pipelined like
        SVItrainer.train (trainers/_pipeline.py): no host synchronisation per step."""
        sup_batches = len(loader_sup)
        unsup_batches = len(loader_unsup)
        p = (sup_batches + unsup_batches) // sup_batches
        eng = self.svi
        steps = [eng.step, eng.step_aux]
        counted = [0]

        def schedule():
            sup = iter(loader_sup)
            for i, (xs,) in enumerate(loader_unsup):
                counted[0] += xs.shape[0]
                yield (xs,), steps, 1.0            # loss + auxiliary loss enter the epoch loss
                if i % p == 1:
                    xs_l, ys_l = next(sup)
                    yield (xs_l, ys_l), steps, 0.0      # computed, not accumulated (auxsvi.py:126)

        with torch.cuda.device(eng.device):
            if not hasattr(self, "_pipe"):
                self._pipe = StepPipeline(eng)
            epoch_loss = run_epoch(self._pipe, schedule(), **kwargs)
        return epoch_loss / counted[0]

    def evaluate(self, loader_val) -> float:
        if self.task == "regression":
            # mean over validation batches of the per-batch MSE (reference auxsvi.py:151-162)
            acc, n = 0., 0
            for data, gt in loader_val:
                pred = self.model.regressor(data)
                acc += torch.nn.functional.mse_loss(pred, gt.cpu()).item()
                n += 1
            return acc / n
        correct, total = 0, 0
        for data, labels in loader_val:
            predicted = self.model.classifier(data)
            _, lab_idx = torch.max(labels.cpu(), 1)
            correct += (predicted == lab_idx).sum().item()
            total += data.size(0)
        return correct / total

    def step(self, loader_unsup, loader_sup, loader_val=None, **kwargs) -> None:
        train_loss = self.train(loader_unsup, loader_sup, **kwargs)
        self.history["training_loss"].append(train_loss)
        if loader_val is not None:
            self.history["test"].append(self.evaluate(loader_val))
        self.current_epoch += 1

    def save_running_weights(self, net: str) -> None:
        net = getattr(self.model, net)
        sd = OrderedDict()
        for k, v in net.state_dict().items():
            sd[k] = dc(v).cpu()
        self.running_weights[self.current_epoch] = sd

    def average_weights(self, net: str) -> None:
        net = getattr(self.model, net)
        net.load_state_dict(average_weights(self.running_weights))

    def print_statistics(self) -> None:
        e = self.current_epoch
        if len(self.history["test"]) > 0:
            template = ('Epoch: {} Training loss: {:.4f}, Test accuracy: {:.4f}'
                        if self.task == "classification" else
                        'Epoch: {} Training loss: {:.4f}, Test MSE: {:.4f}')
            print(template.format(e, self.history["training_loss"][-1], self.history["test"][-1]))
        else:
            template = 'Epoch: {} Training loss: {:.4f}'
            print(template.format(e, self.history["training_loss"][-1]))
