"""Local-SGD loop run by every federated client between two aggregations.

Parity target: demo ``Model.train`` (reference demo.py:29-49): one ``randperm``
per call (demo.py:33, kept as the default -- quirk 12), plain SGD, per epoch a
pass over ``torch.split(idxs, batch_size)`` with zero_grad / gather / forward /
loss / backward / step, returning one running-mean loss per epoch.

Two executions of the same contract:

* ``run_local_sgd`` -- portable PyTorch loop (CPU, gloo plumbing config, test
  oracle).  The loss is accumulated in a tensor and read once per epoch (the
  reference does ``float(loss)`` per batch, utils.py:88).
* ``GraphedLocalSGD`` (CUDA) -- the whole step (on-device batch gather, forward,
  loss, backward, fused arena SGD, loss accumulation) is captured once into a
  CUDA graph and replayed per batch; parameters, gradients and momentum live in
  the flat arena so the optimizer is one kernel (``ops.fused_sgd``) instead of
  one launch per tensor.  No host synchronisation inside an epoch.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
from torch import nn

from .utils.progress import EpochProgress


def _loss_fn(kind):
    if callable(kind):
        return kind
    if kind == "mse":
        return nn.functional.mse_loss
    if kind in ("ce", "cross_entropy"):
        return nn.functional.cross_entropy
    raise ValueError("unknown loss {!r}".format(kind))


def run_local_sgd(model: nn.Module, X: torch.Tensor, y: torch.Tensor, *, n_epoch: int = 32,
                  lr: float = 0.001, batch_size: int = 32, momentum: float = 0.0,
                  weight_decay: float = 0.0, loss: "str | Callable" = "mse",
                  verbose: bool = False, reshuffle_each_epoch: bool = False,
                  generator: Optional[torch.Generator] = None) -> List[float]:
    """Portable local SGD; returns the per-epoch mean loss."""
    criterion = _loss_fn(loss)
    n = X.shape[0]
    nn.Module.train(model, True)
    optimizer = torch.optim.SGD(model.parameters(), lr=lr, momentum=momentum,
                                weight_decay=weight_decay)
    idxs = torch.randperm(n, generator=generator).to(X.device)
    loss_history: List[float] = []
    for epoch in range(n_epoch):
        if reshuffle_each_epoch and epoch > 0:
            idxs = torch.randperm(n, generator=generator).to(X.device)
        batch_iter = EpochProgress(epoch, torch.split(idxs, batch_size), verbose=verbose)
        for batch_idxs in batch_iter:
            optimizer.zero_grad(set_to_none=True)
            output = model(X[batch_idxs])
            target = y[batch_idxs]
            if output.shape != target.shape and target.dtype.is_floating_point:
                target = target.reshape(output.shape)  # (N,) vs (N,1) -- quirk 13
            loss_batch = criterion(output, target)
            batch_iter.update_loss(loss_batch)
            loss_batch.backward()
            optimizer.step()
        loss_history.append(batch_iter.loss)
    return loss_history
