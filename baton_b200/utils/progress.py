"""Epoch iterator with a running loss.

Parity target: ``EpochProgress`` (reference utils.py:70-90): wrap the batch
iterable of one epoch, optionally show a tqdm bar, expose ``update_loss`` and a
``loss`` attribute read at the end of the epoch (demo.py:38-48).

Deliberate changes:
  * the running value is a true arithmetic mean (the reference recurrence
    ``loss*i/(i+1) + l/i`` over-reports: a constant 4.0 reads 4.0, 4.667, ...),
  * a CUDA tensor loss is accumulated on the device and only read back when
    ``loss`` is queried (once per epoch) instead of ``float(loss)`` per batch.
"""
from __future__ import annotations

from collections.abc import Iterator
from typing import Any, Iterable

try:  # tqdm is optional at runtime; the bar is cosmetic
    from tqdm import tqdm
except Exception:  # pragma: no cover
    tqdm = None


class EpochProgress(Iterator):
    def __init__(self, epoch: int, _iter: Iterable[Any], verbose: bool = True,
                 postfix_every: int = 16):
        self.verbose = bool(verbose) and tqdm is not None
        self.epoch = epoch
        if self.verbose:
            self.pbar = tqdm(_iter, desc="Epoch {}".format(epoch))
        else:
            self.pbar = _iter
        self.pbar_iter = iter(self.pbar)
        self.N = 0                 # batches handed out
        self._n_loss = 0           # losses folded in
        self._host_sum = 0.0       # python-number losses
        self._dev_sum = None       # lazily created device/CPU tensor accumulator
        self._postfix_every = max(1, int(postfix_every))

    def __next__(self):
        item = next(self.pbar_iter)
        self.N += 1
        return item

    def update_loss(self, loss) -> None:
        self._n_loss += 1
        if hasattr(loss, "detach"):
            val = loss.detach()
            if val.dim() != 0:
                val = val.mean()
            val = val.float()
            self._dev_sum = val.clone() if self._dev_sum is None else self._dev_sum.add_(val)
        else:
            self._host_sum += float(loss)
        if self.verbose and (self._n_loss % self._postfix_every == 0):
            self.pbar.set_postfix({"loss": "{:0.4f}".format(self.loss)})

    @property
    def loss(self) -> float:
        """Mean of every loss passed to ``update_loss`` (one device read)."""
        if self._n_loss == 0:
            return 0.0
        total = self._host_sum
        if self._dev_sum is not None:
            total += float(self._dev_sum)
        return total / self._n_loss
