"""Flat parameter arena.

Every floating ``state_dict`` entry of a model (parameters first, then float
buffers such as BatchNorm running statistics) is re-homed into ONE contiguous
fp32 buffer at a fixed, 8-element-aligned offset; integer buffers
(``num_batches_tracked``) go to a small int64 side arena.  The module's
parameters/buffers become views, so ``state_dict()`` / ``load_state_dict()`` and
the checkpoint layout are unchanged (reference: the global model is addressed by
``state_dict`` keys, manager.py:78,123) while

* the optimizer is one kernel over ``theta[:n_param]`` (``ops.fused_sgd``),
* the FedAvg collective is one kernel over ``theta[:n]`` (parameters AND
  running statistics -- the reference averages every entry, manager.py:123),
* gradients (``grad``), momentum, the bf16 shadow weights consumed by the GEMMs
  (``theta_bf16``) and the frozen global copy used for delta uploads
  (``global_w``) are parallel flat buffers with identical offsets on every rank.

Conv weights keep their logical ``[Cout, Cin, KH, KW]`` shape with channels_last
strides, i.e. they are physically ``[Cout, KH, KW, Cin]`` in the arena.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
from torch import nn

ALIGN = 8  # elements: 16 B for bf16, 32 B for fp32 -> every view satisfies TMA / vector alignment


@dataclass
class Slot:
    name: str
    offset: int
    numel: int
    shape: Tuple[int, ...]
    channels_last: bool
    is_param: bool


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class ParamArena:
    def __init__(self, model: nn.Module, device=None, *, momentum: bool = False, bf16_shadow: bool = True,
                 keep_global: bool = True, theta_storage: Optional[torch.Tensor] = None, total_align: int = 2048):
        self.model = model
        params = [(n, p) for n, p in model.named_parameters()]
        device = torch.device(device) if device is not None else (params[0][1].device if params else torch.device("cpu"))
        self.device = device
        self.slots: "OrderedDict[str, Slot]" = OrderedDict()
        self.int_slots: "OrderedDict[str, Tuple[int, int]]" = OrderedDict()
        off = 0
        seen = set()
        for name, p in params:
            if id(p) in seen:
                continue
            seen.add(id(p))
            cl = p.dim() == 4
            self.slots[name] = Slot(name, off, p.numel(), tuple(p.shape), cl, True)
            off = _round_up(off + p.numel(), ALIGN)
        self.n_param = _round_up(off, ALIGN)
        off = self.n_param
        ioff = 0
        for name, b in model.named_buffers():
            if name.split(".")[-1] in getattr(self._owner(name), "_non_persistent_buffers_set", ()):
                continue
            if b.is_floating_point():
                self.slots[name] = Slot(name, off, b.numel(), tuple(b.shape), False, False)
                off = _round_up(off + b.numel(), ALIGN)
            else:
                self.int_slots[name] = (ioff, b.numel())
                ioff += b.numel()
        self.n = _round_up(max(off, ALIGN), total_align)   # padded so tiles / vectors never straddle the end
        self.n_int = ioff

        if theta_storage is not None:
            assert theta_storage.numel() >= self.n and theta_storage.dtype == torch.float32
            self.theta = theta_storage[: self.n]
            self.theta.zero_()
        else:
            self.theta = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.grad = torch.zeros(self.n_param, dtype=torch.float32, device=device)
        self.momentum = torch.zeros(self.n_param, dtype=torch.float32, device=device) if momentum else None
        self.theta_bf16 = torch.zeros(self.n, dtype=torch.bfloat16, device=device) if bf16_shadow else None
        self.global_w = torch.zeros(self.n, dtype=torch.float32, device=device) if keep_global else None
        self.int_arena = torch.zeros(max(self.n_int, 1), dtype=torch.int64, device=device)
        self._adopt()

    # ------------------------------------------------------------------
    def _owner(self, qualified: str) -> nn.Module:
        mod = self.model
        parts = qualified.split(".")[:-1]
        for p in parts:
            mod = getattr(mod, p)
        return mod

    def _view(self, flat: torch.Tensor, slot: Slot) -> torch.Tensor:
        v = flat[slot.offset: slot.offset + slot.numel]
        if slot.channels_last:
            co, ci, kh, kw = slot.shape
            return v.view(co, kh, kw, ci).permute(0, 3, 1, 2)
        return v.view(slot.shape)

    @torch.no_grad()
    def _adopt(self) -> None:
        for name, slot in self.slots.items():
            owner = self._owner(name)
            leaf = name.split(".")[-1]
            view = self._view(self.theta, slot)
            if slot.is_param:
                p = getattr(owner, leaf)
                view.copy_(p.detach().to(self.device))
                p.data = view
                p.grad = self._view(self.grad, slot)
                if self.theta_bf16 is not None:
                    sh = self.theta_bf16[slot.offset: slot.offset + slot.numel]
                    sh = sh.view(slot.shape[0], -1) if len(slot.shape) >= 2 else sh.view(slot.shape)
                    # plain attribute (not a registered buffer): stays out of state_dict
                    object.__setattr__(owner, leaf + "_bf16", sh)
            else:
                b = getattr(owner, leaf)
                view.copy_(b.detach().to(self.device))
                owner._buffers[leaf] = view
        for name, (ioff, n) in self.int_slots.items():
            owner = self._owner(name)
            leaf = name.split(".")[-1]
            b = getattr(owner, leaf)
            view = self.int_arena[ioff: ioff + n].view(b.shape)
            view.copy_(b.detach().to(self.device))
            owner._buffers[leaf] = view
        self.sync_shadow()
        if self.global_w is not None:
            self.global_w.copy_(self.theta)

    # ------------------------------------------------------------------
    @torch.no_grad()
    def sync_shadow(self) -> None:
        """Refresh the bf16 shadow from the fp32 master (after load_state_dict etc.)."""
        if self.theta_bf16 is None:
            return
        if self.theta.is_cuda:
            from ..ops import functional as F
            F.cast(self.theta, torch.bfloat16, out=self.theta_bf16)
        else:
            self.theta_bf16.copy_(self.theta.to(torch.bfloat16))

    @torch.no_grad()
    def commit_global(self) -> None:
        """Declare the current weights to be the global model (start of training / after a
        manual ``load_state_dict``)."""
        if self.global_w is not None:
            self.global_w.copy_(self.theta)
        self.sync_shadow()

    def zero_grad(self) -> None:
        self.grad.zero_()

    def first_weight_slot(self) -> Optional[Slot]:
        for s in self.slots.values():
            if s.is_param and len(s.shape) >= 2:
                return s
        return None

    def nbytes(self) -> Dict[str, int]:
        out = {"theta": self.theta.numel() * 4, "grad": self.grad.numel() * 4}
        if self.momentum is not None:
            out["momentum"] = self.momentum.numel() * 4
        if self.theta_bf16 is not None:
            out["theta_bf16"] = self.theta_bf16.numel() * 2
        if self.global_w is not None:
            out["global_w"] = self.global_w.numel() * 4
        return out

    def describe(self) -> str:
        return "ParamArena(n={}, n_param={}, n_int={}, tensors={}, device={})".format(
            self.n, self.n_param, self.n_int, len(self.slots), self.device)
