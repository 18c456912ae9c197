"""SPMD federated engine: one process per GPU, every rank is a federated client.

This is the in-process counterpart of the HTTP control plane for jobs launched
with ``torchrun`` on one NVSwitch box (the BASELINE.json ResNet / BERT configs).
A *round* on every rank is

    (host -> device copy of this round's private shard, from pinned memory)
    local SGD for ``n_epoch`` epochs        (reference worker.py:103-106, demo.py:29-49)
    fused weighted reduce + broadcast + apply over NVLink   (manager.py:113-126 + :77-86)
    (device -> host read of the per-epoch losses)

with no host-side exchange between ranks: the sample counts n_k travel on the
collective's barrier flags.  Client sampling / logical clients stay in Python
(BASELINE.json: "client sampling and round bookkeeping stay in Python"): with
``logical_clients > world`` each rank time-slices several logical clients and
folds their sample-weighted deltas locally before the cross-GPU reduce; the
per-round draw is a seeded ``random.Random`` shared by all ranks, so no
communication is needed to agree on the participants.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from ..metrics import phase
from ..train import GraphedLocalSGD, PortableLocalSGD
from .arena import ParamArena
from .fedavg import FedAvgSession, NcclSession


@dataclass
class RoundResult:
    update_name: str
    n_samples: int
    loss_history: List[float]
    participants: List[int] = field(default_factory=list)
    global_loss: Optional[List[float]] = None


class FederatedEngine:
    def __init__(self, model, device, *, backend: str = "fused", group=None, loss: str = "ce",
                 lr: float = 0.05, batch_size: int = 128, momentum: float = 0.0, weight_decay: float = 0.0,
                 wire_dtype: str = "bf16", mode: str = "delta", n_ctas: int = 148, use_graph: bool = True,
                 logical_clients: int = 0, sample_k: Optional[int] = None, seed: int = 0, name: str = "exp",
                 nvls: "bool | str" = "auto", tile_flags: bool = False):
        self.device = torch.device(device)
        self.model = model
        self.name = name
        self.arena = ParamArena(model, self.device, momentum=momentum > 0)
        if hasattr(model, "build_workspace"):
            model.build_workspace(self.device)
        if self.device.type == "cuda":
            self.trainer = GraphedLocalSGD(model, self.arena, loss=loss, use_graph=use_graph)
            model._graphed_trainer = self.trainer
        else:
            # CPU / gloo: the plumbing configuration -- same engine, portable PyTorch training loop, and the
            # torch.distributed session (the fused collective needs NVLink peer memory)
            if backend == "fused":
                raise ValueError("backend='fused' needs CUDA devices; use backend='nccl' (torch.distributed, "
                                 "gloo on CPU) for CPU runs")
            self.trainer = PortableLocalSGD(model, self.arena, loss=loss)
        Session = {"fused": FedAvgSession, "nccl": NcclSession}[backend]
        self.session = Session(self.arena, group, wire_dtype=wire_dtype, mode=mode, n_ctas=n_ctas, nvls=nvls,
                               tile_flags=tile_flags)
        self.backend = backend
        self.rank, self.world = self.session.rank, self.session.world
        # K4: the last SGD step of the captured epoch writes the upload copy itself (no pack phase in the collective);
        # only for the plain one-client-per-GPU rounds -- logical clients fold their deltas after training
        self.prepack = (backend == "fused" and self.device.type == "cuda" and not (logical_clients and logical_clients > self.world)
                        and __import__("os").environ.get("BATON_PREPACK", "1") != "0")
        if self.prepack and hasattr(self.session, "pack_spec"):
            self.trainer.pack = self.session.pack_spec()
        # the round-end collective runs on the session's high-priority side stream: the NEXT round's host->device shard
        # copy (and anything else that does not touch the arena) overlaps it; local training joins first
        self.overlap_collective = (backend == "fused" and self.device.type == "cuda"
                                   and __import__("os").environ.get("BATON_COLLECTIVE_OVERLAP", "1") != "0")
        # K3 (bcast_gemm) on the flagship path: the first convolution's weight staging + GEMM acquire the collective's
        # arrival flags, and the head of the next round's captured epoch runs while the collective is still in flight
        self.k3 = bool(self.overlap_collective and tile_flags and hasattr(model, "conv1")
                       and isinstance(self.session, FedAvgSession) and hasattr(self.session, "gate_first_conv")
                       and hasattr(self.trainer, "k3_join"))
        if self.k3:
            self.session.gate_first_conv(model.conv1)
            self.trainer.k3_join = self.sync
        self.hp = dict(lr=lr, batch_size=batch_size, momentum=momentum, weight_decay=weight_decay)
        self.n_rounds = 0
        self.logical_clients = logical_clients if logical_clients and logical_clients > self.world else 0
        self.sample_k = sample_k
        self._rng = random.Random(seed)            # identical stream on every rank
        self._stage: Dict[Tuple, Tuple[torch.Tensor, torch.Tensor]] = {}
        self._acc = None
        self.last_losses_dev = None
        self.samples_trained = 0          # samples this rank pushed through local SGD (per epoch)
        self.phase_s: Dict[str, float] = {}   # host seconds per NVTX phase (launch cost; device time is in bench.py)

    # ------------------------------------------------------------------ data staging
    def stage(self, X_host: torch.Tensor, y_host: torch.Tensor, slot: int = 0):
        """Asynchronous host->device copy of a shard into persistent staging buffers (so the
        captured epoch graph keeps pointing at the same addresses).  ``X_host``/``y_host`` should
        be pinned.  Returns the device views."""
        key = (slot, tuple(X_host.shape), X_host.dtype, tuple(y_host.shape), y_host.dtype)
        bufs = self._stage.get(key)
        if bufs is None:
            bufs = (torch.empty(X_host.shape, dtype=X_host.dtype, device=self.device),
                    torch.empty(y_host.shape, dtype=y_host.dtype, device=self.device))
            self._stage[key] = bufs
        bufs[0].copy_(X_host, non_blocking=True)
        bufs[1].copy_(y_host, non_blocking=True)
        return bufs

    @staticmethod
    def h2d_bytes(X_host: torch.Tensor, y_host: torch.Tensor) -> int:
        return X_host.numel() * X_host.element_size() + y_host.numel() * y_host.element_size()

    # ------------------------------------------------------------------ participants
    def draw_participants(self) -> List[int]:
        """Logical client ids taking part in this round (same on every rank)."""
        total = self.logical_clients or self.world
        ids = list(range(total))
        if self.sample_k is None or self.sample_k >= total:
            return ids
        return sorted(self._rng.sample(ids, self.sample_k))

    def hosted(self, client_id: int) -> bool:
        return client_id % self.world == self.rank

    # ------------------------------------------------------------------ one round
    def run_round(self, shards, n_epoch: int = 1, read_loss: bool = True) -> RoundResult:
        """``shards``: for a plain run a ``(X, y)`` pair (device tensors, or pinned host tensors
        that are staged first); with logical clients a callable ``client_id -> (X, y)``."""
        update_name = "update_{}_{:05d}".format(self.name, self.n_rounds)
        participants = self.draw_participants()
        mine = [c for c in participants if self.hosted(c)]
        a = self.arena
        total_n = 0
        losses_dev = None
        if not self.logical_clients:
            if mine:
                X, y = shards(self.rank) if callable(shards) else shards
                if not X.is_cuda:
                    with phase("baton.h2d_shard", self.phase_s):
                        X, y = self.stage(X, y)
                if not self.k3:
                    self.sync()              # the previous round's collective must have landed before training reads theta
                if self.prepack and self.trainer.pack is not None:
                    self.session.arm_prepack(float(X.shape[0]))
                with phase("baton.local_train", self.phase_s):
                    losses_dev = self.trainer.run(X, y, n_epoch=n_epoch, return_device=True, **self.hp)
                total_n = X.shape[0]
        else:
            # time-sliced logical clients: fold n_k * (theta_k - global) locally, then upload the mean
            self.sync()
            if len(mine) > 1 and self._acc is None:
                self._acc = torch.zeros_like(a.theta)
            if len(mine) > 1 and not a.theta.is_cuda:
                self._acc.zero_()
            for j, cid in enumerate(mine):
                X, y = shards(cid)
                if not X.is_cuda:
                    X, y = self.stage(X, y, slot=j)
                ld = self.trainer.run(X, y, n_epoch=n_epoch, return_device=True, **self.hp)
                nk = X.shape[0]
                losses_dev = ld * nk if losses_dev is None else losses_dev + ld * nk
                total_n += nk
                if len(mine) > 1:
                    more = j + 1 < len(mine)           # the next co-resident client starts from the global model
                    if a.theta.is_cuda:
                        from ..ops import functional as F     # ONE kernel: fold the delta + reset the replica
                        F.fold_client(self._acc, a.theta, a.global_w, nk, first=(j == 0), reset=more,
                                      w_bf16=a.theta_bf16, momentum=a.momentum)
                    else:
                        self._acc.add_(a.theta - a.global_w, alpha=float(nk))
                        if more:
                            a.theta.copy_(a.global_w)
                            a.sync_shadow()
                            if a.momentum is not None:
                                a.momentum.zero_()
            if len(mine) > 1:
                if a.theta.is_cuda:
                    from ..ops import functional as F
                    F.fold_finish(self._acc, a.theta, a.global_w, total_n)
                else:
                    torch.add(a.global_w, self._acc, alpha=1.0 / total_n, out=a.theta)
            if losses_dev is not None and total_n:
                losses_dev = losses_dev / total_n
        self.last_losses_dev = losses_dev
        self.samples_trained += int(total_n)
        loss_for_wire = None
        if losses_dev is not None:
            steps = max(1, self.trainer.last_steps)
            loss_for_wire = losses_dev[:, 0] / steps
        with phase("baton.aggregate_broadcast", self.phase_s):
            self._aggregate(float(total_n), loss_for_wire)
        self.n_rounds += 1
        hist: List[float] = []
        if read_loss and losses_dev is not None:
            hist = loss_for_wire.tolist()       # device -> host read of the round's result
        return RoundResult(update_name, int(total_n), hist, participants)

    def sync(self) -> None:
        """Make the compute stream wait for a collective that is still running on the side stream (call before
        anything reads or writes the arena: training, ``state_dict()``, checkpoints)."""
        join = getattr(self.session, "join", None)
        if join is not None:
            join()

    def _aggregate(self, my_n: float, loss_dev) -> None:
        s = self.session
        side = bool(self.overlap_collective and isinstance(s, FedAvgSession))
        if loss_dev is not None and hasattr(s, "loss_local"):
            k = min(loss_dev.numel(), s.loss_local.numel())
            s.loss_local.zero_()
            s.loss_local[:k].copy_(loss_dev[:k])
            pre = bool(self.prepack and getattr(self.trainer, "emitted_wire", False) and my_n > 0
                       and not getattr(self.trainer, "last_had_tail_step", False))
            s.aggregate(my_n=my_n, prepacked=pre, on_side_stream=side) if isinstance(s, FedAvgSession) else s.aggregate(my_n=my_n)
        elif loss_dev is not None:
            s.aggregate(my_n=my_n, loss_history=loss_dev.tolist())
        else:
            s.aggregate(my_n=my_n)

    def global_loss(self, n_epoch: int) -> List[float]:
        return self.session.reduced_loss(n_epoch)

    def state_dict(self):
        self.sync()
        return self.model.state_dict()
