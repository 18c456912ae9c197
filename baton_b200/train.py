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

from typing import Callable, List, Optional

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


class GraphedLocalSGD:
    """CUDA local-SGD engine for an arena-adopted model.

    One *epoch* -- ``n // batch_size`` steps of {on-device batch gather, forward,
    fused loss, hand-written backward, ONE fused SGD kernel over the arena, loss
    accumulation} -- is captured into a single CUDA graph and replayed once per
    epoch, so the host issues one launch per epoch and never synchronises inside
    a round: the per-epoch losses are read back together when the round ends.

    ``model`` must already be adopted by a :class:`~baton_b200.parallel.arena.ParamArena`
    (``arena``); the engine is what ``FederatedModule.local_train`` dispatches to
    for CUDA shards (``model._graphed_trainer``).
    """

    def __init__(self, model: nn.Module, arena, *, loss: str = "ce", nesterov: bool = False,
                 use_graph: bool = True, input_dtype=torch.bfloat16):
        from .ops import functional as F
        from .ops import nn as bnn
        self.F, self.bnn = F, bnn
        self.model, self.arena = model, arena
        self.loss_kind = loss
        self.nesterov = nesterov
        self.use_graph = use_graph
        self.input_dtype = input_dtype
        import os
        self.explicit = os.environ.get("BATON_EXPLICIT_STEP", "1") != "0"   # models that offer a hand-scheduled step
        # optimizer slice of the deep layers beside the rest of the backward pass: implemented and validated, but measured
        # neutral on B200 (the HBM-bound slice slows the latency-bound kernels it runs beside by as much as it
        # hides: 4.52 vs 4.54 ms per 8 steps, profiles/r2_step_experiments.txt) -> opt-in
        self.tail_overlap = os.environ.get("BATON_SGD_OVERLAP", "0") == "1"
        self.tail_ctas = int(os.environ.get("BATON_SGD_TAIL_CTAS", "148"))    # grid cap of the overlapped SGD slice
        self.k3_join = None           # set by the engine: callable joining the round-end collective (enables the graph split)
        self._first_gemm_hook = None
        self.pack = None              # set by the engine: FedAvgSession.pack_spec() -> last SGD step emits the upload copy
        self.emitted_wire = False
        self.graph_emits_wire = False
        self._split = None
        self._split_active = 0
        self._tail_stream = None
        self._tail_pending = False
        self._tail_done = False
        dev = arena.device
        self.device = dev
        self.hyper = torch.zeros(4, dtype=torch.float32, device=dev)
        self.loss_acc = torch.zeros(2, dtype=torch.float32, device=dev)
        self._graphs = {}           # (n, batch, x_shape, y_shape) -> captured epoch
        self._hyper_host = None
        self.n_kernels_per_step = None
        self.last_stats = {}

    # -------------------------------------------------------------- one SGD step (capturable)
    def _loss(self, out, yb):
        if self.loss_kind in ("ce", "cross_entropy"):
            loss, stats = self.bnn.cross_entropy(out, yb)
            return loss, stats
        loss = self.bnn.mse_loss(out, yb)
        return loss, torch.stack([loss.detach(), torch.zeros_like(loss.detach())])

    def _gather(self, X, y, idx):
        F = self.F
        xb = F.gather_rows(X, idx)
        yb = F.gather_rows(y, idx) if y.dtype == torch.int64 and y.dim() == 1 else y.index_select(0, idx)
        return xb, yb

    def _step(self, X, y, idx, batch=None, emit_wire=False):
        """One SGD step on ``X[idx], y[idx]`` (or on the already gathered ``batch``).  ``emit_wire``: last step of
        an epoch -- the optimizer kernel also writes the upload copy for the round-end collective (``self.pack``)."""
        F = self.F
        xb, yb = batch if batch is not None else self._gather(X, y, idx)
        ws = getattr(self.model, "stats_workspace", None)
        if ws is not None and not getattr(self.model, "zeroes_own_workspace", False):
            ws.zero_()
        explicit = getattr(self.model, "explicit_step", None) if self.explicit else None
        if getattr(self.model, "compute_dtype", "bf16") != "bf16":
            explicit = None            # the hand-scheduled step drives the bf16 conv kernels; MXFP8 convs go through autograd
        a = self.arena
        bf = a.theta_bf16
        if explicit is not None and self.loss_kind in ("ce", "cross_entropy"):
            # hand-scheduled forward + loss + backward (no autograd engine): two-piece block gradients, parallel shortcut
            # branch; the loss kernel accumulates straight into the epoch's running sums.  The optimizer step of the deep
            # layers (their gradients are complete early in the backward pass) runs on a side stream beside the rest
            # of the backward pass and the first layers of the NEXT step's forward.
            # the epoch's last step emits the upload copy from ONE optimizer launch over the whole arena: no split there
            split = 0 if (emit_wire and self.pack is not None) else self._tail_split()
            self._split_active = split
            self._tail_done = False
            explicit(xb, yb, loss_acc=self.loss_acc, hooks=self if (split or self._first_gemm_hook is not None) else None)
            end = split if (split and self._tail_done) else a.n_param
            pack = self.pack if (emit_wire and end == a.n_param) else None
            F.fused_sgd(a.theta[:end], a.grad[:end], self.hyper,
                        a.momentum[:end] if a.momentum is not None else None,
                        bf[:end] if bf is not None else None, zero_grad=True, nesterov=self.nesterov, pack=pack)
            self.emitted_wire = pack is not None
            return
        out = self.model(xb)
        loss, stats = self._loss(out, yb)
        loss.backward()
        self.bnn.WGRAD.join()      # weight-gradient GEMMs run on a side stream; they must land before the step
        pack = self.pack if emit_wire else None
        F.fused_sgd(a.theta[: a.n_param], a.grad, self.hyper, a.momentum,
                    bf[: a.n_param] if bf is not None else None, zero_grad=True, nesterov=self.nesterov, pack=pack)
        self.emitted_wire = pack is not None
        self.loss_acc.add_(stats)

    # ---- optimizer / backward overlap (hooks called by ``model.explicit_step``) ----
    def _tail_split(self) -> int:
        """Arena offset where the deep layers' parameters start (0: no split)."""
        if not self.tail_overlap:
            return 0
        if self._split is None:
            prefix = getattr(self.model, "tail_split_prefix", None)
            self._split = 0
            if prefix:
                for name, slot in self.arena.slots.items():
                    if slot.is_param and name.startswith(prefix):
                        self._split = slot.offset - slot.offset % 8
                        break
        return self._split

    def tail_grads_ready(self):
        """Gradients of ``theta[split:n_param]`` are complete (once the weight-gradient branch has drained): run their
        SGD slice on its own stream, on a capped grid, beside the remaining backward pass."""
        a, dev, split = self.arena, self.device, self._split
        if self._tail_stream is None:
            self._tail_stream = torch.cuda.Stream(device=dev)
        side = self._tail_stream
        side.wait_stream(torch.cuda.current_stream(dev))
        wg = self.bnn.WGRAD.streams.get(dev)
        if wg is not None:
            side.wait_stream(wg)
        bf = a.theta_bf16
        with torch.cuda.stream(side):
            self.F.fused_sgd(a.theta[split: a.n_param], a.grad[split:], self.hyper,
                             a.momentum[split:] if a.momentum is not None else None,
                             bf[split: a.n_param] if bf is not None else None, zero_grad=True, nesterov=self.nesterov,
                             max_ctas=self.tail_ctas)
        self._tail_done = True
        self._tail_pending = True

    def after_first_gemm(self):
        """Called by ``model.explicit_step`` right after the GEMM of the model's first convolution."""
        if self._first_gemm_hook is not None:
            self._first_gemm_hook()

    def before_tail_forward(self):
        if self._tail_pending:
            torch.cuda.current_stream(self.device).wait_stream(self._tail_stream)
            self._tail_pending = False

    def _set_hyper(self, lr, momentum, weight_decay, dampening=0.0):
        vals = (float(lr), float(momentum), float(weight_decay), float(dampening))
        if vals != self._hyper_host:
            self.hyper.copy_(torch.tensor(vals, dtype=torch.float32))
            self._hyper_host = vals

    # -------------------------------------------------------------- epoch graph
    def _capture(self, X, y, n_steps, batch_size):
        perm = torch.zeros(n_steps * batch_size, dtype=torch.int64, device=self.device)
        perm.copy_(torch.arange(n_steps * batch_size, device=self.device) % X.shape[0])
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):   # warm-up outside capture (allocator, lazy init, autograd)
            # hyper lr=0 during warm-up/capture would still move BN statistics; save/restore the state
            snap = self.arena.theta.clone()
            snap_i = self.arena.int_arena.clone()
            snap_m = self.arena.momentum.clone() if self.arena.momentum is not None else None
            for _ in range(2):
                self._step(X, y, perm[:batch_size])
            self.before_tail_forward()       # drain the overlapped optimizer slice
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        from .ops._ext import total_launches
        graph = torch.cuda.CUDAGraph()
        graph2 = None
        c0 = total_launches()

        def body():
            # the epoch's batches are gathered by ONE launch pair (a permuted copy of the shard, 25 MB for the
            # flagship config) instead of two latency-bound gathers at the head of every step
            Xp, yp = self._gather(X, y, perm)
            for s in range(n_steps):
                self._step(X, y, None, batch=(Xp[s * batch_size:(s + 1) * batch_size],
                                              yp[s * batch_size:(s + 1) * batch_size]),
                           emit_wire=(s == n_steps - 1 and self.pack is not None))
            self.graph_emits_wire = self.pack is not None
            self.before_tail_forward()       # every forked stream must rejoin before the capture ends

        if (self.k3_join is not None and self.explicit and hasattr(self.model, "explicit_step")
                and getattr(self.model, "compute_dtype", "bf16") == "bf16"):
            # bcast_gemm (K3): the epoch is captured as TWO graphs that share one memory pool.  Graph 1 ends right after
            # the first convolution's GEMM of the first step -- everything in it either does not touch the parameter
            # arena (batch gather, im2col) or acquires the collective's arrival flags (weight staging, TMA producer of
            # the GEMM), so it is replayed WITHOUT waiting for the round-end collective that is still running on its
            # side stream.  Graph 2 (the rest of the epoch) is replayed after the join.
            graph2 = torch.cuda.CUDAGraph()
            pool = torch.cuda.graph_pool_handle()
            cap = torch.cuda.Stream(device=self.device)
            cap.wait_stream(torch.cuda.current_stream(self.device))
            state = {"cut": False}

            def cut():
                if not state["cut"]:
                    state["cut"] = True
                    graph.capture_end()
                    graph2.capture_begin(pool=pool)
            self._first_gemm_hook = cut
            with torch.cuda.stream(cap):
                graph.capture_begin(pool=pool)
                try:
                    body()
                finally:
                    self._first_gemm_hook = None
                (graph2 if state["cut"] else graph).capture_end()
            torch.cuda.current_stream(self.device).wait_stream(cap)
            if not state["cut"]:
                graph2 = None
        else:
            with torch.cuda.graph(graph):
                body()
        self.kernels_per_epoch = total_launches() - c0      # our kernels inside one epoch graph
        self.n_kernels_per_step = self.kernels_per_epoch // max(1, n_steps)
        # undo the side effects of warm-up + capture-time execution (capture does not execute,
        # warm-up did)
        self.arena.theta.copy_(snap)
        self.arena.int_arena.copy_(snap_i)
        if snap_m is not None:
            self.arena.momentum.copy_(snap_m)
        self.arena.grad.zero_()
        self.arena.sync_shadow()
        self.loss_acc.zero_()
        return {"graph": graph, "graph2": graph2, "perm": perm, "X": X, "y": y}

    # -------------------------------------------------------------- public
    @torch.no_grad()
    def _shuffle_into(self, perm, n, generator=None):
        perm.copy_(torch.randperm(n, device=self.device)[: perm.numel()])

    def run(self, X, y, n_epoch: int = 1, lr: float = 0.001, batch_size: int = 32, momentum: float = 0.0,
            weight_decay: float = 0.0, reshuffle_each_epoch: bool = False, return_device: bool = False,
            **_ignored):
        assert X.is_cuda, "GraphedLocalSGD needs a device-resident shard"
        nn.Module.train(self.model, True)
        n = X.shape[0]
        batch_size = min(batch_size, n)
        n_steps = n // batch_size
        tail = n - n_steps * batch_size
        self._set_hyper(lr, momentum, weight_decay)
        if momentum and self.arena.momentum is None:
            self.arena.momentum = torch.zeros_like(self.arena.grad)
        key = (n, batch_size, tuple(X.shape[1:]), tuple(y.shape[1:]), X.data_ptr(), y.data_ptr(), bool(momentum))
        epoch_losses = torch.zeros(n_epoch, 2, dtype=torch.float32, device=self.device)
        if self.use_graph:
            ent = self._graphs.get(key)
            if ent is None:
                ent = self._graphs[key] = self._capture(X, y, n_steps, batch_size)
            perm_full = torch.randperm(n, device=self.device)
            for e in range(n_epoch):
                if reshuffle_each_epoch and e > 0:
                    perm_full = torch.randperm(n, device=self.device)
                ent["perm"].copy_(perm_full[: n_steps * batch_size])
                self.loss_acc.zero_()
                ent["graph"].replay()
                if ent.get("graph2") is not None:
                    self.k3_join()           # the collective of the previous round must have landed from here on
                    ent["graph2"].replay()
                if tail:
                    with torch.enable_grad():
                        self._step(X, y, perm_full[n_steps * batch_size:])
                    self.before_tail_forward()
                epoch_losses[e].copy_(self.loss_acc)
        else:
            perm_full = torch.randperm(n, device=self.device)
            for e in range(n_epoch):
                if reshuffle_each_epoch and e > 0:
                    perm_full = torch.randperm(n, device=self.device)
                self.loss_acc.zero_()
                for idx in torch.split(perm_full, batch_size):
                    with torch.enable_grad():
                        self._step(X, y, idx)
                self.before_tail_forward()
                epoch_losses[e].copy_(self.loss_acc)
        steps = n_steps + (1 if tail else 0)
        self.last_steps = steps
        self.last_had_tail_step = bool(tail)        # a ragged eager step ran after the graph: its SGD did not emit the wire
        if self.use_graph:
            self.emitted_wire = bool(self.graph_emits_wire and not tail)
        if return_device:                     # caller reads (or forwards) the losses itself: no host sync here
            return epoch_losses
        host = epoch_losses.tolist()          # the ONLY host read of the round
        self.last_stats = {"accuracy": [h[1] / n for h in host], "steps_per_epoch": steps}
        return [h[0] / steps for h in host]


class PortableLocalSGD:
    """Same interface as :class:`GraphedLocalSGD` on plain PyTorch ops -- CPU / gloo runs of the SPMD engine
    (plumbing config, host-side logic tests).  Parameters stay views of the arena, so the session's reduce and
    broadcast work unchanged."""

    def __init__(self, model: nn.Module, arena, *, loss: str = "ce", **_unused):
        self.model, self.arena = model, arena
        self.loss_kind = loss
        self.device = arena.device
        self.last_steps = 1
        self.kernels_per_epoch = 0
        self.n_kernels_per_step = 0
        self.last_stats = {}

    def run(self, X, y, n_epoch: int = 1, lr: float = 0.001, batch_size: int = 32, momentum: float = 0.0,
            weight_decay: float = 0.0, reshuffle_each_epoch: bool = False, return_device: bool = False, **_ignored):
        criterion = _loss_fn(self.loss_kind)
        n = X.shape[0]
        batch_size = min(batch_size, n)
        nn.Module.train(self.model, True)
        opt = torch.optim.SGD(self.model.parameters(), lr=lr, momentum=momentum, weight_decay=weight_decay)
        perm = torch.randperm(n)
        out = torch.zeros(n_epoch, 2, dtype=torch.float32)
        steps = 1
        for e in range(n_epoch):
            if reshuffle_each_epoch and e > 0:
                perm = torch.randperm(n)
            batches = torch.split(perm, batch_size)
            steps = len(batches)
            for idx in batches:
                opt.zero_grad(set_to_none=True)
                pred = self.model(X[idx])
                tgt = y[idx]
                if pred.shape != tgt.shape and tgt.dtype.is_floating_point:
                    tgt = tgt.reshape(pred.shape)
                loss = criterion(pred.float() if tgt.dtype.is_floating_point else pred, tgt)
                loss.backward()
                opt.step()
                out[e, 0] += float(loss.detach())
                if not tgt.dtype.is_floating_point:
                    out[e, 1] += float((pred.argmax(-1) == tgt).sum())
        self.last_steps = steps
        if self.arena.theta_bf16 is not None:
            self.arena.sync_shadow()
        if return_device:
            return out
        host = out.tolist()
        self.last_stats = {"accuracy": [h[1] / n for h in host], "steps_per_epoch": steps}
        return [h[0] / steps for h in host]
