"""The fused NVLink FedAvg data plane (host side).

``FedAvgSession`` owns the symmetric wire buffer of one rank and launches the
single-kernel round-end collective (``csrc/fedavg.cu``): pack/cast -> weighted
reduce over peer memory -> broadcast -> running-mean apply into the fp32 master,
the bf16 shadow and (optionally) per-tile arrival flags that gate the first GEMM
of the next forward (``bcast_gemm``).

It replaces, for GPU-seated clients, the reference's upload (worker.py:108-118),
the manager's CPU reduce (manager.py:119-126), the broadcast (manager.py:77-86)
and ``load_state_dict`` (worker.py:98).  ``NcclSession`` implements the same
interface with ``torch.distributed`` collectives: it is the BASELINE the fused
kernel is measured against and the oracle the tests compare with -- not the
product path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .arena import ParamArena
from .symm import SymmetricBuffer

MAX_LOSS = 64       # per-epoch loss slots carried through the collective
MAX_CTAS = 296      # pad slots; the kernel runs one 512-thread CTA per SM (148 on B200), clamped by the launcher


def _align(x: int, a: int) -> int:
    return (x + a - 1) // a * a


class FedAvgSession:
    def __init__(self, arena: ParamArena, group=None, *, wire_dtype: str = "bf16", mode: str = "delta",
                 nvls: "bool | str" = "auto", n_ctas: int = 148, tile_elems: int = 0, timeout_log2: int = 24,
                 reset_momentum: bool = True, tile_flags: bool = False):
        from ..ops._ext import load
        self._C = load()
        assert wire_dtype in ("bf16", "fp32", "fp8") and mode in ("delta", "weights")
        self.arena = arena
        self.device = arena.device
        self.group = group
        self.wire_dtype = wire_dtype
        self.wire_bf16 = wire_dtype == "bf16"
        self.wire_kind = {"fp32": 0, "bf16": 1, "fp8": 2}[wire_dtype]
        self.delta = mode == "delta"
        self.n_ctas = max(1, min(int(n_ctas), MAX_CTAS))
        self.tile_elems = int(tile_elems)
        # every cross-GPU spin is BOUNDED by default (2^24 polls, several seconds): a seat that dies between the
        # manager's plan and its launch turns into an error status on the survivors (check()), not into eight GPUs
        # spinning forever -- the NCCL failure mode this data plane exists to avoid.  0 = spin without limit.
        self.timeout_log2 = int(timeout_log2)
        self.reset_momentum = reset_momentum
        # wire / int / loss pages exist TWICE (round parity): the kernel has no closing barrier, a rank that races ahead
        # packs the next round into the other half while a slow peer still applies this one (csrc/fedavg.cu)
        self.half_wire = _align(self.wire_bytes(), 2 << 20)          # multicast-friendly granularity
        self.half_int = _align(max(arena.n_int, 1) * 8, 256)
        self.half_loss = _align(MAX_LOSS * 4, 256)
        self.off_wire = 0
        self.off_int = 2 * self.half_wire
        self.off_loss = self.off_int + 2 * self.half_int
        self.off_pads = _align(self.off_loss + 2 * self.half_loss, 256)
        total = _align(self.off_pads + (MAX_CTAS + 8) * self._C.MAX_RANKS * 8, 2 << 20)
        self.symm = SymmetricBuffer(total, self.device, group)
        self.rank, self.world = self.symm.rank, self.symm.world
        assert self.world <= self._C.MAX_RANKS
        self.use_nvls = self.symm.has_multicast if nvls == "auto" else (bool(nvls) and self.symm.has_multicast)
        if self.wire_kind == 2:
            self.use_nvls = False      # the switch adds raw elements; block scales need the P2P path
        self.epoch = 0
        self.rounds = 0
        self.stale = False          # True after a round this seat sat out: its weights are no longer the global model
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.loss_local = torch.zeros(MAX_LOSS, dtype=torch.float32, device=self.device)
        self.loss_out = torch.zeros(MAX_LOSS, dtype=torch.float32, device=self.device)
        # tile_elems == 0: sized per launch so that every CTA of every live rank owns ~one tile
        self.min_tile = 1024
        n_tiles = (arena.n + self.min_tile - 1) // self.min_tile
        self.tile_flags = torch.zeros(n_tiles, dtype=torch.int32, device=self.device) if tile_flags else None
        # flag value consumers must wait for (= rounds launched so far): device-resident so captured graphs follow
        self.epoch_word = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.last_tile_elems = self.tile_elems or self.min_tile
        # a high-priority stream lets the collective's CTAs become resident ahead of a flag-gated
        # GEMM that is launched right behind it on the compute stream
        self.stream = torch.cuda.Stream(device=self.device, priority=-1) if self.device.type == "cuda" else None
        self.symm.barrier()
        # K4: the last SGD step of an epoch may write this rank's wire copy itself (ops.fused_sgd(pack=...)); the wire
        # address of the upcoming round (parity half) and the pack scale live in device words so a captured graph follows
        self.wire_slot = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.pack_scale = torch.ones(1, dtype=torch.float32, device=self.device)
        self._armed_for = None     # (epoch, scale) the wire words were armed for
        self.phase_ns = None       # enable_phase_timing(): int64[16] of %globaltimer stamps written by the kernel
        self.nvls_choice = "forced" if nvls != "auto" else "default"
        if nvls == "auto" and self.use_nvls and self.delta and self.world > 1 and arena.global_w is not None:
            self.autotune_nvls()
        # barrier epoch at the end of construction (the autotune above already ran collectives): identical on every rank,
        # and the origin of the manager-dictated round index (aggregate(round_index=...))
        self.base_epoch = self.epoch

    FLAG_GRANULE = 1024        # elements per arrival flag (csrc/fedavg.cu)

    PHASES = ("pack", "barrier1", "reduce_bcast", "barrier2", "apply", "exit")

    def enable_phase_timing(self) -> None:
        """Ask the kernel to record %globaltimer at its phase boundaries (first and last CTA) -- the way to
        see where a multi-GPU round goes, since kernels with cross-GPU spin barriers cannot run under ncu.
        Needs an extension built with ``BATON_BUILD_PHASE_TIMING=1 python -m baton_b200.build_ext`` (the default
        build compiles the stamps out, so all durations read 0)."""
        self.phase_ns = torch.zeros(16, dtype=torch.int64, device=self.device)

    def phase_breakdown_us(self) -> dict:
        """Phase durations of the LAST launch in microseconds: ``{"first_cta": {...}, "last_cta": {...}}``."""
        if self.phase_ns is None:
            return {}
        t = self.phase_ns.tolist()
        out = {}
        for name, base in (("first_cta", 0), ("last_cta", 8)):
            out[name] = {p: (t[base + i + 1] - t[base + i]) / 1e3 for i, p in enumerate(self.PHASES)}
        return out

    def autotune_nvls(self, iters: int = 3) -> None:
        """Measure, don't guess: time the collective both ways (peer loads/stores vs in-switch
        ``multimem`` reduce + multicast store) on THIS box and world size and keep the faster one.  Runs on
        zero deltas (``theta == global`` right after the arena is built), so it leaves the model untouched;
        every rank takes the max over ranks, so all ranks agree."""
        import torch.distributed as dist
        if not torch.equal(self.arena.theta[: self.arena.n_param], self.arena.global_w[: self.arena.n_param]):
            return                                   # replicas already drifted: keep the default
        best = {}
        for mode in (False, True):
            self.use_nvls = mode
            ts = []
            for it in range(iters + 1):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                self.aggregate(my_n=1.0)
                e1.record()
                torch.cuda.synchronize(self.device)
                if it:
                    ts.append(e0.elapsed_time(e1))
            best[mode] = min(ts)
        t = torch.tensor([best[False], best[True]], device=self.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        self.use_nvls = bool(t[1] < t[0])
        self.nvls_choice = "autotuned p2p {:.0f} us vs nvls {:.0f} us".format(float(t[0]) * 1e3, float(t[1]) * 1e3)
        self.check()
        self.symm.barrier()

    # ------------------------------------------------------------------ K4: upload copy emitted by the optimizer
    def pack_spec(self) -> Optional[dict]:
        """Arguments for ``ops.fused_sgd(pack=...)`` (stable device tensors: safe to capture), or None when the wire
        format needs the in-kernel pack (block-scaled fp8)."""
        if self.wire_kind == 2 or self.device.type != "cuda":
            return None
        a = self.arena
        return {"wire_slot": self.wire_slot, "global_w": a.global_w if self.delta else None, "scale": self.pack_scale,
                "n_pack": a.n, "wire_fp32": self.wire_kind == 0}

    def arm_prepack(self, my_n: float) -> None:
        """Point the device words at the wire half of the UPCOMING round and set the pack scale (n_k under NVLS, where
        the switch can only add; 1 for peer loads, which weight on the reader side).  Call before the local epoch(s)."""
        par = (self.epoch // 3) & 1
        addr = self.symm.peer_ptrs(self.off_wire + par * self.half_wire)[self.rank]
        scale = float(my_n) if (self.use_nvls and self.world > 1) else 1.0
        key = (self.epoch, scale)
        if self._armed_for != key:
            self.wire_slot.fill_(int(addr))
            self.pack_scale.fill_(scale)
            self._armed_for = key

    # ------------------------------------------------------------------ the collective
    def aggregate(self, n_samples_by_rank: Optional[Sequence[float]] = None,
                  alive_ranks: Optional[Sequence[int]] = None, my_n: Optional[float] = None,
                  loss_history: Optional[Sequence[float]] = None, on_side_stream: bool = False,
                  round_index: Optional[int] = None, prepacked: bool = False) -> None:
        """Launch the fused reduce+broadcast+apply.  Either the full per-rank sample
        counts are given (manager-driven rounds: the plan comes over HTTP) or only this
        rank's own count ``my_n`` (SPMD engine: peers' counts ride on the barrier flags).

        ``round_index``: number of aggregations the MANAGER has dispatched before this one.  The manager is the single
        authority for it: the barrier epoch becomes ``base_epoch + 3 * round_index`` on every seat, so a seat that sat
        out rounds (evicted, re-registered) re-enters in step with its peers instead of racing them with a lagging
        counter (epochs must never run backwards: the pads keep the highest epoch ever seen)."""
        world = self.world
        if round_index is not None:
            self.epoch = (self.base_epoch + 3 * int(round_index)) & 0xFFFFFFFF
        if n_samples_by_rank is not None:
            counts = [float(x) for x in n_samples_by_rank] + [0.0] * (world - len(n_samples_by_rank))
            counts = counts[:world]
            from_flags = False
        else:
            assert my_n is not None
            counts = [0.0] * world
            counts[self.rank] = float(my_n)
            from_flags = True
        alive = list(range(world)) if alive_ranks is None else [int(r) for r in alive_ranks if 0 <= int(r) < world]
        if self.rank not in alive:
            # not part of this round: the replica goes stale (``stale`` tells the worker to pull the global model
            # before it takes part again) but the barrier epoch and the round counter keep pace with the peers
            self.epoch = (self.epoch + 3) & 0xFFFFFFFF
            self.rounds += 1
            self.stale = True
            return
        mask = 0
        for r in alive:
            mask |= 1 << r
        if loss_history is not None:
            k = min(len(loss_history), MAX_LOSS)
            self.loss_local.zero_()
            self.loss_local[:k].copy_(torch.tensor([float(x) for x in loss_history[:k]]), non_blocking=True)
        a = self.arena
        # the optimizer's wire copy is only valid for the round / scale it was armed for and when the collective runs
        # the way arm_prepack assumed (NVLS needs every rank alive); otherwise the kernel packs itself (always correct)
        nvls_now = bool(self.use_nvls and len(alive) == world)
        want_scale = (counts[self.rank] if not from_flags else float(my_n)) if (self.use_nvls and world > 1) else 1.0
        prepacked = bool(prepacked and self.wire_kind != 2 and self._armed_for == (self.epoch, float(want_scale))
                         and (nvls_now == bool(self.use_nvls and world > 1)))
        self.last_prepacked = prepacked
        if self.tile_elems:
            tile = self.tile_elems
        else:   # one tile per (live rank, CTA): n / (A * G), rounded up to a multiple of 8 elements
            per = -(-a.n // (len(alive) * self.n_ctas))
            tile = max(self.min_tile, (per + 31) // 32 * 32)
        if self.tile_flags is not None:      # arrival flags cover fixed 1024-element granules: tiles must not split one
            tile = (tile + self.FLAG_GRANULE - 1) // self.FLAG_GRANULE * self.FLAG_GRANULE
        self.last_tile_elems = tile
        flag_value = self.rounds + 1
        par = (self.epoch // 3) & 1            # round parity: which half of the wire / int / loss pages this round uses
        o_wire, o_int, o_loss = (self.off_wire + par * self.half_wire, self.off_int + par * self.half_int,
                                 self.off_loss + par * self.half_loss)
        cur = torch.cuda.current_stream(self.device)
        stream = self.stream if on_side_stream else cur
        if self.tile_flags is not None:
            self.epoch_word.fill_(flag_value)       # compute stream: whatever is enqueued after this call waits for THIS round
        if on_side_stream:
            stream.wait_stream(cur)
        with torch.cuda.stream(stream):
            self._C.fedavg_allreduce(
                self.symm.peer_ptrs(o_wire), self.symm.peer_ptrs(self.off_pads),
                self.symm.mc(o_wire) if self.use_nvls else 0,
                a.theta, a.global_w, a.theta_bf16,
                a.momentum if self.reset_momentum else None,
                a.int_arena if a.n_int > 0 else None,
                self.symm.peer_ptrs(o_int) if a.n_int > 0 else [],
                self.loss_local, self.symm.peer_ptrs(o_loss), self.loss_out,
                counts, from_flags, mask, self.rank, world, self.wire_kind, self.delta,
                bool(self.use_nvls and len(alive) == world), self.epoch,
                self.tile_flags, flag_value, tile, self.n_ctas, self.timeout_log2, self.status, self.phase_ns,
                prepacked)
        self.epoch = (self.epoch + 3) & 0xFFFFFFFF     # uint32 wrap: the kernel compares signed differences
        self.rounds += 1
        self._side_pending = on_side_stream

    def join(self) -> None:
        """Make the compute stream wait for a side-stream collective."""
        if getattr(self, "_side_pending", False):
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
            self._side_pending = False

    def gate_first_layer(self, module, slot_name: Optional[str] = None) -> None:
        """bcast_gemm: make the NEXT forward of ``module`` (an ``ops.nn.Linear`` whose
        weight is the first GEMM operand of the model) wait, tile by tile, on the arrival
        flags this round's collective publishes, instead of on the whole kernel."""
        if self.tile_flags is None:
            raise RuntimeError("session was created with tile_flags=False")
        slot = None
        for name, s in self.arena.slots.items():
            if slot_name is not None and name == slot_name:
                slot = s
                break
            if slot_name is None and s.is_param and len(s.shape) == 2:
                owner = self.arena._owner(name)
                if owner is module:
                    slot = s
                    break
        if slot is None:
            raise ValueError("module weight not found in the arena")
        bias_off = -1
        if getattr(module, "bias", None) is not None:
            for name, s in self.arena.slots.items():
                if s.is_param and self.arena._owner(name) is module and name.endswith(".bias"):
                    bias_off = s.offset
        module.flags_cfg = {"flags": self.tile_flags, "epoch": self.rounds, "elem_off": slot.offset,
                            "tile_elems": self.FLAG_GRANULE, "bias_off": bias_off}

    def gate_first_conv(self, conv) -> None:
        """bcast_gemm for a convolutional first layer, usable INSIDE a captured CUDA graph: the staging kernel of the
        layer's weights and the TMA producer of its GEMM acquire the arrival flags of the arena granules under
        ``conv.weight`` and compare them with ``self.epoch_word`` -- a device word the session bumps (on the compute
        stream) every time it launches a collective.  Steps after the first of a round find the flags already there
        (one cached load), so every replay of the captured step can stay gated."""
        if self.tile_flags is None:
            raise RuntimeError("session was created with tile_flags=False")
        slot = None
        for name, s in self.arena.slots.items():
            if s.is_param and self.arena._owner(name) is conv and name.endswith(".weight"):
                slot = s
                break
        if slot is None:
            raise ValueError("conv weight not found in the arena")
        conv.flags_cfg = {"flags": self.tile_flags, "epoch_word": self.epoch_word, "elem_off": slot.offset,
                          "tile_elems": self.FLAG_GRANULE}

    def reduced_loss(self, n_epoch: int) -> List[float]:
        return self.loss_out[: min(n_epoch, MAX_LOSS)].tolist()

    def check(self) -> None:
        code = int(self.status.item())
        if code:
            self.status.zero_()
            raise RuntimeError("FedAvg collective timed out waiting for rank {}".format(code - 1))

    def wire_bytes(self) -> int:
        """Bytes one client uploads per round (fp8: e4m3 payload + one UE8M0 scale byte per 32 elements)."""
        n = self.arena.n
        if self.wire_kind == 2:
            return n + (n + 31) // 32
        return n * (2 if self.wire_kind == 1 else 4)


class NcclSession:
    """Same contract through ``torch.distributed`` collectives (baseline / oracle)."""

    def __init__(self, arena: ParamArena, group=None, *, wire_dtype: str = "bf16", mode: str = "delta",
                 reset_momentum: bool = True, **_unused):
        import torch.distributed as dist
        self.dist = dist
        self.arena, self.group, self.device = arena, group, arena.device
        self.wire_dtype = torch.bfloat16 if wire_dtype == "bf16" else torch.float32
        self.delta = mode == "delta"
        self.reset_momentum = reset_momentum
        inited = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if inited else 1
        self.rank = dist.get_rank(group) if inited else 0
        self.wire = torch.zeros(arena.n, dtype=self.wire_dtype, device=self.device)
        self.counts = torch.zeros(self.world, dtype=torch.float32, device=self.device)
        self.loss_buf = torch.zeros(MAX_LOSS, dtype=torch.float32, device=self.device)
        self.loss_out = torch.zeros(MAX_LOSS, dtype=torch.float32, device=self.device)
        self.rounds = 0

    @torch.no_grad()
    def aggregate(self, n_samples_by_rank=None, alive_ranks=None, my_n=None, loss_history=None, **_unused) -> None:
        a, dist = self.arena, self.dist
        if n_samples_by_rank is not None:
            counts = torch.tensor([float(x) for x in n_samples_by_rank][: self.world], device=self.device)
        else:
            self.counts.zero_()
            self.counts[self.rank] = float(my_n)
            if self.world > 1:
                dist.all_reduce(self.counts, group=self.group)
            counts = self.counts
        total = counts.sum()
        w = counts[self.rank] / total
        src = (a.theta - a.global_w) if self.delta else a.theta
        self.wire.copy_((src * w).to(self.wire_dtype))
        if self.world > 1:
            dist.all_reduce(self.wire, group=self.group)
        # every rank joins the loss reduce every round -- a rank that hosts no sampled client this round
        # contributes zeros (weight 0); skipping the call there would desynchronise the collectives
        self.loss_buf.zero_()
        if loss_history is not None:
            k = min(len(loss_history), MAX_LOSS)
            self.loss_buf[:k] = torch.tensor([float(x) for x in loss_history[:k]], device=self.device) * w
        if self.world > 1:
            dist.all_reduce(self.loss_buf, group=self.group)
        self.loss_out.copy_(self.loss_buf)
        if self.delta:
            a.global_w.add_(self.wire.float())
        else:
            a.global_w.copy_(self.wire.float())
        a.theta.copy_(a.global_w)
        if a.theta_bf16 is not None:
            a.theta_bf16.copy_(a.theta.to(torch.bfloat16))
        if a.n_int > 0 and self.world > 1:
            dist.all_reduce(a.int_arena, op=dist.ReduceOp.MAX, group=self.group)
        if self.reset_momentum and a.momentum is not None:
            a.momentum.zero_()
        self.rounds += 1

    def join(self) -> None:
        pass

    def reduced_loss(self, n_epoch: int) -> List[float]:
        return self.loss_out[: min(n_epoch, MAX_LOSS)].tolist()

    def check(self) -> None:
        pass

    def wire_bytes(self) -> int:
        return self.arena.n * (2 if self.wire_dtype == torch.bfloat16 else 4)
