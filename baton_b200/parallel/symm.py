"""Symmetric (peer-mapped) device memory for the NVLink data plane.

One process per GPU; every rank allocates a buffer of identical size, the
handles are exchanged once through the process group's store (rendezvous), and
afterwards each rank holds

  * ``ptrs[k]``      -- a device pointer valid on THIS GPU that aliases rank k's
                        buffer (NVLink peer mapping),
  * ``mc_ptr``       -- a multicast address bound to every replica (NVLS): one
                        ``multimem.st`` is replicated by the NVSwitch, one
                        ``multimem.ld_reduce`` is summed in the switch; 0 when the
                        fabric has no multicast support.

The rendezvous uses ``torch.distributed._symmetric_memory`` (CUDA VMM
allocations + fabric/fd handle exchange); NCCL itself is only the bootstrap
transport, the hot path issues peer loads/stores from inside our own kernels
(SURVEY.md section 5.1).  With ``world_size == 1`` (single-GPU runs, CPU tests)
the buffer is a plain local tensor whose only "peer" is itself.
"""
from __future__ import annotations

from typing import List, Optional

import torch


class SymmetricBuffer:
    def __init__(self, nbytes: int, device, group=None, zero: bool = True):
        import torch.distributed as dist
        self.nbytes = int((nbytes + 15) // 16 * 16)
        self.device = torch.device(device)
        self.group = group
        self.handle = None
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.world = world
        self.rank = dist.get_rank(group) if world > 1 else 0
        if world > 1 and self.device.type == "cuda":
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else dist.group.WORLD
            self.local = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=self.device)
            if zero:
                self.local.zero_()
            self.handle = symm_mem.rendezvous(self.local, grp)
            self.ptrs: List[int] = [int(p) for p in self.handle.buffer_ptrs]
            mc = 0
            try:
                mc = int(self.handle.multicast_ptr or 0)
            except Exception:
                mc = 0
            self.mc_ptr = mc
        else:
            self.local = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.device)
            self.ptrs = [self.local.data_ptr()] * max(world, 1) if world == 1 else [self.local.data_ptr()]
            self.mc_ptr = 0

    @property
    def has_multicast(self) -> bool:
        return self.mc_ptr != 0

    def view(self, byte_offset: int, numel: int, dtype: torch.dtype) -> torch.Tensor:
        esz = torch.empty((), dtype=dtype).element_size()
        return self.local[byte_offset: byte_offset + numel * esz].view(dtype)

    def peer_ptrs(self, byte_offset: int = 0) -> List[int]:
        return [p + byte_offset for p in self.ptrs]

    def mc(self, byte_offset: int = 0) -> int:
        return self.mc_ptr + byte_offset if self.mc_ptr else 0

    def measure_link_gbps(self, iters: int = 8, max_bytes: int = 256 << 20) -> Optional[float]:
        """Measured NVLink bandwidth of THIS box, per direction: every rank pulls its right neighbour's buffer with
        the copy engine at the same time (peer view -> local scratch), timed with CUDA events on the device.  This
        is the denominator of the collective's roofline (bench.py), measured in-repo instead of quoted.  ``None``
        without peers."""
        if self.handle is None or self.world < 2:
            return None
        n = min(self.nbytes, max_bytes) // 16 * 16
        peer = (self.rank + 1) % self.world
        src = self.handle.get_buffer(peer, (n,), torch.uint8)
        dst = torch.empty(n, dtype=torch.uint8, device=self.device)
        dst.copy_(src)                                   # maps + warms the path
        torch.cuda.synchronize(self.device)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            dst.copy_(src)
        e1.record()
        torch.cuda.synchronize(self.device)
        self.barrier()
        return n * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9

    def barrier(self) -> None:
        """Host-visible rendezvous barrier (setup/teardown only, never on the hot path)."""
        if self.handle is not None:
            self.handle.barrier()
        elif self.world > 1:
            import torch.distributed as dist
            dist.barrier(self.group)
