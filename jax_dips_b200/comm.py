"""Gradient all-reduce over NVLink peer memory, fused with the partial-row reduction (one process per
GPU on ONE node; CUDA IPC).  Stands in for `jax.lax.psum(grads/loss, "devices")` (trainer.py:829-830).

`PeerComm.reduce_allreduce(partials, rows, np1, out)` enqueues ONE single-CTA kernel that sums the
per-CTA partial rows of this rank, publishes the 168 floats in this rank's communication block,
waits for the peers' step flags and adds all blocks in rank order.  NCCL (`torch.distributed`) is
only used to exchange the 64-byte IPC handles at construction time.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi as cabi


class PeerComm:
    def __init__(self, device, group=None):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise cabi.NbmError("PeerComm needs an initialised torch.distributed process group")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise cabi.NbmError("peer all-reduce supports up to 8 ranks on one node")
        self.device = torch.device(device)
        L = cabi.lib()
        with torch.cuda.device(self.device):
            local = C.c_void_p()
            handle = C.create_string_buffer(64)
            cabi.check(L.nbm_comm_alloc(C.byref(local), handle), "nbm_comm_alloc")
            self.local = local
            handles = [None] * self.world
            dist.all_gather_object(handles, handle.raw, group=group)
            self.blocks = (C.c_void_p * self.world)()
            self._peers = []
            for r in range(self.world):
                if r == self.rank:
                    self.blocks[r] = local.value
                else:
                    p = C.c_void_p()
                    cabi.check(L.nbm_comm_open_peer(handles[r], C.byref(p)), "nbm_comm_open_peer")
                    self.blocks[r] = p.value
                    self._peers.append(p)
            self.step = torch.zeros(1, dtype=torch.int32, device=self.device)
            torch.cuda.synchronize()
        dist.barrier(group=group)   # every rank has mapped every block before the first kernel spins on one

    def reduce_allreduce(self, partials: torch.Tensor, rows: int, np1: int, out: torch.Tensor) -> torch.Tensor:
        cabi.check(cabi.lib().nbm_reduce_allreduce_f32(cabi.ptr(partials), rows, np1, self.rank, self.world,
                                                       self.blocks, cabi.ptr(self.step), cabi.ptr(out),
                                                       cabi.stream_ptr()), "nbm_reduce_allreduce_f32")
        return out

    def error(self) -> int:
        return int(cabi.lib().nbm_comm_error(self.local))

    def close(self):
        L = cabi.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for p in self._peers:
                L.nbm_comm_close_peer(p)
            self._peers = []
            if self.local is not None:
                L.nbm_comm_free(self.local)
                self.local = None
