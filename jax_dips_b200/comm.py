"""Gradient all-reduce over NVLink peer memory, fused with the partial-row reduction (one process per
GPU on ONE node; CUDA IPC).  Stands in for `jax.lax.psum(grads/loss, "devices")` (trainer.py:829-830).

`PeerComm.reduce_allreduce(partials, rows, np1, out)` enqueues ONE single-CTA kernel that sums the
per-CTA partial rows of this rank, publishes the 168 floats in this rank's communication block,
waits for the peers' step flags and adds all blocks in rank order.  NCCL (`torch.distributed`) is
only used to exchange the 64-byte IPC handles at construction time.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _cabi as cabi


def _apply_timeout() -> float:
    """NBM_PEER_TIMEOUT_S: how long the exchange kernel waits for a peer before it reports an error (default 30 s;
    0 = wait without a bound, as NCCL does).  Returns the value in force."""
    t = float(os.environ.get("NBM_PEER_TIMEOUT_S", "30"))
    cabi.check(cabi.lib().nbm_comm_set_timeout(t), "nbm_comm_set_timeout")
    return t


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
        self.timeout_s = _apply_timeout()
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

    def reduce_allreduce_finalize(self, partials, rows, np1, out, fin) -> torch.Tensor:
        """exchange + optax chain + parameter staging as one kernel; fin = (optimizer struct, net struct, params,
        opt_state, opt_count, loss_hist or None)"""
        o, net, params, state, count, hist = fin
        cabi.check(cabi.lib().nbm_reduce_allreduce_finalize_f32(
            C.byref(o), C.byref(net), cabi.ptr(partials), rows, np1, self.rank, self.world, self.blocks,
            cabi.ptr(self.step), cabi.ptr(out), cabi.ptr(params), cabi.ptr(state), cabi.ptr(count), cabi.ptr(hist),
            cabi.stream_ptr()), "nbm_reduce_allreduce_finalize_f32")
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


class LocalPeerComm:
    """The same exchange inside ONE process that drives several devices (the reference's `pmap` model,
    trainer.py:727-743): one block per device, peer access enabled pairwise, no IPC.  `handle(r)` is device r's view
    (what `SharedPlan.loss_grad_launch(comm=...)` takes)."""

    class _Handle:
        def __init__(self, parent, rank):
            self.parent, self.rank = parent, rank

        def reduce_allreduce(self, partials, rows, np1, out):
            p = self.parent
            cabi.check(cabi.lib().nbm_reduce_allreduce_f32(cabi.ptr(partials), rows, np1, self.rank, p.world, p.blocks,
                                                           cabi.ptr(p.steps[self.rank]), cabi.ptr(out),
                                                           cabi.stream_ptr()), "nbm_reduce_allreduce_f32")
            return out

        def reduce_allreduce_finalize(self, partials, rows, np1, out, fin):
            p = self.parent
            o, net, params, state, count, hist = fin
            cabi.check(cabi.lib().nbm_reduce_allreduce_finalize_f32(
                C.byref(o), C.byref(net), cabi.ptr(partials), rows, np1, self.rank, p.world, p.blocks,
                cabi.ptr(p.steps[self.rank]), cabi.ptr(out), cabi.ptr(params), cabi.ptr(state), cabi.ptr(count),
                cabi.ptr(hist), cabi.stream_ptr()), "nbm_reduce_allreduce_finalize_f32")
            return out

    def __init__(self, devices):
        self.devices = [torch.device(d) for d in devices]
        self.world = len(self.devices)
        if not 1 <= self.world <= 8:
            raise cabi.NbmError("peer all-reduce supports up to 8 devices")
        L = cabi.lib()
        self.timeout_s = _apply_timeout()
        for a in self.devices:
            for b in self.devices:
                if a != b:
                    cabi.check(L.nbm_enable_peer_access(a.index, b.index), "nbm_enable_peer_access")
        self.blocks = (C.c_void_p * self.world)()
        self.steps = []
        for r, dev in enumerate(self.devices):
            with torch.cuda.device(dev):
                p = C.c_void_p()
                cabi.check(L.nbm_comm_alloc_local(C.byref(p)), "nbm_comm_alloc_local")
                self.blocks[r] = p.value
                self.steps.append(torch.zeros(1, dtype=torch.int32, device=dev))
                torch.cuda.synchronize(dev)
        self._handles = [self._Handle(self, r) for r in range(self.world)]

    def handle(self, r: int):
        return self._handles[r]

    def error(self) -> int:
        e = 0
        for r, dev in enumerate(self.devices):
            with torch.cuda.device(dev):
                e |= int(cabi.lib().nbm_comm_error(self.blocks[r]))
        return e

    def raise_on_error(self, where: str) -> None:
        if self.error():
            raise cabi.NbmError(f"peer all-reduce timed out ({where}): a device did not reach the exchange within "
                                f"{self.timeout_s:g} s (NBM_PEER_TIMEOUT_S)")

    def close(self):
        L = cabi.lib()
        for r, dev in enumerate(self.devices):
            with torch.cuda.device(dev):
                torch.cuda.synchronize(dev)
                if self.blocks[r]:
                    L.nbm_comm_free(self.blocks[r])
                    self.blocks[r] = None
