"""Peer exchange: candidates of the row-sharded search pushed to every rank over NVLink and merged on arrival.

Replaces the gathers of reference src/index.py:135-157 — and the NCCL all-gather in front of ``mips_merge_topk`` —
with two launches per search (``mips_xchg_merge``: push, wait + merge).  One node, one process per GPU; set-up is
collective (64-byte CUDA IPC handles travel through the existing process group).
"""
from __future__ import annotations

import ctypes
import os
import socket
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _native as N


class PeerExchangeTimeout(RuntimeError):
    """A peer's block did not arrive within the exchange timeout (JSA_MIPS_XCHG_TIMEOUT_S, default 30 minutes — the
    order of a collective watchdog; the reference tolerates 100000 s, src/slurm.py:181).  Nothing trapped: the CUDA
    context is intact, the step that timed out returned padding, and the exchange has to be rebuilt."""


class PeerExchange:
    """Collective object: every rank constructs it with the same capacity, calls ``merge`` the same number of times
    with the same (batch, k) and closes it together.  A rank may be arbitrarily late (checkpointing, logging, a GC
    pause): its peers' wait kernels simply keep polling until the timeout."""

    def __init__(self, device: torch.device, block_capacity: int):
        self._lib = N.load()
        self.device = torch.device(device)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self._h = ctypes.c_void_p()
        rc = self._lib.mips_xchg_create(ctypes.byref(self._h), self.device.index or 0, self.rank, self.world,
                                        int(block_capacity))
        handle = ctypes.create_string_buffer(self._lib.mips_xchg_handle_bytes())
        if rc == N.MIPS_OK:
            rc = self._lib.mips_xchg_export(self._h, handle)
        # the handles ride the process group as bytes; a rank that failed sends an empty one
        everyone = [None] * self.world
        dist.all_gather_object(everyone, (socket.gethostname(), handle.raw if rc == N.MIPS_OK else b""))
        ok = rc == N.MIPS_OK and all(h for _, h in everyone) and len({host for host, _ in everyone}) == 1
        if ok:
            rc = self._lib.mips_xchg_connect(self._h, b"".join(h for _, h in everyone))
            ok = rc == N.MIPS_OK
        votes = [None] * self.world
        dist.all_gather_object(votes, bool(ok))
        self.ok = all(votes)
        self.why = "" if self.ok else (self._lib.mips_xchg_last_error(self._h) or b"").decode("utf-8", "replace") \
            if self._h else "create failed"
        self.capacity = int(self._lib.mips_xchg_capacity(self._h)) if self._h else 0
        if not self.ok:
            self._free()

    def merge(self, local_block: torch.Tensor, batch: int, k_in: int, k_out: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """``local_block``: this rank's ``packed_result_buffer`` block (uint8 [block]).  Returns the merged
        (scores fp32 [batch, k_out], global ids int64 [batch, k_out]) of all ranks' blocks."""
        if not self.ok:
            raise RuntimeError("peer exchange is not connected")
        dev = local_block.device
        s_bytes = (batch * k_in * 4 + 7) // 8 * 8
        out_s = torch.empty((batch, k_out), dtype=torch.float32, device=dev)
        out_i = torch.empty((batch, k_out), dtype=torch.int64, device=dev)
        rc = self._lib.mips_xchg_merge(self._h, ctypes.c_void_p(local_block.data_ptr()), local_block.numel(), s_bytes, batch,
                                       k_in, k_out, ctypes.c_void_p(out_s.data_ptr()), ctypes.c_void_p(out_i.data_ptr()),
                                       ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        self._check(rc, "mips_xchg_merge")
        return out_s, out_i

    def _check(self, rc: int, what: str) -> None:
        if rc == N.MIPS_OK:
            return
        msg = f"{what}: " + (self._lib.mips_xchg_last_error(self._h) or b"").decode("utf-8", "replace")
        if rc == N.MIPS_ETIMEOUT:
            raise PeerExchangeTimeout(msg)
        raise RuntimeError(msg)

    def status(self) -> None:
        """Raises PeerExchangeTimeout if a wait kernel of this exchange has given up (host-side read, no launch)."""
        if self._h:
            self._check(self._lib.mips_xchg_status(self._h), "mips_xchg_status")

    def set_timeout(self, seconds: float) -> None:
        self._check(self._lib.mips_xchg_set_timeout_ms(self._h, max(1, int(seconds * 1000))), "mips_xchg_set_timeout_ms")

    def gather(self, local: torch.Tensor) -> torch.Tensor:
        """All-gather of equally shaped contiguous tensors: returns [W, *local.shape] (rank order)."""
        if not self.ok:
            raise RuntimeError("peer exchange is not connected")
        local = local.contiguous()
        out = torch.empty((self.world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
        rc = self._lib.mips_xchg_gather(self._h, ctypes.c_void_p(local.data_ptr()), local.numel() * local.element_size(),
                                        ctypes.c_void_p(out.data_ptr()),
                                        ctypes.c_void_p(torch.cuda.current_stream(local.device).cuda_stream))
        self._check(rc, "mips_xchg_gather")
        return out

    def _free(self):
        if self._h:
            self._lib.mips_xchg_destroy(self._h)
            self._h = ctypes.c_void_p()

    def close(self):
        """Collective: nobody frees its buffer while a peer may still be storing into it."""
        if self._h:
            torch.cuda.synchronize(self.device)
            if dist.is_initialized():
                dist.barrier()
            self._free()
        self.ok = False

    def __del__(self):   # best effort at interpreter exit; close() is the orderly way
        try:
            self._free()
        except Exception:
            pass


def exchange_mode() -> str:
    """JSA_MIPS_EXCHANGE = p2p (default: peer stores, NCCL when the GPUs cannot map each other) | nccl."""
    return os.environ.get("JSA_MIPS_EXCHANGE", "p2p").lower()


def make_peer_exchange(device, block_capacity: int) -> Optional[PeerExchange]:
    """Collective.  Returns a connected exchange or None (caller keeps using the all-gather path)."""
    if exchange_mode() != "p2p" or not dist.is_initialized() or dist.get_backend() != "nccl" or dist.get_world_size() > 16:
        return None
    x = PeerExchange(device, block_capacity)
    if not x.ok:
        if dist.get_rank() == 0:
            import warnings
            warnings.warn(f"peer exchange unavailable ({x.why}); using the NCCL all-gather + merge path")
        return None
    return x
