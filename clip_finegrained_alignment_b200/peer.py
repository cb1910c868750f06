"""Peer-memory exchange blocks for the all-gathered global InfoNCE (csrc/peer_exchange.cu).

One `PeerExchange` per (process group, B, D): the rank's own block is allocated by the library
(cudaMalloc + CUDA IPC handle), the 64-byte handles travel once through `torch.distributed`
(all_gather_object — plumbing, setup time only), every peer block is mapped with
cudaIpcOpenMemHandle, and from then on a gathered step issues no collective call at all: the
kernels read the peers' rows over NVLink and two in-stream device barriers order them.

All ranks of the group must sit on one NVLink/NVSwitch box (one process per GPU); otherwise the
mapping fails and `SPARCLoss(gather=True)` keeps the NCCL all-gather path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import torch

from . import _lib

_L = _lib.lib


class PeerExchange:
    def __init__(self, B: int, D: int, group=None):
        import torch.distributed as dist
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.B, self.D = B, D
        self.step = 0                                   # gathered forwards issued so far (same on every rank)
        self._status = None                             # pinned int32: barriers of this rank that timed out
        self._reported = 0
        self.nbytes = _L.cfa_peer_exchange_bytes(B, D)
        self._own = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        self._opened = []
        _lib.check(_L.cfa_peer_alloc(self.nbytes, C.byref(self._own), handle), "cfa_peer_alloc")
        handles = [None] * self.world
        dist.all_gather_object(handles, (bytes(handle), torch.cuda.current_device()), group=group)
        self.blocks = (C.c_void_p * self.world)()
        ok = 1
        try:
            for r, (h, _dev) in enumerate(handles):
                if r == self.rank:
                    self.blocks[r] = self._own.value
                    continue
                p = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                _lib.check(_L.cfa_peer_open(buf, C.byref(p)), f"cfa_peer_open(rank {r})")
                self._opened.append(p.value)
                self.blocks[r] = p.value
        except _lib.CfaError:
            ok = 0
        # every rank must take the same decision (peer path or NCCL path)
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        torch.cuda.synchronize()
        self.ok = bool(flag.item())
        if not self.ok:
            self.close()

    def next_step(self) -> int:
        s = self.step
        self.step = (s + 1) & 0x7FFFFFFF        # epochs 2s+1, 2s+2 run round the uint32 ring
        if (s & 63) == 63:
            self.check()                        # cheap: looks at the value an earlier asynchronous copy brought back
        return s

    def check(self, sync: bool = False) -> None:
        """Raise CfaError if a device barrier of this rank timed out (a peer never arrived: its step produced NaN losses,
        which GradScaler would silently turn into skipped steps).  Called every 64 gathered steps; `sync=True` waits for
        the stream first (end of an epoch, before a checkpoint)."""
        if self._status is None:
            self._status = torch.zeros(1, dtype=torch.int32).pin_memory()
        if sync:
            torch.cuda.current_stream().synchronize()
        n = int(self._status[0])                # value copied back by the PREVIOUS call (no host synchronisation)
        _lib.check(_L.cfa_peer_status(self._own.value, self._status.data_ptr(), _lib.stream_ptr()), "cfa_peer_status")
        if sync:
            torch.cuda.current_stream().synchronize()
            n = int(self._status[0])
        if n > self._reported:
            self._reported = n
            raise _lib.CfaError(f"gathered loss: {n} device barrier(s) of rank {self.rank} timed out waiting for a peer "
                                "(CFA_PEER_TIMEOUT_MS); the losses of those steps are NaN")

    def close(self):
        for p in self._opened:
            _L.cfa_peer_close(p)
        self._opened = []
        if self._own.value:
            _L.cfa_peer_free(self._own.value)
            self._own = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_EXCHANGES: Dict[Tuple, PeerExchange] = {}


def get_exchange(B: int, D: int, group=None):
    """Cached exchange for this (group, device, B, D); None when peer mapping is not possible."""
    key = (id(group) if group is not None else 0, torch.cuda.current_device(), B, D)
    ex = _EXCHANGES.get(key)
    if ex is None:
        ex = _EXCHANGES[key] = PeerExchange(B, D, group)
    return ex if ex.ok else None
