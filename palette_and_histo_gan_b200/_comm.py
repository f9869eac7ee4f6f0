"""The one exchange of the sharded path over peer memory (NVLink / NVSwitch), `ph_comm_*` of the C ABI.

The Hellinger distance takes one square root over the whole batch (histogram.py:88-89); ranks holding shards of
the batch add up one float64 per step.  `PeerComm.allreduce_` does that with ONE 32-thread kernel: stores into
every peer's mailbox, polls of the local mailbox, sum in rank order (bit-identical on all ranks) — ~3 us instead
of the ~30 us of an 8-byte NCCL all-reduce and no framework kernels around it.  `torch.distributed` is only used
once, to hand the 64-byte CUDA IPC handles around.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._tensor import ptr, stream_ptr

_cache: dict = {}


class PeerComm:
    """Mailboxes of the ranks of `group` (None = default group) mapped into each other's address space."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist

        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        self._h = C.c_void_p()
        # Every rank takes part in the exchange of the handles whatever happened before it: a rank whose mailbox could
        # not be created sends an empty handle instead of leaving its peers waiting in the collective.
        failure = None
        mine = b""
        try:
            _lib.call("ph_comm_create", dev.index, self.rank, self.world, C.byref(self._h))
            buf = C.create_string_buffer(_lib.COMM_HANDLE_BYTES)
            _lib.call("ph_comm_export", self._h, buf)
            mine = bytes(buf.raw)
        except _lib.PalHistError as e:
            failure = e
        handles = [None] * self.world
        with torch.cuda.device(dev):
            dist.all_gather_object(handles, mine, group=group)
        if failure is None and any(len(h) != _lib.COMM_HANDLE_BYTES for h in handles):
            failure = RuntimeError("a peer rank could not create its mailbox")
        if failure is not None:
            self.close()
            raise failure
        try:
            _lib.call("ph_comm_connect", self._h, b"".join(handles))
        except _lib.PalHistError:
            self.close()
            raise

    @property
    def handle(self):
        return self._h

    def allreduce_(self, t: torch.Tensor) -> torch.Tensor:
        """In-place sum over the ranks of a float64 CUDA tensor of 1 or 2 elements, on the current stream."""
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and 1 <= t.numel() <= 2):
            raise ValueError("allreduce_ expects a contiguous float64 CUDA tensor of 1 or 2 elements")
        with torch.cuda.device(t.device):
            _lib.call("ph_comm_allreduce_sum_f64", self._h, ptr(t), t.numel(), stream_ptr(t.device))
        return t

    def close(self):
        if self._h:
            _lib.load().ph_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def peer_comm(group, device):
    """Cached PeerComm of (group, device); None when the ranks cannot map each other's memory (no P2P between the
    devices, ranks on different nodes) or PH_COLLECTIVE=nccl asks for the library collective — decided jointly, so
    every rank takes the same path."""
    import torch.distributed as dist

    pg = None if group is True else group
    key = (id(pg) if pg is not None else 0, torch.device(device).index)
    if key in _cache:
        return _cache[key]
    comm = None
    # two joint decisions (MIN over the ranks), so that every rank takes the same path: whether to try at all, and
    # whether every rank managed to map every mailbox
    want = 0 if os.environ.get("PH_COLLECTIVE", "peer") == "nccl" or dist.get_backend(pg) != "nccl" else 1
    flag = torch.tensor([want], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=pg)
    if int(flag) == 1:
        ok = 1
        try:
            comm = PeerComm(pg, device)
        except (_lib.PalHistError, ValueError, RuntimeError):
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=pg)
        if int(flag) == 0:
            if comm is not None:
                comm.close()
            comm = None
    _cache[key] = comm
    return comm
