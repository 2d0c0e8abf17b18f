"""Multi-GPU plumbing: one process per GPU, units (chains / pool members / live points) sharded by rank.

The reference has no distributed layer (single JVM process).  Chains and pool members need
no exchange at all; the only collectives on the path are the SLQ global lowest-logL selection and sample-moment
sums, both through the C ABI's ``thy_comm_*`` (NCCL over NVLink).  ``torch.distributed`` is used for the
rendezvous only: it carries the 128-byte NCCL unique id from rank 0 to the other ranks (gloo or nccl backend).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

from ._capi import check, lib, register

_vp = C.c_void_p
register("thy_comm_unique_id", C.c_int, [C.POINTER(C.c_ubyte)])
register("thy_comm_create", C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_ubyte), C.POINTER(_vp)])
register("thy_comm_create_all", C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(_vp)])
register("thy_comm_destroy", C.c_int, [_vp])
register("thy_comm_rank", C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)])
register("thy_comm_all_reduce_sum", C.c_int, [_vp, _vp, C.c_int64, _vp])


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of ``total`` units owned by ``rank``: the first ``total % world`` ranks get one extra
    unit.  Every unit belongs to exactly one rank and the blocks are in rank order."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_world() -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def exchange_unique_id(make_id, rank: int, world: int) -> Optional[bytes]:
    """Rank 0 calls ``make_id()`` and the result is broadcast over torch.distributed (any backend)."""
    if world == 1:
        return None
    import torch.distributed as dist
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


class Communicator:
    """``thy_comm_t``: an NCCL communicator with one rank per process/GPU."""

    def __init__(self, rank: int = 0, world: int = 1, unique_id: Optional[bytes] = None):
        self._lib = lib()
        self.rank, self.world = rank, world
        h = _vp()
        uid = None
        if world > 1:
            if unique_id is None or len(unique_id) != 128:
                raise ValueError("a 128-byte NCCL unique id is required for world > 1")
            uid = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        check(self._lib.thy_comm_create(rank, world, uid, C.byref(h)))
        self._h = h

    @classmethod
    def createAll(cls, devices) -> list:
        """Single-process mode (``thy_comm_create_all``): one communicator per listed GPU of THIS process; use each from
        its own host thread, bound with ``setDevice``."""
        devices = list(devices)
        n = len(devices)
        arr = (C.c_int * n)(*devices)
        handles = (_vp * n)()
        check(lib().thy_comm_create_all(n, arr, handles))
        out = []
        for i in range(n):
            c = object.__new__(cls)
            c._lib, c.rank, c.world, c._h = lib(), i, n, _vp(handles[i])
            out.append(c)
        return out

    @staticmethod
    def uniqueId() -> bytes:
        buf = (C.c_ubyte * 128)()
        check(lib().thy_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def fromTorchDistributed(cls) -> "Communicator":
        """Build from an initialised torch.distributed process group (torchrun launch); the CUDA device must already
        be selected (``torch.cuda.set_device(LOCAL_RANK)``)."""
        import torch.distributed as dist
        if not dist.is_initialized():
            return cls(0, 1, None)
        rank, world = dist.get_rank(), dist.get_world_size()
        return cls(rank, world, exchange_unique_id(cls.uniqueId, rank, world))

    @property
    def handle(self):
        return self._h

    def allReduceSum(self, dev_ptr: int, count: int, stream: int = 0) -> None:
        check(self._lib.thy_comm_all_reduce_sum(self._h, dev_ptr, count, stream))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.thy_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def deviceCount() -> int:
    n = C.c_int()
    check(lib().thy_device_count(C.byref(n)))
    return n.value


def setDevice(device: int) -> None:
    """Bind the calling thread to a GPU (``thy_set_device``); handles created afterwards live there."""
    check(lib().thy_set_device(device))
