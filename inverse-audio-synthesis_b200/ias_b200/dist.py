"""Process-group plumbing for the one collective of the hot path: the all-gather of the embeddings.

``Communicator`` owns an NCCL communicator created through the C ABI (``ias_comm_*`` in libias_comm.so); the unique id
is made on rank 0 and handed to the other ranks with ``torch.distributed.broadcast_object_list`` (any initialised
backend works for that one-off exchange).  ``FullGatherLayer`` uses it when installed with ``use_communicator``; by
default it goes through ``torch.distributed`` (same NCCL underneath).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib

_active: Optional["Communicator"] = None


class Communicator:
    def __init__(self):
        if not (dist.is_available() and dist.is_initialized()):
            raise _lib.IasError("Communicator needs an initialised torch.distributed process group for the id exchange")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        lib = _lib.comm_lib()
        ident = (ctypes.c_uint8 * 128)()
        if self.rank == 0:
            _lib.check_comm(lib.ias_comm_unique_id(ident), "ias_comm_unique_id")
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=0)
        ident = (ctypes.c_uint8 * 128).from_buffer_copy(box[0])
        handle = ctypes.c_void_p()
        _lib.check_comm(lib.ias_comm_init(ident, self.rank, self.world, ctypes.byref(handle)), "ias_comm_init")
        self.handle = handle

    def all_gather(self, local: torch.Tensor) -> torch.Tensor:
        """[...] -> [W, ...] in rank order."""
        _lib.require_cuda(local, "all_gather input")
        local = local.contiguous().float()
        out = torch.empty((self.world,) + tuple(local.shape), dtype=torch.float32, device=local.device)
        rc = _lib.comm_lib().ias_comm_allgather(self.handle, _lib.ptr(local), _lib.ptr(out), local.numel(),
                                                _lib.current_stream(local.device))
        _lib.check_comm(rc, "ias_comm_allgather")
        return out

    def reduce_scatter(self, stacked: torch.Tensor) -> torch.Tensor:
        """[W, ...] (this rank's gradient w.r.t. the gathered tensor) -> sum over ranks of the own slice."""
        _lib.require_cuda(stacked, "reduce_scatter input")
        stacked = stacked.contiguous().float()
        out = torch.empty(tuple(stacked.shape[1:]), dtype=torch.float32, device=stacked.device)
        rc = _lib.comm_lib().ias_comm_reduce_scatter(self.handle, _lib.ptr(stacked), _lib.ptr(out), out.numel(),
                                                     _lib.current_stream(stacked.device))
        _lib.check_comm(rc, "ias_comm_reduce_scatter")
        return out

    def close(self):
        if self.handle:
            _lib.check_comm(_lib.comm_lib().ias_comm_destroy(self.handle), "ias_comm_destroy")
            self.handle = ctypes.c_void_p()


def use_communicator(comm: Optional[Communicator]) -> None:
    """Route FullGatherLayer through ``comm`` (None: back to torch.distributed)."""
    global _active
    _active = comm


def active() -> Optional[Communicator]:
    return _active
