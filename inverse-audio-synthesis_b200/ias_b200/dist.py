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


# ----------------------------------------------------------------------------------------------------------------
# Fused gather: the loss kernels read every rank's embeddings over NVLink themselves (ias_vicreg_loss_gather)
# ----------------------------------------------------------------------------------------------------------------
_exchange: Optional["EmbeddingExchange"] = None


class EmbeddingExchange:
    """Peer-visible staging buffer for this rank's [2, B_local, D] embeddings (torch symmetric memory: one
    cudaMalloc'd block per rank, IPC-mapped into every peer, plus a signal pad for device-side barriers)."""

    def __init__(self, b_local: int, D: int, device, group=None):
        import torch.distributed._symmetric_memory as symm

        if not (dist.is_available() and dist.is_initialized()):
            raise _lib.IasError("EmbeddingExchange needs an initialised torch.distributed process group")
        group = group if group is not None else dist.group.WORLD
        self.b_local, self.D, self.device = int(b_local), int(D), torch.device(device)
        self.buf = symm.empty((2, self.b_local, self.D), dtype=torch.float32, device=self.device)
        self.hdl = symm.rendezvous(self.buf, group)
        self.world, self.rank = int(self.hdl.world_size), int(self.hdl.rank)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        half = self.b_local * self.D * 4
        self.x_ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        self.y_ptrs = (ctypes.c_void_p * self.world)(*[p + half for p in ptrs])
        self.workspace: Optional[torch.Tensor] = None

    def publish(self, x: torch.Tensor, y: torch.Tensor) -> None:
        """Stream-ordered: wait until every peer has finished reading the previous contents, write the new ones,
        make them visible to every peer."""
        self.hdl.barrier(channel=0)
        self.buf[0].copy_(x.detach())
        self.buf[1].copy_(y.detach())
        self.hdl.barrier(channel=1)

    def ws(self) -> torch.Tensor:
        need = _lib.lib().ias_vicreg_gather_workspace_bytes(self.world, self.b_local, self.D)
        if self.workspace is None or self.workspace.numel() * 4 < need:
            raw = torch.empty(need // 4 + 256, dtype=torch.float32, device=self.device)
            shift = (-raw.data_ptr() % 1024) // 4
            self.workspace = raw[shift:shift + need // 4]
            self._raw = raw
        return self.workspace


def use_fused_gather(exchange: Optional[EmbeddingExchange]) -> None:
    """Route VICReg.loss through the fused gather kernels (None: back to FullGatherLayer + ias_vicreg_loss)."""
    global _exchange
    _exchange = exchange


def fused_exchange() -> Optional[EmbeddingExchange]:
    return _exchange


class FusedGatherLoss(torch.autograd.Function):
    """loss(x_local, y_local) over the global batch without a separate collective: forward = ias_vicreg_loss_gather,
    backward = ias_vicreg_loss_gather_backward (own rows, no communication)."""

    @staticmethod
    def forward(ctx, x, y, ex, cfg_batch, embeddim, sim, stdc, covc):
        _lib.require_cuda(x, "VICReg.loss x")
        if tuple(x.shape) != (ex.b_local, ex.D) or tuple(y.shape) != (ex.b_local, ex.D):
            raise _lib.IasError(f"EmbeddingExchange was built for [{ex.b_local},{ex.D}], got {tuple(x.shape)}")
        ex.publish(x.float(), y.float())
        ws = ex.ws()
        out4 = torch.empty(4, dtype=torch.float32, device=x.device)
        with _lib.on_device(x):
            rc = _lib.lib().ias_vicreg_loss_gather(ex.x_ptrs, ex.y_ptrs, ex.world, ex.rank, ex.b_local, cfg_batch, ex.D,
                                                   embeddim, sim, stdc, covc, _lib.ptr(out4), _lib.ptr(ws),
                                                   ws.numel() * 4, _lib.current_stream(x.device))
        _lib.check(rc, "ias_vicreg_loss_gather")
        ctx.ex, ctx.args = ex, (cfg_batch, embeddim, sim, stdc, covc)
        return out4[0], out4[1], out4[2], out4[3]

    @staticmethod
    def backward(ctx, g0, g1, g2, g3):
        ex = ctx.ex
        cfg_batch, embeddim, sim, stdc, covc = ctx.args
        zero = torch.zeros((), dtype=torch.float32, device=ex.device)
        gout = torch.stack([g if g is not None else zero for g in (g0, g1, g2, g3)]).float().contiguous()
        gx = torch.empty((ex.b_local, ex.D), dtype=torch.float32, device=ex.device)
        gy = torch.empty_like(gx)
        ws = ex.ws()
        with _lib.on_device(gx):
            rc = _lib.lib().ias_vicreg_loss_gather_backward(ex.world, ex.rank, ex.b_local, cfg_batch, ex.D, embeddim,
                                                            sim, stdc, covc, _lib.ptr(gout), _lib.ptr(gx), _lib.ptr(gy),
                                                            _lib.ptr(ws), ws.numel() * 4,
                                                            _lib.current_stream(ex.device))
        _lib.check(rc, "ias_vicreg_loss_gather_backward")
        return gx, gy, None, None, None, None, None, None
