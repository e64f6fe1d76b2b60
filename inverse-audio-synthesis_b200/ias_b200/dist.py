"""Process-group plumbing for the one exchange step of the hot path: what the loss needs from the other
ranks' embeddings.  Three routes, same result as the single-process loss on the rank-ordered concatenation:

  * ``StatsExchange`` + ``FusedStatsLoss`` (default of bench.py): every rank reduces its own rows and pushes a 0.4 MB
    summary (mean, centred second moments, Gram) to its peers over NVLink from inside the loss kernels; no collective;
  * ``EmbeddingExchange`` + ``FusedGatherLoss``: the statistics kernel reads every peer's embeddings over NVLink;
  * ``FullGatherLayer`` (vicreg.py): an NCCL all-gather through torch.distributed or ``Communicator`` below.

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
from .vicreg import aligned_workspace, needs_backward

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


class StatsExchange:
    """Peer-visible inbox of this rank for the statistics exchange (``ias_vicreg_loss_stats``): one symmetric-memory
    block per rank holding a 1 KiB header (arrival flags, step counter) and ``world`` x 2 summary slots.  Allocation and
    the address exchange go through torch symmetric memory; every byte that moves afterwards is written by the loss
    kernels themselves."""

    def __init__(self, D: int, device, group=None):
        import torch.distributed._symmetric_memory as symm

        if not (dist.is_available() and dist.is_initialized()):
            raise _lib.IasError("StatsExchange needs an initialised torch.distributed process group")
        group = group if group is not None else dist.group.WORLD
        self.D, self.device = int(D), torch.device(device)
        world = dist.get_world_size(group)
        nbytes = _lib.lib().ias_vicreg_stats_buffer_bytes(world, self.D)
        self.buf = symm.empty((nbytes // 4,), dtype=torch.float32, device=self.device)
        self.buf.zero_()
        self.buf.view(torch.int32)[32] = 1  # step counter (ias_b200.h: the int at byte offset 128 starts at 1)
        self.hdl = symm.rendezvous(self.buf, group)
        self.world, self.rank = int(self.hdl.world_size), int(self.hdl.rank)
        self.ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in self.hdl.buffer_ptrs])
        torch.cuda.synchronize(self.device)
        dist.barrier(group)  # nobody pushes into an inbox that is still being zeroed
        self._scratch = None

    def ws(self, b_local: int) -> torch.Tensor:
        need = _lib.lib().ias_vicreg_workspace_bytes(b_local, self.D)
        if self._scratch is None or self._scratch.numel() * 4 < need:
            self._scratch = aligned_workspace(need, self.device)
        return self._scratch

    def step(self) -> int:
        """Step counter as the device sees it (synchronises; diagnostics only)."""
        return int(self.buf.view(torch.int32)[32].item())


def use_fused_gather(exchange) -> None:
    """Route VICReg.loss through ``exchange``: a ``StatsExchange`` (summary exchange inside the loss kernels) or an
    ``EmbeddingExchange`` (peer reads of the embeddings).  None: back to FullGatherLayer + ias_vicreg_loss."""
    global _exchange
    _exchange = exchange


def fused_exchange():
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
        # a forward that may be differentiated owns its workspace (see vicreg._VicregLossFn)
        grad = needs_backward(ctx)
        ws = aligned_workspace(_lib.lib().ias_vicreg_gather_workspace_bytes(ex.world, ex.b_local, ex.D),
                               x.device) if grad else ex.ws()
        out4 = torch.empty(4, dtype=torch.float32, device=x.device)
        with _lib.on_device(x):
            rc = _lib.lib().ias_vicreg_loss_gather(ex.x_ptrs, ex.y_ptrs, ex.world, ex.rank, ex.b_local, cfg_batch, ex.D,
                                                   embeddim, sim, stdc, covc, _lib.ptr(out4), _lib.ptr(ws),
                                                   ws.numel() * 4, _lib.current_stream(x.device))
        _lib.check(rc, "ias_vicreg_loss_gather")
        ctx.ex, ctx.args, ctx.ws = ex, (cfg_batch, embeddim, sim, stdc, covc), ws
        return out4[0], out4[1], out4[2], out4[3]

    @staticmethod
    def backward(ctx, g0, g1, g2, g3):
        ex = ctx.ex
        cfg_batch, embeddim, sim, stdc, covc = ctx.args
        zero = torch.zeros((), dtype=torch.float32, device=ex.device)
        gout = torch.stack([g if g is not None else zero for g in (g0, g1, g2, g3)]).float().contiguous()
        gx = torch.empty((ex.b_local, ex.D), dtype=torch.float32, device=ex.device)
        gy = torch.empty_like(gx)
        ws = ctx.ws
        with _lib.on_device(gx):
            rc = _lib.lib().ias_vicreg_loss_gather_backward(ex.world, ex.rank, ex.b_local, cfg_batch, ex.D, embeddim,
                                                            sim, stdc, covc, _lib.ptr(gout), _lib.ptr(gx), _lib.ptr(gy),
                                                            _lib.ptr(ws), ws.numel() * 4,
                                                            _lib.current_stream(ex.device))
        _lib.check(rc, "ias_vicreg_loss_gather_backward")
        return gx, gy, None, None, None, None, None, None


class FusedStatsLoss(torch.autograd.Function):
    """loss(x_local, y_local) over the global batch through the statistics exchange: forward =
    ``ias_vicreg_loss_stats``, backward = ``ias_vicreg_loss_stats_backward`` (own rows, no communication)."""

    @staticmethod
    def forward(ctx, x, y, ex, cfg_batch, embeddim, sim, stdc, covc):
        _lib.require_cuda(x, "VICReg.loss x")
        _lib.require_cuda(y, "VICReg.loss y")
        if x.dim() != 2 or x.shape[1] != ex.D or tuple(y.shape) != tuple(x.shape):
            raise _lib.IasError(f"StatsExchange was built for D={ex.D}, got x {tuple(x.shape)} y {tuple(y.shape)}")
        xc = x.detach().to(torch.float32).contiguous()
        yc = y.detach().to(torch.float32).contiguous()
        b_local = xc.shape[0]
        lib = _lib.lib()
        grad = needs_backward(ctx)
        ws = aligned_workspace(lib.ias_vicreg_workspace_bytes(b_local, ex.D), xc.device) if grad else ex.ws(b_local)
        out4 = torch.empty(4, dtype=torch.float32, device=xc.device)
        with _lib.on_device(xc):
            rc = lib.ias_vicreg_loss_stats(_lib.ptr(xc), _lib.ptr(yc), ex.ptrs, ex.world, ex.rank, b_local, cfg_batch,
                                           ex.D, embeddim, sim, stdc, covc, _lib.ptr(out4), _lib.ptr(ws),
                                           ws.numel() * 4, _lib.current_stream(xc.device))
        _lib.check(rc, "ias_vicreg_loss_stats")
        ctx.save_for_backward(xc, yc)
        ctx.args, ctx.ws, ctx.world = (cfg_batch, ex.D, embeddim, sim, stdc, covc), ws, ex.world
        return out4[0], out4[1], out4[2], out4[3]

    @staticmethod
    def backward(ctx, g0, g1, g2, g3):
        xc, yc = ctx.saved_tensors
        cfg_batch, D, embeddim, sim, stdc, covc = ctx.args
        zero = torch.zeros((), dtype=torch.float32, device=xc.device)
        gout = torch.stack([g if g is not None else zero for g in (g0, g1, g2, g3)]).float().contiguous()
        gx, gy = torch.empty_like(xc), torch.empty_like(yc)
        ws = ctx.ws
        with _lib.on_device(xc):
            rc = _lib.lib().ias_vicreg_loss_stats_backward(_lib.ptr(xc), _lib.ptr(yc), ctx.world, xc.shape[0], cfg_batch,
                                                           D, embeddim, sim, stdc, covc, _lib.ptr(gout), _lib.ptr(gx),
                                                           _lib.ptr(gy), _lib.ptr(ws), ws.numel() * 4,
                                                           _lib.current_stream(xc.device))
        _lib.check(rc, "ias_vicreg_loss_stats_backward")
        return gx, gy, None, None, None, None, None, None
