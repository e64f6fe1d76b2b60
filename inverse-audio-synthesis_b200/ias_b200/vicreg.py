"""Drop-in for the reference's ``vicreg.py`` (VICReg / Projector / off_diagonal / FullGatherLayer) on libias_b200.so.

  * ``VICReg(cfg, backbone_audio, backbone_param)`` ........ vicreg.py:11-28, same attributes
  * ``VICReg.forward(audio, params) -> (x, y)`` ............ vicreg.py:30-33 (torch: the backbones are out of scope)
  * ``VICReg.loss(x, y) -> (loss, repr, std, cov)`` ........ vicreg.py:35-58, computed by ``ias_vicreg_loss``
  * ``FullGatherLayer`` .................................... vicreg.py:79-95, the gather the reference has commented
    out at vicreg.py:38-39; here it is live whenever ``torch.distributed`` is initialised with world size > 1, so the
    variance and covariance terms see the global batch while the invariance term stays on the local rows.

``cfg`` is any object with ``dim``, ``embeddim`` and ``vicreg.{mlp,batch_size,sim_coeff,std_coeff,cov_coeff}`` (an
omegaconf tree in the reference, a SimpleNamespace in the tests).  With the gather enabled ``cfg.vicreg.batch_size``
must be the *global* batch size: it is the covariance divisor (vicreg.py:47-48).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn

from . import _lib


def off_diagonal(x: torch.Tensor) -> torch.Tensor:
    """vicreg.py:73-76: the n(n-1) off-diagonal entries of a square matrix, row-major."""
    n, m = x.shape
    assert n == m
    return x.flatten()[:-1].view(n - 1, n + 1)[:, 1:].flatten()


def Projector(cfg, reprdim: int) -> nn.Sequential:
    """vicreg.py:61-70 (dense MLP, cuBLAS; not part of the hot path)."""
    spec = f"{reprdim}-{cfg.vicreg.mlp}" % cfg.embeddim
    f = [int(v) for v in spec.split("-")]
    layers = []
    for a, b in zip(f[:-2], f[1:-1]):
        layers += [nn.Linear(a, b), nn.BatchNorm1d(b), nn.ReLU(True)]
    layers.append(nn.Linear(f[-2], f[-1], bias=False))
    return nn.Sequential(*layers)


def _world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class FullGatherLayer(torch.autograd.Function):
    """Gather tensors from all processes, with gradients (vicreg.py:79-95).

    forward: tuple of W tensors in rank order;  backward: sum over ranks of the incoming gradients, own slice
    (all-reduce + slice in the reference == one reduce-scatter here)."""

    @staticmethod
    def forward(ctx, x):
        rank, world = _world()
        ctx.rank, ctx.world = rank, world
        if world == 1:
            return (x,)
        x = x.contiguous()
        from . import dist as ias_dist

        comm = ias_dist.active()
        if comm is not None and x.is_cuda:
            return tuple(comm.all_gather(x).to(x.dtype).unbind(0))
        out = torch.empty((world,) + tuple(x.shape), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out.view(-1), x.view(-1))
        return tuple(out.unbind(0))

    @staticmethod
    def backward(ctx, *grads):
        if ctx.world == 1:
            return grads[0]
        stacked = torch.stack(grads).contiguous()
        from . import dist as ias_dist

        comm = ias_dist.active()
        if comm is not None and stacked.is_cuda:
            return comm.reduce_scatter(stacked).to(stacked.dtype)
        if dist.get_backend() == "nccl":
            own = torch.empty_like(stacked[0])
            dist.reduce_scatter_tensor(own.view(-1), stacked.view(-1))
            return own
        dist.all_reduce(stacked)  # gloo has no reduce-scatter: the reference's all-reduce + slice (vicreg.py:92-95)
        return stacked[ctx.rank]


def aligned_workspace(nbytes: int, device) -> torch.Tensor:
    """float32 scratch of at least ``nbytes`` whose address is 1 KiB aligned (the caching allocator hands out
    512-byte aligned blocks: over-allocate and slice; the view keeps the block alive)."""
    raw = torch.empty(nbytes // 4 + 256, dtype=torch.float32, device=device)
    shift = (-raw.data_ptr() % 1024) // 4
    return raw[shift:shift + nbytes // 4]


def needs_backward(ctx) -> bool:
    """Inside ``autograd.Function.forward`` (where grad mode is off): will autograd call ``backward`` for x or y?"""
    return bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])


class _VicregLossFn(torch.autograd.Function):
    """``ias_vicreg_loss`` / ``ias_vicreg_loss_backward``.  The C ABI keeps what the backward needs (column means, std,
    the full Gram) in the workspace of the forward call "until the next forward call on it" (ias_b200.h).  A forward
    that may be differentiated therefore gets a workspace of its own, owned by its autograd node, so two ``loss()``
    calls before the first ``backward()`` -- two view pairs, a diagnostic loss on another batch -- cannot overwrite each
    other's state; only forwards under ``no_grad`` share the module-wide scratch buffer."""

    @staticmethod
    def forward(ctx, x, y, local_row0, b_local, cfg_batch, embeddim, sim, stdc, covc, holder):
        _lib.require_cuda(x, "VICReg.loss x")
        _lib.require_cuda(y, "VICReg.loss y")
        xc = x.detach().to(torch.float32).contiguous()
        yc = y.detach().to(torch.float32).contiguous()
        B, D = xc.shape
        lib = _lib.lib()
        if needs_backward(ctx):
            ws = aligned_workspace(lib.ias_vicreg_workspace_bytes(B, D), xc.device)
        else:
            ws = holder.workspace(B, D, xc.device)
        out4 = torch.empty(4, dtype=torch.float32, device=xc.device)
        with _lib.on_device(xc):
            rc = lib.ias_vicreg_loss(_lib.ptr(xc), _lib.ptr(yc), B, local_row0, b_local, cfg_batch, D, embeddim, sim,
                                     stdc, covc, _lib.ptr(out4), _lib.ptr(ws), ws.numel() * ws.element_size(),
                                     _lib.current_stream(xc.device))
        _lib.check(rc, "ias_vicreg_loss")
        ctx.save_for_backward(xc, yc)
        ctx.args = (B, local_row0, b_local, cfg_batch, D, embeddim, sim, stdc, covc)
        ctx.ws = ws
        return out4[0], out4[1], out4[2], out4[3]

    @staticmethod
    def backward(ctx, g0, g1, g2, g3):
        xc, yc = ctx.saved_tensors
        B, local_row0, b_local, cfg_batch, D, embeddim, sim, stdc, covc = ctx.args
        zero = torch.zeros((), dtype=torch.float32, device=xc.device)
        gout = torch.stack([g if g is not None else zero for g in (g0, g1, g2, g3)]).to(torch.float32).contiguous()
        gx = torch.empty_like(xc)
        gy = torch.empty_like(yc)
        ws = ctx.ws
        with _lib.on_device(xc):
            rc = _lib.lib().ias_vicreg_loss_backward(
                _lib.ptr(xc), _lib.ptr(yc), B, local_row0, b_local, cfg_batch, D, embeddim, sim, stdc, covc,
                _lib.ptr(gout), _lib.ptr(gx), _lib.ptr(gy), _lib.ptr(ws), ws.numel() * ws.element_size(),
                _lib.current_stream(xc.device))
        _lib.check(rc, "ias_vicreg_loss_backward")
        return gx, gy, None, None, None, None, None, None, None, None


class _Workspace:
    """Module-wide scratch for forwards that will not be differentiated."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None

    def workspace(self, B: int, D: int, device) -> torch.Tensor:
        need = _lib.lib().ias_vicreg_workspace_bytes(B, D)
        if self.buf is None or self.buf.device != device or self.buf.numel() * 4 < need:
            self.buf = aligned_workspace(need, device)
        return self.buf


def vicreg_loss(x, y, cfg_batch_size: int, embeddim: int, sim_coeff: float, std_coeff: float, cov_coeff: float,
                gather: bool = True, _holder: Optional[_Workspace] = None, strict_batch: bool = True):
    """Functional form of ``VICReg.loss`` -> (loss, repr_loss, std_loss, cov_loss), 0-d tensors.

    ``gather`` (and an initialised process group with world size > 1) makes the variance / covariance terms see the
    global batch, as the reference intends (vicreg.py:38-39).  ``cfg_batch_size`` is the covariance divisor
    (vicreg.py:47-48) and must then be the GLOBAL batch size: a per-GPU value would inflate ``cov_loss`` by ~W^2
    without any other symptom, so a mismatch with ``world * x.shape[0]`` raises unless ``strict_batch=False``."""
    holder = _holder if _holder is not None else _Workspace()
    rank, world = _world()
    b_local = x.shape[0]
    if gather and world > 1:
        if strict_batch and int(cfg_batch_size) != world * b_local:
            raise _lib.IasError(
                f"VICReg.loss gathers the embeddings of {world} ranks ({world} x {b_local} rows) but "
                f"cfg.vicreg.batch_size = {cfg_batch_size}: with the gather enabled it is the covariance divisor of the "
                f"GLOBAL batch (vicreg.py:47-48) and must be {world * b_local}; the per-GPU size belongs in "
                "SynthConfig.batch_size.  Pass gather=False for the reference's per-rank statistics, or "
                "strict_batch=False to keep this divisor.")
        from . import dist as ias_dist

        ex = ias_dist.fused_exchange()
        if isinstance(ex, ias_dist.StatsExchange):
            return ias_dist.FusedStatsLoss.apply(x, y, ex, int(cfg_batch_size), int(embeddim), float(sim_coeff),
                                                 float(std_coeff), float(cov_coeff))
        if ex is not None:
            return ias_dist.FusedGatherLoss.apply(x, y, ex, int(cfg_batch_size), int(embeddim), float(sim_coeff),
                                                  float(std_coeff), float(cov_coeff))
        x_all = torch.cat(FullGatherLayer.apply(x), dim=0)
        y_all = torch.cat(FullGatherLayer.apply(y), dim=0)
        row0 = rank * b_local
    else:
        x_all, y_all, row0 = x, y, 0
    return _VicregLossFn.apply(x_all, y_all, row0, b_local, int(cfg_batch_size), int(embeddim), float(sim_coeff),
                               float(std_coeff), float(cov_coeff), holder)


class VICReg(nn.Module):
    def __init__(self, cfg, backbone_audio: nn.Module, backbone_param: nn.Module, gather: bool = True,
                 strict_batch: bool = True):
        """``gather``: with an initialised process group of world size > 1, the variance / covariance terms are
        computed over the global batch (the FullGatherLayer call the reference has commented out at vicreg.py:38-39 and
        its README calls a bug; BASELINE north star).  ``cfg.vicreg.batch_size`` must then be the global batch size
        (checked, see ``vicreg_loss``).  ``gather=False`` reproduces the reference as it runs today: per-rank
        statistics with ``cfg.vicreg.batch_size`` as the divisor."""
        super().__init__()
        self.cfg = cfg
        self.reprdim = cfg.dim
        self.embeddim = cfg.embeddim
        self.backbone_audio = backbone_audio
        self.backbone_param = backbone_param
        self.projector = Projector(cfg, self.reprdim)
        self.gather = gather
        self.strict_batch = strict_batch
        self._holder = _Workspace()

    def forward(self, audio, params):
        x = self.projector(self.backbone_audio(audio))
        y = self.projector(self.backbone_param(params))
        return x, y

    def loss(self, x, y):
        v = self.cfg.vicreg
        return vicreg_loss(x, y, v.batch_size, self.embeddim, v.sim_coeff, v.std_coeff, v.cov_coeff,
                           gather=self.gather, _holder=self._holder, strict_batch=self.strict_batch)


def exclude_bias_and_norm(p):
    return p.ndim == 1
