"""Drop-in for the reference's ``pqmf.PQMF`` (pqmf.py:9-55) running on ``libias_b200.so``.

Same constructor, attributes, persistent buffers (``H[N,1,taps+1]``, ``G[1,N,taps+1]``, ``updown_filter[N,N,N]``
-- so checkpoints with ``gram.H`` etc. load unchanged) and methods ``forward`` / ``analysis`` / ``synthesis``.
The filter design is host-side one-off work (scipy, like the reference); the filtering itself is
``ias_pqmf_analysis`` / ``ias_pqmf_synthesis``.  Callers: ``vicreg_audio_params.py:40``, ``audioembed.py:38``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
from scipy import signal as sig
from torch import nn

from . import _lib


def design_filters(N: int, taps: int, cutoff: float, beta: float) -> Tuple[np.ndarray, np.ndarray]:
    """Cosine-modulated analysis/synthesis banks of the reference (pqmf.py:18-30), float64 [N, taps+1] each.

    The modulation is centred on ``(taps-1)/2`` exactly as the reference does (its own TODO at pqmf.py:26 says the
    textbook value would be ``taps/2``); parity with the reference's bands requires keeping that."""
    prototype = sig.firwin(taps + 1, cutoff, window=("kaiser", beta))
    k = np.arange(N)[:, None]
    j = np.arange(taps + 1)[None, :]
    theta = (2 * k + 1) * (np.pi / (2 * N)) * (j - ((taps - 1) / 2))
    phase = ((-1.0) ** k) * np.pi / 4
    return 2 * prototype * np.cos(theta + phase), 2 * prototype * np.cos(theta - phase)


def cosine_modulation_factors(N: int, taps: int, cutoff: float, beta: float) -> Tuple[np.ndarray, np.ndarray]:
    """(g[taps+1], c[N, 2N]) float32 with H[k, j] == g[j] * c[k, j % 2N] up to fp32 rounding: the cosine of
    pqmf.py:21-30 has period 4N in the tap index and flips sign every 2N taps, so the sign goes into the prototype
    and the bank collapses to 2N polyphase partial sums followed by an N x 2N modulation."""
    prototype = sig.firwin(taps + 1, cutoff, window=("kaiser", beta))
    j = np.arange(taps + 1)
    g = 2.0 * prototype * np.where((j // (2 * N)) % 2 == 0, 1.0, -1.0)
    k = np.arange(N)[:, None]
    r = np.arange(2 * N)[None, :]
    c = np.cos((2 * k + 1) * (np.pi / (2 * N)) * (r - ((taps - 1) / 2)) + ((-1.0) ** k) * np.pi / 4)
    return g.astype(np.float32), c.astype(np.float32)


class PQMF(nn.Module):
    def __init__(self, N: int = 4, taps: int = 62, cutoff: float = 0.15, beta: float = 9.0):
        super().__init__()
        self.N = N
        self.taps = taps
        self.cutoff = cutoff
        self.beta = beta
        H, G = design_filters(N, taps, cutoff, beta)
        self.register_buffer("H", torch.from_numpy(H[:, None, :]).float())
        self.register_buffer("G", torch.from_numpy(G[None, :, :]).float())
        updown = torch.zeros((N, N, N)).float()
        updown[torch.arange(N), torch.arange(N), 0] = 1.0
        self.register_buffer("updown_filter", updown)  # unused by the kernels; kept for state-dict parity
        self.pad_fn = nn.ConstantPad1d(taps // 2, 0.0)
        self.polyphase = True  # use the cosine-modulated fast path when H is the designed filter
        self._host_cache = {}

    def _taps(self, which: str):
        """(device [N,K], host [N,K], polyphase factors or None) of buffer ``which``; refreshed when the buffer
        changes.  The factors are offered to the kernel only while the buffer still holds exactly the filter this
        module designs (a checkpoint may have loaded other taps: then the direct form runs)."""
        buf = getattr(self, which)
        key = (buf.data_ptr(), buf._version, buf.device)
        hit = self._host_cache.get(which)
        if hit is None or hit[0] != key:
            dev = buf.reshape(self.N, -1).contiguous()
            host = dev.detach().to("cpu", torch.float32).contiguous()
            factors = None
            H, G = design_filters(self.N, self.taps, self.cutoff, self.beta)
            if torch.equal(host, torch.from_numpy(H if which == "H" else G).float()):
                # the synthesis bank shares the signed prototype; its modulation matrix is rebuilt by the kernel
                g, c = cosine_modulation_factors(self.N, self.taps, self.cutoff, self.beta)
                factors = (torch.from_numpy(g).contiguous(), torch.from_numpy(c).contiguous())
            hit = (key, dev, host, factors)
            self._host_cache[which] = hit
        return hit[1], hit[2], hit[3]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.analysis(x)

    def analysis_image(self, x: torch.Tensor, mean, std, image_shape: Optional[Tuple[int, int]] = None,
                       row_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``AudioEmbedding._preprocess`` in one kernel (audioembed.py:36-49): bands -> ``reshape(-1, N, H, W)`` ->
        ``torchvision.transforms.Normalize(mean, std)``.  ``mean`` / ``std`` are length-N sequences (the reference:
        ImageNet constants, vicreg_audio_params.py:60-62).  Returns ``[B, N, H, W]`` when ``image_shape=(H, W)`` is given
        (H*W must equal L; the reference uses (240, 245) for 4 s), else ``[B, N, L]``."""
        mean_t = torch.as_tensor(mean, dtype=torch.float32).reshape(-1).contiguous().cpu()
        std_t = torch.as_tensor(std, dtype=torch.float32).reshape(-1).contiguous().cpu()
        if mean_t.numel() != self.N or std_t.numel() != self.N:
            raise ValueError(f"analysis_image: mean/std need {self.N} entries")
        out = self.analysis(x, row_scale=row_scale, _norm=(mean_t, std_t))
        if image_shape is not None:
            h, w = image_shape
            if h * w != out.shape[2]:
                raise ValueError(f"analysis_image: {h}x{w} != L={out.shape[2]}")
            out = out.reshape(-1, self.N, h, w)
        return out

    def analysis(self, x: torch.Tensor, row_scale: Optional[torch.Tensor] = None, _norm=None) -> torch.Tensor:
        """x [B,1,T] -> [B,N,L]  (pqmf.py:49-50).  ``row_scale`` [B] optionally scales each row on the fly."""
        if x.dim() != 3 or x.shape[1] != 1:
            raise ValueError(f"PQMF.analysis expects [B,1,T], got {tuple(x.shape)}")
        _lib.require_cuda(x, "PQMF.analysis input")
        if x.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError("PQMF.analysis has no backward: the synth output it filters carries no gradient")
        x = x.detach().to(torch.float32).contiguous()
        B, _, T = x.shape
        dev, host, factors = self._taps("H")
        proto, mod = factors if (factors is not None and self.polyphase) else (None, None)
        K = host.shape[1]
        lib = _lib.lib()
        L = lib.ias_pqmf_out_len(T, self.N, K)
        if L <= 0:
            raise ValueError(f"PQMF.analysis: input too short (T={T})")
        out = torch.empty((B, self.N, L), dtype=torch.float32, device=x.device)
        if row_scale is not None:
            row_scale = row_scale.detach().to(torch.float32).contiguous()
        if _norm is not None:
            mean_t, std_t = _norm
            norm_dev = torch.cat([mean_t, std_t]).to(x.device)
            with _lib.on_device(x):
                rc = lib.ias_pqmf_analysis_image(_lib.ptr(x), _lib.ptr(dev), _lib.ptr(host), _lib.ptr(proto),
                                                 _lib.ptr(mod), _lib.ptr(row_scale), _lib.ptr(mean_t), _lib.ptr(std_t),
                                                 _lib.ptr(norm_dev), _lib.ptr(out), B, T, self.N, K,
                                                 _lib.current_stream(x.device))
            _lib.check(rc, "ias_pqmf_analysis_image")
            return out
        with _lib.on_device(x):
            rc = lib.ias_pqmf_analysis(_lib.ptr(x), _lib.ptr(dev), _lib.ptr(host), _lib.ptr(proto), _lib.ptr(mod),
                                       _lib.ptr(row_scale), _lib.ptr(out), B, T, self.N, K,
                                       _lib.current_stream(x.device))
        _lib.check(rc, "ias_pqmf_analysis")
        return out

    def analysis_pooled(self, x: torch.Tensor, bins: int, row_scale: Optional[torch.Tensor] = None):
        """``analysis(x)`` plus ``adaptive_avg_pool1d(bands.abs().reshape(B, 1, -1), bins)`` computed by the same kernel
        (the harness bridge between the bands and the embeddings; not a reference surface) -> (bands
        [B,N,L], feat [B,bins]).  Raises ``IasError`` when the shape has no fused path; pool separately then."""
        if x.dim() != 3 or x.shape[1] != 1:
            raise ValueError(f"PQMF.analysis_pooled expects [B,1,T], got {tuple(x.shape)}")
        _lib.require_cuda(x, "PQMF.analysis_pooled input")
        x = x.detach().to(torch.float32).contiguous()
        B, _, T = x.shape
        dev, host, factors = self._taps("H")
        proto, mod = factors if (factors is not None and self.polyphase) else (None, None)
        K = host.shape[1]
        lib = _lib.lib()
        L = lib.ias_pqmf_out_len(T, self.N, K)
        if L <= 0:
            raise ValueError(f"PQMF.analysis_pooled: input too short (T={T})")
        out = torch.empty((B, self.N, L), dtype=torch.float32, device=x.device)
        feat = torch.empty((B, bins), dtype=torch.float32, device=x.device)
        nbytes = lib.ias_pqmf_pool_workspace_bytes(B, T, self.N, K)
        ws = torch.empty((max(nbytes, 4) + 3) // 4, dtype=torch.float32, device=x.device)
        if row_scale is not None:
            row_scale = row_scale.detach().to(torch.float32).contiguous()
        with _lib.on_device(x):
            rc = lib.ias_pqmf_analysis_pooled(_lib.ptr(x), _lib.ptr(dev), _lib.ptr(host), _lib.ptr(proto), _lib.ptr(mod),
                                              _lib.ptr(row_scale), _lib.ptr(out), _lib.ptr(feat), bins, _lib.ptr(ws),
                                              ws.numel() * 4, B, T, self.N, K, _lib.current_stream(x.device))
        _lib.check(rc, "ias_pqmf_analysis_pooled")
        return out, feat

    def synthesis(self, x: torch.Tensor) -> torch.Tensor:
        """z [B,N,L] -> [B,1,L*N]  (pqmf.py:52-55)."""
        if x.dim() != 3 or x.shape[1] != self.N:
            raise ValueError(f"PQMF.synthesis expects [B,{self.N},L], got {tuple(x.shape)}")
        _lib.require_cuda(x, "PQMF.synthesis input")
        if x.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError("PQMF.synthesis has no backward")
        x = x.detach().to(torch.float32).contiguous()
        B, _, L = x.shape
        dev, host, factors = self._taps("G")
        proto = factors[0] if (factors is not None and self.polyphase) else None
        K = host.shape[1]
        out = torch.empty((B, 1, L * self.N), dtype=torch.float32, device=x.device)
        with _lib.on_device(x):
            rc = _lib.lib().ias_pqmf_synthesis(_lib.ptr(x), _lib.ptr(dev), _lib.ptr(host), _lib.ptr(proto),
                                               _lib.ptr(out), B, L, self.N, K, _lib.current_stream(x.device))
        _lib.check(rc, "ias_pqmf_synthesis")
        # conv1d(padding=taps//2) keeps L*N samples for even `taps` (the default) and drops the last one for odd
        keep = L * self.N + 2 * (self.taps // 2) - self.taps
        return out if keep == out.shape[2] else out[:, :, :keep]
