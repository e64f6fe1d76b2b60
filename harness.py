"""Measurement / test harness shared by bench.py, __graft_entry__.smoke() and tests/ (not part of the package).

* ``bridge``: the stand-in for the out-of-scope backbone between PQMF bands and the VICReg embeddings
  (SURVEY.md 8d "harness bridge"): abs-mean pool of the flattened bands to 256 bins, fixed seeded projections.
  Identical torch code runs on the oracle (CPU) and device paths; it is not a reference component and is timed
  separately from the three hot-path stages.
* ``oracle_front_end``: synth -> PQMF -> bridge -> VICReg loss on the CPU through ``oracle/`` -- the checker and the
  CPU baseline.  Only tests, smoke() and bench.py's baseline legs call it.
"""
from __future__ import annotations

import os
import sys
import time
from typing import Dict, Tuple

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "inverse-audio-synthesis_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

EMBED_DIM = 256
NPARAMS = 78
LAST_BRIDGE_PATH = None  # which device path analysis_bridge took last (recorded in the bench line)


def bridge_weights(device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    g = torch.Generator().manual_seed(1234)
    wa = torch.randn((EMBED_DIM, EMBED_DIM), generator=g) / EMBED_DIM ** 0.5
    wp = torch.randn((NPARAMS, EMBED_DIM), generator=g) / NPARAMS ** 0.5
    return wa.to(device), wp.to(device)


def bridge(bands: torch.Tensor, params: torch.Tensor, wa: torch.Tensor, wp: torch.Tensor):
    """bands [B,N,L], params [B,78] -> x [B,256], y [B,256]."""
    B = bands.shape[0]
    if bands.is_cuda:  # same pooling through the library's harness kernel (torch's CUDA adaptive pool is ~13x slower)
        from ias_b200 import _lib

        flat = bands.contiguous().view(B, -1)
        feat = torch.empty((B, EMBED_DIM), dtype=torch.float32, device=bands.device)
        with _lib.on_device(bands):
            _lib.check(_lib.lib().ias_abs_avg_pool(_lib.ptr(flat), _lib.ptr(feat), B, flat.shape[1], EMBED_DIM,
                                                   _lib.current_stream(bands.device)), "ias_abs_avg_pool")
    else:
        feat = torch.nn.functional.adaptive_avg_pool1d(bands.abs().reshape(B, 1, -1), EMBED_DIM).squeeze(1)
    return feat @ wa, params @ wp


def analysis_bridge(gram, audio: torch.Tensor, params: torch.Tensor, wa: torch.Tensor, wp: torch.Tensor,
                    row_scale: torch.Tensor | None = None, after_analysis=None):
    """Device path of ``bridge(gram(audio), params)`` with the pooling fused into the PQMF analysis kernel
    (``PQMF.analysis_pooled``): the bands make one trip to HBM.  ``row_scale`` = ``Voice(normalize="defer").row_scale``
    folds normalize_if_clipping into the analysis (the filter bank is linear).  ``after_analysis()`` is called once
    the analysis kernel is enqueued (bench.py records an event there).  -> (bands, x, y)."""
    from ias_b200 import IasError, _lib

    global LAST_BRIDGE_PATH
    x3 = audio.unsqueeze(1) if audio.dim() == 2 else audio
    try:
        bands, feat = gram.analysis_pooled(x3, EMBED_DIM, row_scale=row_scale)
    except IasError as exc:
        # only "this shape has no fused kernel" (bins narrower than a CTA tile: short clips) may take the unfused
        # device kernels; a launch failure, a workspace error or a missing library must surface
        if exc.code != _lib.IAS_ERR_UNSUPPORTED:
            raise
        bands = gram.analysis(x3, row_scale=row_scale)
        if after_analysis is not None:
            after_analysis()
        x, y = bridge(bands, params, wa, wp)
        LAST_BRIDGE_PATH = "unfused: ias_pqmf_analysis + ias_abs_avg_pool"
        return bands, x, y
    LAST_BRIDGE_PATH = "fused: ias_pqmf_analysis_pooled"
    if after_analysis is not None:
        after_analysis()
    return bands, feat @ wa, params @ wp


def oracle_front_end(batch_idx: int, B: int, N: int = 3, seconds: float = 4.0, cfg_batch: int | None = None,
                     reproducible: bool = True, timings: Dict[str, float] | None = None, torch_ops: bool = False):
    """CPU oracle of one step: returns dict(audio, params, bands, x, y, loss4).  fp32 throughout, like config 1.
    ``torch_ops`` runs PQMF and the loss through torch's CPU kernels (what the reference calls) instead of the float64
    numpy restatements -- the timing path of the CPU baseline."""
    from oracle import pqmf as OP
    from oracle import vicreg as OV
    from oracle import voice as OVc

    # (a partial batch of a reproducible render -- B % 32 != 0 -- still uses the 32-row noise table, row b % 32)
    cfg = OVc.SynthConfigO(batch_size=B, buffer_size_seconds=seconds, reproducible=reproducible and B % 32 == 0)
    t0 = time.perf_counter()
    u = OVc.seeded_params(batch_idx, B)
    t1 = time.perf_counter()
    noise = _noise_cache(32 if reproducible else B, cfg.buffer_size)
    out = OVc.voice_render(u, noise, cfg.buffer_size, cfg.control_buffer_size, cfg.sample_rate, cfg.control_rate,
                           cfg.eps)
    audio = out["audio"]
    t2 = time.perf_counter()
    H, _ = _filters(N)
    # the reference runs F.conv1d here; the oracle restatement is numpy (float64 accumulate)
    bands = OP.analysis_conv1d(audio, H, N) if torch_ops else torch.from_numpy(OP.analysis(audio.numpy(), H, N))
    t3 = time.perf_counter()
    params = OVc.sorted_to_registration(u)
    wa, wp = bridge_weights()
    x, y = bridge(bands, params, wa, wp)
    t4 = time.perf_counter()
    cb = cfg_batch if cfg_batch is not None else B
    loss4 = OV.loss_torch(x, y, cb, EMBED_DIM) if torch_ops else OV.loss(x.numpy(), y.numpy(), cb, EMBED_DIM,
                                                                        dtype=np.float32)
    t5 = time.perf_counter()
    if timings is not None:
        timings.update(seed=t1 - t0, synth=t2 - t1, pqmf=t3 - t2, bridge=t4 - t3, vicreg=t5 - t4)
    return dict(audio=audio, params=params, bands=bands, x=x, y=y, loss4=loss4, peak=out["peak"])


_NOISE: Dict[Tuple[int, int], torch.Tensor] = {}
_FILT: Dict[int, Tuple[np.ndarray, np.ndarray]] = {}


def _noise_cache(rows: int, T: int) -> torch.Tensor:
    from oracle import voice as OVc

    if (rows, T) not in _NOISE:
        _NOISE[(rows, T)] = OVc.noise_table(rows, T)
    return _NOISE[(rows, T)]


def _filters(N: int):
    from oracle import pqmf as OP

    if N not in _FILT:
        _FILT[N] = OP.design(N)
    return _FILT[N]
