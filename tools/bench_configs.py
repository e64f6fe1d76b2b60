#!/usr/bin/env python
"""Device-timed numbers for the BASELINE.json configs that bench.py's headline line does not cover (GPU, one device):

  config 2  Voice render only, 1024 x 4 s: reproducible (32-row noise table) and not (noise [B,T] from HBM)
  config 3  PQMF analysis + inverse on 1024 x 4 s voices, N = 16 (cutoff 0.15 and 0.03) and N = 3, with the
            reconstruction error next to the value the reference's own filters give
  config 5  long clips: 512 x 30 s through synth -> PQMF(3) -> bridge -> VICReg

One JSON line per measurement: ms per call (CUDA events around `--iters` calls after 3 warm-ups), sounds/s and the
achieved algorithmic GB/s against MEASURED_PEAKS.json.  python tools/bench_configs.py [--iters 10] [--skip-long]
"""
import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import harness  # noqa: E402
import ias_b200  # noqa: E402


def timed(fn, iters):
    out = None
    for _ in range(3):
        out = fn()  # keep the previous result alive, as the timed loop does: the allocator then owns two buffers
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--skip-long", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0

    def emit(config, what, ms, sounds, bytes_per_sound, **extra):
        gbs = sounds * bytes_per_sound / (ms * 1e-3) / 1e9
        print(json.dumps(dict(config=config, what=what, ms=round(ms, 4), sounds_per_s=round(sounds / (ms * 1e-3)),
                              algorithmic_GBps=round(gbs, 1), frac_of_hbm_peak=round(gbs / peak, 4), hbm_peak_GBps=peak,
                              **extra)), flush=True)

    B, T = 1024, 176400
    # ---- config 2 ----
    audio = None
    for reproducible in (True, False):
        cfg = ias_b200.SynthConfig(batch_size=B, reproducible=reproducible, sample_rate=44100, buffer_size_seconds=4.0)
        voice = ias_b200.Voice(synthconfig=cfg).to(dev)
        state = {"i": 0}

        def render():
            state["i"] += 1
            return voice(state["i"])[0]

        ms, audio = timed(render, args.iters)
        emit(2, "Voice(batch_idx) 1024 x 4 s, reproducible=%s (noise %s)" % (
            reproducible, "32-row table, L2 resident" if reproducible else "[B,T] read from HBM"), ms, B,
            4 * T * (1 if reproducible else 2))
        del voice
    # ---- config 3 ----
    x = audio.unsqueeze(1)
    for N, cutoff in ((16, 0.15), (16, 0.03), (3, 0.15)):
        m = ias_b200.PQMF(N=N, cutoff=cutoff).to(dev)
        ms_a, z = timed(lambda: m.analysis(x), args.iters)
        ms_s, y = timed(lambda: m.synthesis(z), args.iters)
        lag = 1  # the reference's synthesis output trails the input by one sample
        num = torch.sqrt(torch.mean((y[:, 0, lag:T] - x[:, 0, :T - lag]) ** 2))
        rel = float(num / torch.sqrt(torch.mean(x ** 2)))
        emit(3, f"PQMF(N={N}, cutoff={cutoff}).analysis 1024 x 4 s", ms_a, B, 8 * T)
        emit(3, f"PQMF(N={N}, cutoff={cutoff}).synthesis", ms_s, B, 8 * T)
        emit(3, f"PQMF(N={N}, cutoff={cutoff}) analysis + inverse", ms_a + ms_s, B, 16 * T,
             reconstruction_rel_rms_lag1=round(rel, 4),
             note="error of the reference's own filter design on synth audio; parity target is the reference's "
                  "output, not perfect reconstruction")
        del m, z, y
    del x, audio
    torch.cuda.empty_cache()
    # ---- config 5 ----
    if not args.skip_long:
        B5, sec = 512, 30.0
        T5 = int(sec * 44100)
        cfg = ias_b200.SynthConfig(batch_size=B5, reproducible=True, sample_rate=44100, buffer_size_seconds=sec)
        voice = ias_b200.Voice(synthconfig=cfg).to(dev)
        gram = ias_b200.PQMF(N=3).to(dev)
        vcfg = types.SimpleNamespace(dim=256, embeddim=256, vicreg=types.SimpleNamespace(
            mlp="8-8-%d", batch_size=B5, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0))
        vic = ias_b200.VICReg(vcfg, torch.nn.Identity(), torch.nn.Identity())
        wa, wp = harness.bridge_weights(dev)
        state = {"i": 0}

        def step():
            state["i"] += 1
            a, p, _ = voice(state["i"])
            bands = gram(a.unsqueeze(1))
            xe, ye = harness.bridge(bands, p, wa, wp)
            with torch.no_grad():
                return torch.stack(vic.loss(xe, ye))

        ms, out = timed(step, max(3, args.iters // 2))
        emit(5, "512 x 30 s: synth -> PQMF(3) -> bridge -> VICReg, eager launches", ms, B5, 12 * T5 + 8 * 256,
             loss4=[float(v) for v in out])
        ms_v, _ = timed(lambda: voice.output(), max(3, args.iters // 2))
        emit(5, "512 x 30 s: Voice.output() alone", ms_v, B5, 4 * T5)


if __name__ == "__main__":
    main()
