#!/usr/bin/env python
"""Experiment (GPU): render batch k+1 (k_voice_audio, bound on the instruction side, 30 % of HBM) WHILE the PQMF analysis,
bridge and loss of batch k (HBM bound) and the control stage of batch k+2 run on a second stream.  The audio kernel
must leave room on the SMs: IAS_VOICE_GRID_PER_SM=3 launches three persistent CTAs per SM (49 k of 64 k registers).

    IAS_VOICE_GRID_PER_SM=3 python tools/exp_overlap.py [--steps 40] [--mode deep|pipe]
"""
import argparse
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import harness  # noqa: E402
import ias_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--mode", default="deep", choices=["deep", "pipe"])
    ap.add_argument("--priority", type=int, default=0, help="priority of the side stream (-1 = high)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    B = args.batch
    cfg = ias_b200.SynthConfig(batch_size=B, reproducible=False, sample_rate=44100, buffer_size_seconds=4.0)
    voices = [ias_b200.Voice(synthconfig=cfg, normalize="defer").to(dev) for _ in range(2)]
    voices[1].noise.noise = voices[0].noise.noise  # one [B,T] table
    gram = ias_b200.PQMF(N=3).to(dev)
    vcfg = types.SimpleNamespace(dim=256, embeddim=256, vicreg=types.SimpleNamespace(
        mlp="8-8-%d", batch_size=B, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0))
    vic = ias_b200.VICReg(vcfg, torch.nn.Identity(), torch.nn.Identity())
    wa, wp = harness.bridge_weights(dev)
    idx = torch.zeros(1, dtype=torch.int64, device=dev)
    side = torch.cuda.Stream(device=dev, priority=args.priority)
    main_s = torch.cuda.current_stream()
    T = cfg.buffer_size
    audio = [torch.empty((B, T), device=dev) for _ in range(2)]
    params = [torch.empty((B, 78), device=dev) for _ in range(2)]
    scale = [torch.empty(B, device=dev) for _ in range(2)]
    losses = []

    def render(v, slot):
        voices[v]._prepared = None
        voices[v]._render(voices[v].STAGE_AUDIO, audio[slot])
        scale[slot].copy_(voices[v].row_scale)

    def prepare(v, slot):
        idx.add_(1)
        voices[v].prepare(idx)
        params[slot].copy_(voices[v]._store.t())

    def consume(slot):
        _, x, y = harness.analysis_bridge(gram, audio[slot], params[slot], wa, wp, scale[slot])
        with torch.no_grad():
            return torch.stack(vic.loss(x, y))

    def run(steps):
        # prologue: batch 0 rendered in slot 0 by voice 0, batch 1 prepared by voice 1
        idx.fill_(-1)
        prepare(0, 0)
        render(0, 0)
        prepare(1, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for k in range(steps):
            a, b = k & 1, (k + 1) & 1
            if args.mode == "deep":
                side.wait_stream(main_s)
                render(b, b)                      # audio of batch k+1 on the main stream
                with torch.cuda.stream(side):     # PQMF / bridge / loss of batch k, then the control stage of batch k+2
                    out = consume(a)
                    prepare(a, a)
                main_s.wait_stream(side)
            else:                                 # the bench's pipeline: audio(k+1); then consume(k+1) || prepare(k+2)
                out = consume(a)
                render(b, b)
                side.wait_stream(main_s)
                with torch.cuda.stream(side):
                    prepare(a, a)
                main_s.wait_stream(side)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, out

    run(6)
    ms, out = run(args.steps)
    print(f"mode {args.mode} grid/SM {os.environ.get('IAS_VOICE_GRID_PER_SM', 'default')} prio {args.priority}: "
          f"{ms:.4f} ms/step = {B / ms * 1e3:.0f} sounds/s  loss {[round(float(v), 5) for v in out]}", flush=True)


if __name__ == "__main__":
    main()
