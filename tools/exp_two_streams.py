#!/usr/bin/env python
"""Experiment (GPU): two independent front-end pipelines on two streams, steps alternating between them, against one
pipeline on one stream.  Measures how much of the under-utilised head (seed/ADSR/control) and tail (pool/loss) of a
step can be hidden behind the neighbouring step.  python tools/exp_two_streams.py [--steps 60]"""
import argparse
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import harness  # noqa: E402
import ias_b200  # noqa: E402


def make_pipeline(dev, B):
    cfg = ias_b200.SynthConfig(batch_size=B, reproducible=True, sample_rate=44100, buffer_size_seconds=4.0)
    voice = ias_b200.Voice(synthconfig=cfg).to(dev)
    gram = ias_b200.PQMF(N=3).to(dev)
    vcfg = types.SimpleNamespace(dim=256, embeddim=256, vicreg=types.SimpleNamespace(
        mlp="8-8-%d", batch_size=B, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0))
    vic = ias_b200.VICReg(vcfg, torch.nn.Identity(), torch.nn.Identity())
    wa, wp = harness.bridge_weights(dev)
    idx = torch.zeros(1, dtype=torch.int64, device=dev)

    def step():
        audio, params, _ = voice(idx)
        bands = gram(audio.unsqueeze(1))
        x, y = harness.bridge(bands, params, wa, wp)
        with torch.no_grad():
            out = torch.stack(vic.loss(x, y))
        idx.add_(2)
        return out

    return step, idx


def capture(step, stream):
    with torch.cuda.stream(stream):
        for _ in range(3):
            step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        out = step()
    return g, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--batch", type=int, default=1024)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    stepA, idxA = make_pipeline(dev, args.batch)
    stepB, idxB = make_pipeline(dev, args.batch)
    gA, outA = capture(stepA, sa)
    gB, outB = capture(stepB, sb)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def run(two):
        idxA.fill_(10)
        idxB.fill_(11)
        torch.cuda.synchronize()
        e0.record()
        if two:
            sa.wait_stream(torch.cuda.current_stream())
            sb.wait_stream(torch.cuda.current_stream())
            for i in range(args.steps // 2):
                with torch.cuda.stream(sa):
                    gA.replay()
                with torch.cuda.stream(sb):
                    gB.replay()
            torch.cuda.current_stream().wait_stream(sa)
            torch.cuda.current_stream().wait_stream(sb)
        else:
            sa.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(sa):
                for i in range(args.steps // 2):
                    gA.replay()
                    gB.replay()
            torch.cuda.current_stream().wait_stream(sa)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (args.steps // 2 * 2)

    for _ in range(2):
        one = run(False)
        two = run(True)
        print(f"one stream: {one:.4f} ms/step ({args.batch / one * 1e3:.0f} sounds/s)   two streams: {two:.4f} ms/step "
              f"({args.batch / two * 1e3:.0f} sounds/s)   gain {one / two:.3f}x", flush=True)
    print("last losses", outA.tolist(), outB.tolist())


if __name__ == "__main__":
    main()
