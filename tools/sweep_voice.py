#!/usr/bin/env python
"""Tuning run (GPU): time ias_voice_render for several audio-kernel shapes (IAS_VOICE_SHAPE=<threads>x<samples per
thread>x<CTAs per SM>) on the bench workload and check that every shape produces the same audio bit for bit.

    python tools/sweep_voice.py [--batch 1024] [--seconds 4] [--iters 20] [shape ...]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import harness  # noqa: E402,F401  (sets sys.path)
import ias_b200  # noqa: E402
from ias_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-normalize", action="store_true", help="Voice(normalize=False): skip normalize_if_clipping")
    ap.add_argument("--non-reproducible", action="store_true", help="[B,T] noise table from HBM")
    ap.add_argument("shapes", nargs="*", default=["128x8x7", "128x8x6", "128x16x4", "128x16x3", "256x8x3", "256x16x2"])
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = ias_b200.lib()
    cfg = ias_b200.SynthConfig(batch_size=args.batch, reproducible=not args.non_reproducible, sample_rate=44100,
                               buffer_size_seconds=args.seconds)
    voice = ias_b200.Voice(synthconfig=cfg, normalize=not args.no_normalize).to(dev)
    voice.randomize(seed=7)
    ref = None
    for shape in args.shapes:
        os.environ["IAS_VOICE_SHAPE"] = shape
        for _ in range(3):
            audio = voice.output()
        torch.cuda.synchronize()
        lib.ias_prof_reset()
        lib.ias_prof_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            audio = voice.output()
        e1.record()
        torch.cuda.synchronize()
        lib.ias_prof_enable(0)
        kern = {}
        for k in range(lib.ias_prof_kernel_count()):
            tot, n = ctypes.c_double(0), ctypes.c_longlong(0)
            _lib.check(lib.ias_prof_read(k, ctypes.byref(tot), ctypes.byref(n)), "ias_prof_read")
            if n.value:
                kern[lib.ias_prof_kernel_name(k).decode()] = round(tot.value / n.value, 4)
        same = None
        if ref is None:
            ref = audio.clone()
        else:
            same = float((ref - audio).abs().max())  # amplitude path rounds differently per shape (<= ~2e-7)
        print(json.dumps({"shape": shape, "ms_per_render": round(e0.elapsed_time(e1) / args.iters, 4), "kernels_ms": kern,
                          "max_abs_diff_vs_first": same, "finite": bool(torch.isfinite(audio).all()),
                          "absmax": float(audio.abs().max())}), flush=True)


if __name__ == "__main__":
    main()
