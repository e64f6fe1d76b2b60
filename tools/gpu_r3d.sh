#!/bin/bash
# per-SM throughput of the audio-kernel shapes without the load-balance confound: every voice rendered in full
# (IAS_VOICE_RENDER_ALL), B = 3552 = 6 x 592 = 4 x 888 = 3 x 1184 voices
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=r3d
IAS_VOICE_RENDER_ALL=1 timeout 600 python tools/sweep_voice.py --batch 3552 --iters 8 128x16x4 p128x16x4 s128x16x4 s128x16x6 s128x16x8 > gpurun_out/sweep_$TAG.log 2>&1; echo "sweep exit $?"
IAS_VOICE_RENDER_ALL=1 timeout 600 python tools/sweep_voice.py --batch 2960 --iters 8 s128x16x5 >> gpurun_out/sweep_$TAG.log 2>&1; echo "sweep exit $?"
timeout 600 python tools/sweep_voice.py --batch 3552 --iters 8 128x16x4 p128x16x4 s128x16x4 s128x16x6 s128x16x8 >> gpurun_out/sweep_$TAG.log 2>&1; echo "sweep exit $?"
python - <<PY
import json
for l in open("gpurun_out/sweep_$TAG.log"):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d["shape"], d["kernels_ms"].get("k_voice_audio"), d["max_abs_diff_vs_first"], d["finite"])
PY
