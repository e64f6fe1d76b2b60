#!/bin/bash
# k_voice_audio_sx (increments in shared memory) shape sweep against the classic and the pipelined kernel
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=r3c
timeout 600 python tools/sweep_voice.py --non-reproducible --iters 20 128x16x4 p128x16x4 s128x16x4 s128x16x5 s128x16x6 s128x16x8 s256x16x2 s256x16x3 s64x16x8 s64x16x12 > gpurun_out/sweep_$TAG.log 2>&1; echo "sweep exit $?"
python - <<PY
import json
for l in open("gpurun_out/sweep_$TAG.log"):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d["shape"], d["kernels_ms"].get("k_voice_audio"), d["max_abs_diff_vs_first"], d["finite"])
PY
