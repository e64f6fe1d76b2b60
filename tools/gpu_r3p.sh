#!/bin/bash
# early (pinned) loads in the pipelined audio kernel: A/B against the shipped build
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=r3p
V=$PWD/inverse-audio-synthesis_b200/ias_b200/variants
for v in "" spel; do
  if [ -z "$v" ]; then unset IAS_B200_LIB; else export IAS_B200_LIB=$V/libias_$v.so; fi
  timeout 300 python tools/sweep_voice.py --non-reproducible --no-normalize --iters 30 128x16x4 p128x16x4 >> gpurun_out/sweep_$TAG.log 2>&1
  IAS_VOICE_RENDER_ALL=1 timeout 300 python tools/sweep_voice.py --batch 3552 --iters 5 p128x16x4 >> gpurun_out/sweep_$TAG.log 2>&1
done
python - <<PY
import json
for l in open("gpurun_out/sweep_$TAG.log"):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d["shape"], d["kernels_ms"].get("k_voice_audio"), d["max_abs_diff_vs_first"], d["finite"])
PY
