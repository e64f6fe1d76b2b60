#!/bin/bash
# occupancy scaling of one binary: 1..4 resident CTAs per SM of the same 128-register kernels, every voice in full
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=r3e
for g in 1 2 3 4; do
  echo "grid per SM $g" >> gpurun_out/sweep_$TAG.log
  IAS_VOICE_GRID_PER_SM=$g IAS_VOICE_RENDER_ALL=1 timeout 600 python tools/sweep_voice.py --batch 3552 --iters 3 128x16x4 p128x16x4 s128x16x8 >> gpurun_out/sweep_$TAG.log 2>&1
done
python - <<PY
import json
for l in open("gpurun_out/sweep_$TAG.log"):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d["shape"], d["kernels_ms"].get("k_voice_audio"), d["max_abs_diff_vs_first"], d["finite"])
PY
