#!/bin/bash
# k_voice_audio_s2 (two tiles of skew, scan under the arithmetic) against the other audio kernels
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r3f}
SHAPES=${2:-"128x16x4 p128x16x4 q128x16x4 q128x16x5 q128x16x3"}
rm -f gpurun_out/sweep_$TAG.log
echo "render_all B=3552" >> gpurun_out/sweep_$TAG.log
IAS_VOICE_RENDER_ALL=1 timeout 600 python tools/sweep_voice.py --batch 3552 --iters 5 $SHAPES >> gpurun_out/sweep_$TAG.log 2>&1
echo "bench workload B=1024 non-reproducible" >> gpurun_out/sweep_$TAG.log
timeout 600 python tools/sweep_voice.py --non-reproducible --iters 20 $SHAPES >> gpurun_out/sweep_$TAG.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/sweep_$TAG.log"):
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d["shape"], d["kernels_ms"].get("k_voice_audio"), d["max_abs_diff_vs_first"], d["finite"])
PY
