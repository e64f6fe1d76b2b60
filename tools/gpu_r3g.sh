#!/bin/bash
# ncu --set full capture of one audio-kernel shape: bash tools/gpu_r3g.sh <tag> <shape>
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r3g}; SHAPE=${2:-q128x16x4}
NCU="ncu --set full --clock-control none --import-source on"
IAS_VOICE_SHAPE=$SHAPE timeout 600 $NCU -k regex:k_voice_audio -s 3 -c 1 -o gpurun_out/prof_voice_audio_$TAG python tools/sweep_voice.py --non-reproducible --iters 2 $SHAPE > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu exit $?"
tail -2 gpurun_out/ncu_$TAG.log
