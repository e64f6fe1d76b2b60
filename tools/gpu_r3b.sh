#!/bin/bash
# ncu --set full capture of the software-pipelined audio kernel (one launch)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=r3b
NCU="ncu --set full --clock-control none --import-source on"
IAS_VOICE_SHAPE=p128x16x4 timeout 600 $NCU -k regex:k_voice_audio -s 3 -c 1 -o gpurun_out/prof_voice_audio_sp_$TAG python tools/sweep_voice.py --non-reproducible --iters 2 p128x16x4 > gpurun_out/ncu_sp_$TAG.log 2>&1; echo "ncu exit $?"
tail -3 gpurun_out/ncu_sp_$TAG.log
