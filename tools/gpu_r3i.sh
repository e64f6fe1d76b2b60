#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=r3i
for g in 4 3 2; do
  for m in pipe deep; do
    IAS_VOICE_GRID_PER_SM=$g timeout 300 python tools/exp_overlap.py --mode $m 2>&1 | tail -1 | tee -a gpurun_out/overlap_$TAG.log
  done
done
IAS_VOICE_GRID_PER_SM=3 timeout 300 python tools/exp_overlap.py --mode deep --priority -1 2>&1 | tail -1 | tee -a gpurun_out/overlap_$TAG.log
