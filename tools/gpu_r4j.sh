#!/bin/bash
# A/B of the synthesis kernels of every band count: committed build (head) against the working tree.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r4j}
V=$PWD/inverse-audio-synthesis_b200/ias_b200/variants
for v in head default head default; do
  if [ "$v" = "default" ]; then unset IAS_B200_LIB; else export IAS_B200_LIB=$V/libias_$v.so; fi
  echo "== $v" >> gpurun_out/sweep_synth_all_$TAG.log
  timeout 200 python tools/sweep_pqmf_synth_all.py 2>&1 | head -5 >> gpurun_out/sweep_synth_all_$TAG.log; echo "sweep $v exit $?"
done
unset IAS_B200_LIB
cat gpurun_out/sweep_synth_all_$TAG.log
