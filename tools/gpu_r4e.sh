#!/bin/bash
# A/B of the packed N=3 synthesis kernel in ONE session (boxes of the pool differ by ~1 %): the committed r4b build
# (static shared memory, r + r/Q padding in both planes, no register cap), the current default, and one-change variants.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r4e}
V=$PWD/inverse-audio-synthesis_b200/ias_b200/variants
for rep in 1 2; do
for v in head default pad0 minb1 static static_minb1; do
  if [ "$v" = "default" ]; then unset IAS_B200_LIB; else export IAS_B200_LIB=$V/libias_$v.so; fi
  [ "$v" != "default" ] && [ ! -f "$IAS_B200_LIB" ] && continue
  echo "== $v" >> gpurun_out/sweep_synth_$TAG.log
  IAS_SWEEP_ONLY=8 timeout 200 python tools/sweep_pqmf_synth_n3.py >> gpurun_out/sweep_synth_$TAG.log 2>&1; echo "sweep $v exit $?"
done
done
unset IAS_B200_LIB
cat gpurun_out/sweep_synth_$TAG.log
