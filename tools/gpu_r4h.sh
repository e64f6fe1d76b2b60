#!/bin/bash
# Pooled analysis epilogue with per-band accumulator addresses from shared memory: parity tests + A/B against HEAD.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r4h}
V=$PWD/inverse-audio-synthesis_b200/ias_b200/variants
timeout 600 python -m pytest tests/test_gpu_pqmf.py tests/test_gpu_e2e.py -m gpu -q -x > gpurun_out/test_pqmf_$TAG.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/test_pqmf_$TAG.log
for rep in 1 2; do
for v in head default; do
  if [ "$v" = "default" ]; then unset IAS_B200_LIB; else export IAS_B200_LIB=$V/libias_$v.so; fi
  echo "== $v" >> gpurun_out/sweep_pooled_$TAG.log
  timeout 200 python tools/sweep_pqmf_pooled.py >> gpurun_out/sweep_pooled_$TAG.log 2>&1; echo "sweep $v exit $?"
done
done
unset IAS_B200_LIB
cat gpurun_out/sweep_pooled_$TAG.log
