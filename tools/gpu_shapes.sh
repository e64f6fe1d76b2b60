#!/bin/bash
# Diagnostic: k_voice_audio timing for (variant library, shape) pairs.  usage: gpu_shapes.sh "<variant>:<shape> ..."
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for vs in $1; do
  v=${vs%%:*}; sh=${vs##*:}
  if [ "$v" = "default" ]; then unset IAS_B200_LIB; else export IAS_B200_LIB=$PWD/inverse-audio-synthesis_b200/ias_b200/variants/libias_$v.so; fi
  echo -n "$v $sh: "; timeout 300 python tools/sweep_voice.py --iters 30 $sh 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['kernels_ms'].get('k_voice_audio'), d['finite'], d['absmax'])"
done 2>&1 | tee -a gpurun_out/shapes.log
