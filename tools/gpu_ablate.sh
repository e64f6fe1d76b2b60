#!/bin/bash
# Diagnostic: time k_voice_audio with parts of its tile loop ablated (variant builds; results are wrong on purpose).
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for v in "" norad abl_NOXU abl_NOPASS1 abl_NOF64 abl_NOBAR abl_XUF64 abl_ALL $EXTRA_VARIANTS; do
  if [ -z "$v" ]; then unset IAS_B200_LIB; else export IAS_B200_LIB=$PWD/inverse-audio-synthesis_b200/ias_b200/variants/libias_$v.so; fi
  [ -n "$v" ] && [ ! -f "$IAS_B200_LIB" ] && continue
  echo -n "variant ${v:-default}: "; timeout 300 python tools/sweep_voice.py --iters 30 128x16x4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['kernels_ms'].get('k_voice_audio'), d['finite'])"
done 2>&1 | tee gpurun_out/ablate.log
