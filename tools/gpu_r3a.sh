#!/bin/bash
# Round-2 session 2, run a: software-pipelined audio kernel (k_voice_audio_sp) against the classic one.
# Shapes "pNTxSPTxCTAS" select the pipelined kernel; variant libraries change one macro each.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=r3a
V=$PWD/inverse-audio-synthesis_b200/ias_b200/variants
timeout 300 python tools/sweep_voice.py --non-reproducible --iters 20 128x16x4 p128x16x4 p128x16x3 p128x8x6 p128x8x8 > gpurun_out/sweep_$TAG.log 2>&1; echo "sweep exit $?"
cut -c1-260 gpurun_out/sweep_$TAG.log | tail -8
for v in spf2f sprs sppf; do
  IAS_B200_LIB=$V/libias_$v.so timeout 300 python tools/sweep_voice.py --non-reproducible --iters 20 128x16x4 p128x16x4 p128x16x3 > gpurun_out/sweep_${TAG}_$v.log 2>&1; echo "sweep $v exit $?"
  cut -c1-260 gpurun_out/sweep_${TAG}_$v.log | tail -3
done
IAS_VOICE_SHAPE=p128x16x4 timeout 900 python -m pytest tests/test_gpu_voice.py -m gpu -q -x > gpurun_out/test_voice_sp_$TAG.log 2>&1; echo "voice tests (sp) exit $?"; tail -4 gpurun_out/test_voice_sp_$TAG.log
IAS_VOICE_SHAPE=p128x16x4 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-noise-variant > gpurun_out/bench_sp_$TAG.json 2> gpurun_out/bench_sp_$TAG.err; echo "bench sp exit $?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_sp_$TAG.json"))
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d.get("parity_ok"))
    print({k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_sp_$TAG.err
