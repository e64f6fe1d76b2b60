#!/bin/bash
# Quick GPU check during development: voice + pqmf parity, audio-kernel shape sweep, short bench.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_quick.sh <tag>'
TAG=${1:-q}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for f in voice pqmf; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q -s -x > gpurun_out/test_${f}_$TAG.log 2>&1
  echo "test_$f exit $?"; tail -5 gpurun_out/test_${f}_$TAG.log
done
timeout 600 python tools/sweep_voice.py > gpurun_out/sweep_$TAG.log 2>&1; echo "sweep exit $?"; cat gpurun_out/sweep_$TAG.log | tail -12
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$TAG.json"))
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print({k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_$TAG.err
