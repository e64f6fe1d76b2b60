#!/bin/bash
# Quick GPU check during development: chosen parity test files, optional shape sweep, short bench.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_quick.sh <tag> "<test files: voice pqmf vicreg e2e>" [sweep]'
TAG=${1:-q}
TESTS=${2:-"voice pqmf"}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for f in $TESTS; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q -s -x > gpurun_out/test_${f}_$TAG.log 2>&1
  echo "test_$f exit $?"; tail -5 gpurun_out/test_${f}_$TAG.log
done
if [ -n "$3" ]; then
  timeout 600 python tools/sweep_voice.py > gpurun_out/sweep_$TAG.log 2>&1; echo "sweep exit $?"; tail -12 gpurun_out/sweep_$TAG.log
fi
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$TAG.json"))
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("mode"), "host-params", d["e2e_host_params"]["value"])
    print({k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
    print("loss", d["loss4_last_step"], d["e2e"].get("loss4_last_step"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_$TAG.err
