#!/bin/bash
# two-deep control pipeline: e2e tests + bench depth 1 vs 2
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=r3k
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_voice.py -m gpu -q -x > gpurun_out/test_e2e_$TAG.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/test_e2e_$TAG.log
for d in 1 2; do
  timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-noise-variant --pipeline-depth $d > gpurun_out/bench_${TAG}_d$d.json 2> gpurun_out/bench_${TAG}_d$d.err; echo "bench depth $d exit $?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_d$d.json"))
    print("depth $d value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "parity", d.get("parity_ok"), d["loss4_last_step"], d["e2e"]["loss4_last_step"])
except Exception as e:
    print("bench parse failed", e)
PY
  tail -2 gpurun_out/bench_${TAG}_d$d.err
done
