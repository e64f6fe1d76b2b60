#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=r3j
for f in "" "--main-priority" "--no-pipeline"; do
  timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-noise-variant --no-parity $f > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench [$f] exit $?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$TAG.json"))
    print("[$f] value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("bench parse failed", e)
PY
  tail -2 gpurun_out/bench_$TAG.err
done
