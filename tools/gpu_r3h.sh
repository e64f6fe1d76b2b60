#!/bin/bash
# pipelined audio kernel as the default + deferred normalisation: voice / e2e tests, bench with both --normalize modes
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r3h}
timeout 900 python -m pytest tests/test_gpu_voice.py tests/test_gpu_e2e.py -m gpu -q -x > gpurun_out/test_voice_$TAG.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/test_voice_$TAG.log
for m in defer kernel; do
  timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-noise-variant --normalize $m > gpurun_out/bench_${TAG}_$m.json 2> gpurun_out/bench_${TAG}_$m.err; echo "bench $m exit $?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_$m.json"))
    print("$m value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "parity", d.get("parity_ok"))
    print({k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
    print(d["parity"]["oracle_batch"]["voices_le_1e-4"], d["parity"]["oracle_batch"]["loss4_rel_end_to_end"], d["parity"]["oracle_batch"]["pqmf_rel"])
except Exception as e:
    print("bench parse failed", e)
PY
  tail -3 gpurun_out/bench_${TAG}_$m.err
done
