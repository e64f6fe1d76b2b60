#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_voice.py tests/test_gpu_e2e.py -m gpu -q > gpurun_out/test_voice_r2j.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/test_voice_r2j.log
timeout 600 python bench.py --no-cpu-baseline --no-noise-variant --steps 50 > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; echo "bench exit $?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2j.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), {k: round(x["ms_per_launch"], 4) for k, x in d["kernels"].items() if "voice" in k or "seed" in k}, d["parity_ok"])
PY
