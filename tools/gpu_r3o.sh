#!/bin/bash
# 2-D grids in the PQMF kernels: tests, bench, configs 3
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r3o}
timeout 900 python -m pytest tests/test_gpu_pqmf.py tests/test_gpu_e2e.py -m gpu -q -x > gpurun_out/test_pqmf_$TAG.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/test_pqmf_$TAG.log
timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-noise-variant > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$TAG.json"))
    print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "parity", d.get("parity_ok"))
    print({k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python tools/bench_configs.py --skip-long > gpurun_out/configs_$TAG.jsonl 2> gpurun_out/configs_$TAG.err; echo "configs exit $?"
cut -c1-220 gpurun_out/configs_$TAG.jsonl | grep -i "pqmf\|analysis\|config" | head -12
