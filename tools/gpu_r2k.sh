#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pqmf.py tests/test_gpu_e2e.py -m gpu -q > gpurun_out/test_pqmf_r2k.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/test_pqmf_r2k.log
for v in "" nobulk; do
  if [ -z "$v" ]; then unset IAS_B200_LIB; else export IAS_B200_LIB=$PWD/inverse-audio-synthesis_b200/ias_b200/variants/libias_$v.so; fi
  timeout 600 python bench.py --no-cpu-baseline --no-noise-variant --steps 50 > gpurun_out/bench_r2k_${v:-bulk}.json 2> gpurun_out/bench_r2k_${v:-bulk}.err; echo "bench ${v:-bulk} exit $?"
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2k_${v:-bulk}.json"))
print("${v:-bulk}", "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), {k: round(x["ms_per_launch"], 4) for k, x in d["kernels"].items() if "pqmf" in k}, d["parity_ok"])
PY
  timeout 300 python tools/bench_configs.py --skip-long 2>/dev/null | grep "N=3" | grep "analysis 1024" | cut -c1-170
done
