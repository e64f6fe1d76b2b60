#!/bin/bash
# pooled epilogue A/B: fixed-point redux (default) vs shared-memory tree (variant build)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pqmf.py tests/test_gpu_e2e.py -m gpu -q > gpurun_out/test_pool.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/test_pool.log
for v in "" poolold; do
  if [ -z "$v" ]; then unset IAS_B200_LIB; else export IAS_B200_LIB=$PWD/inverse-audio-synthesis_b200/ias_b200/variants/libias_$v.so; fi
  timeout 600 python bench.py --no-cpu-baseline --no-noise-variant --steps 50 > gpurun_out/bench_pool_${v:-new}.json 2> gpurun_out/bench_pool_${v:-new}.err; echo "bench ${v:-new} exit $?"
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_pool_${v:-new}.json"))
print("${v:-new}", "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), {k: round(x["ms_per_launch"], 4) for k, x in d["kernels"].items() if "pqmf" in k or "pool" in k}, d["parity_ok"], d["parity"]["oracle_batch"]["loss4_rel_end_to_end"])
PY
done
