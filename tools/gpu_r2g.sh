#!/bin/bash
# Round 2, session g (1 GPU): full GPU suite + smoke + default bench (reference noise mode) + reference arm.
TAG=${1:-r2g}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/test_all_$TAG.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/test_all_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; tail -2 gpurun_out/bench_$TAG.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref exit $?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "variant", d.get("e2e_noise_variant") and round(d["e2e_noise_variant"]["value"]))
print({k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
print("roofline", json.dumps(d["roofline"]))
print("parity", json.dumps(d["parity"]))
print("parity_ok", d["parity_ok"], "cpu", d.get("cpu_baseline", {}).get("value"), d["config"]["noise"])
r = json.load(open("gpurun_out/bench_ref_$TAG.json"))
print("ref", r["value"], r["ms_per_step"], r["config"])
PY
