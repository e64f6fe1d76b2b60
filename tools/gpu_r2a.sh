#!/bin/bash
# Round 2, session a: full GPU test suite on the new build, bench with the parity block.
TAG=${1:-r2a}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_$TAG.log
for f in multi vicreg pqmf voice e2e; do
  timeout 1500 python -m pytest tests/test_gpu_$f.py -m gpu -q -s > gpurun_out/test_${f}_$TAG.log 2>&1
  echo "test_$f exit $?"; tail -6 gpurun_out/test_${f}_$TAG.log
done
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; tail -3 gpurun_out/bench_$TAG.err
timeout 900 python bench.py --pipeline --no-cpu-baseline --no-parity > gpurun_out/bench_pipe_$TAG.json 2> gpurun_out/bench_pipe_$TAG.err; echo "bench pipe exit $?"
python - <<PY
import json
for f in ("bench_$TAG", "bench_pipe_$TAG"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "nr", d.get("e2e_nonreproducible") and round(d["e2e_nonreproducible"]["value"]))
        print({k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
        print("parity", json.dumps(d.get("parity")), d.get("parity_ok"))
        print("clocks", d.get("clocks"))
    except Exception as e:
        print(f, "parse failed", e)
PY
