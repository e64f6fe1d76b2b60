#!/bin/bash
# Round 2, session c: pooled epilogue rewrite, workspace fix, pipeline default; config 5 at N=1.
TAG=${1:-r2c}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for f in vicreg pqmf e2e; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q -s > gpurun_out/test_${f}_$TAG.log 2>&1
  echo "test_$f exit $?"; tail -4 gpurun_out/test_${f}_$TAG.log
done
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; tail -3 gpurun_out/bench_$TAG.err
timeout 900 python bench.py --no-cpu-baseline --no-pipeline --no-parity --no-nonreproducible > gpurun_out/bench_nopipe_$TAG.json 2> gpurun_out/bench_nopipe_$TAG.err; echo "bench nopipe exit $?"
timeout 900 python bench.py --seconds 30 --batch-per-gpu 512 --steps 20 --warmup 3 --no-nonreproducible > gpurun_out/bench_c5_g1_$TAG.json 2> gpurun_out/bench_c5_g1_$TAG.err; echo "bench c5 exit $?"; tail -3 gpurun_out/bench_c5_g1_$TAG.err
python - <<PY
import json
for f in ("bench_$TAG", "bench_nopipe_$TAG", "bench_c5_g1_$TAG"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "nr", d.get("e2e_nonreproducible") and round(d["e2e_nonreproducible"]["value"]))
        print({k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
        print("parity", json.dumps(d.get("parity")), d.get("parity_ok"))
        print("cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(f, "parse failed", e)
PY
