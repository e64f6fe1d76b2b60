#!/bin/bash
# Round 2, session b: stats-exchange emulation, new tests, audio-kernel variants, config 5 at N=1.
TAG=${1:-r2b}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for f in multi vicreg voice pqmf; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q -s > gpurun_out/test_${f}_$TAG.log 2>&1
  echo "test_$f exit $?"; tail -4 gpurun_out/test_${f}_$TAG.log
done
grep -h "voices\|vs fixture\|PASS\|FAIL" gpurun_out/test_voice_$TAG.log gpurun_out/test_multi_$TAG.log | head -40
echo "== audio kernel variants (128x16x4)"
for v in "" norad pkfold; do
  if [ -z "$v" ]; then unset IAS_B200_LIB; else export IAS_B200_LIB=$PWD/inverse-audio-synthesis_b200/ias_b200/variants/libias_$v.so; fi
  echo "variant ${v:-default}"; timeout 300 python tools/sweep_voice.py --iters 30 128x16x4 2>&1 | tail -1
done
unset IAS_B200_LIB
echo "== config 5 (512 x 30 s) at N=1"
timeout 900 python bench.py --seconds 30 --batch-per-gpu 512 --steps 20 --warmup 3 --no-nonreproducible > gpurun_out/bench_c5_g1_$TAG.json 2> gpurun_out/bench_c5_g1_$TAG.err; echo "bench c5 exit $?"; tail -3 gpurun_out/bench_c5_g1_$TAG.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_c5_g1_$TAG.json"))
    print("c5 value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d["value_mode"])
    print({k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
    print("parity", json.dumps(d.get("parity")), d.get("parity_ok"))
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("parse failed", e)
PY
