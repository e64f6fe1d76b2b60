#!/bin/bash
# Round 2, session e (1 GPU): publish-kernel rewrite under the one-GPU emulation, configs 2/3, ncu launch list + full captures.
TAG=${1:-r2e}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_pqmf.py -m gpu -q > gpurun_out/test_multi_pqmf_$TAG.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/test_multi_pqmf_$TAG.log
timeout 600 python tools/bench_configs.py --skip-long > gpurun_out/configs_$TAG.jsonl 2> gpurun_out/configs_$TAG.err; echo "configs exit $?"; cut -c1-260 gpurun_out/configs_$TAG.jsonl
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; tail -2 gpurun_out/bench_$TAG.err
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity --no-nonreproducible --no-pipeline"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launch list exit $?"
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:k_voice_audio -s 4 -c 1 -o gpurun_out/prof_voice_audio_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu voice_audio exit $?"
timeout 600 $NCU -k regex:k_pqmf_analysis -s 4 -c 1 -o gpurun_out/prof_k_pqmf_analysis_$TAG $CMD > gpurun_out/ncu_full_pq_$TAG.log 2>&1; echo "ncu pqmf_analysis exit $?"
timeout 300 $NCU -k regex:k_pqmf_synthesis -s 2 -c 1 -o gpurun_out/prof_k_pqmf_synthesis_n3_$TAG python tools/prof_pqmf.py > gpurun_out/ncu_syn3_$TAG.log 2>&1; echo "ncu synthesis N=3 exit $?"
for K in k_voice_control k_voice_adsr; do
  timeout 600 $NCU -k regex:$K -s 4 -c 1 -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_full_${K}_$TAG.log 2>&1; echo "ncu $K exit $?"
done
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "nr", d.get("e2e_nonreproducible") and round(d["e2e_nonreproducible"]["value"]))
print({k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
print("roofline", json.dumps(d["roofline"]))
print("parity_ok", d["parity_ok"], "cpu", d.get("cpu_baseline", {}).get("value"))
PY
