#!/bin/bash
# Round 2, closing session (1 GPU): smoke, full GPU suite, bench (both arms), configs 2/3, ncu launch list + full captures
# of the final sources (profiles/capture_k_voice_audio.json is keyed by their hash).
TAG=${1:-r2z}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/test_all_$TAG.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/test_all_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; tail -2 gpurun_out/bench_$TAG.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref exit $?"
timeout 600 python tools/bench_configs.py --skip-long > gpurun_out/configs_$TAG.jsonl 2> gpurun_out/configs_$TAG.err; echo "configs exit $?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity --no-noise-variant --no-pipeline"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launch list exit $?"
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:k_voice_audio -s 4 -c 1 -o gpurun_out/prof_voice_audio_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu voice_audio exit $?"
timeout 600 $NCU -k regex:k_pqmf_analysis -s 4 -c 1 -o gpurun_out/prof_k_pqmf_analysis_$TAG $CMD > gpurun_out/ncu_full_pq_$TAG.log 2>&1; echo "ncu pqmf_analysis exit $?"
timeout 300 $NCU -k regex:k_pqmf_synthesis -s 2 -c 1 -o gpurun_out/prof_k_pqmf_synthesis_n3_$TAG python tools/prof_pqmf.py > gpurun_out/ncu_syn3_$TAG.log 2>&1; echo "ncu synthesis N=3 exit $?"
timeout 300 $NCU -k regex:k_gram_tc -s 4 -c 1 -o gpurun_out/prof_k_gram_tc_$TAG $CMD > gpurun_out/ncu_full_gram_$TAG.log 2>&1; echo "ncu gram exit $?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "variant", d.get("e2e_noise_variant") and round(d["e2e_noise_variant"]["value"]))
print({k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
print("roofline", json.dumps(d["roofline"]))
print("parity_ok", d["parity_ok"], "cpu", d.get("cpu_baseline", {}).get("value"))
PY
