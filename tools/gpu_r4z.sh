#!/bin/bash
# Closing session of round 2 / session 3: smoke, all GPU parity tests, bench (both arms), ncu launch list, full captures
# of the three large kernels of the final sources (audio, pooled analysis, packed N=3 synthesis), BASELINE configs 2/3.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_r4z.sh <tag>'; then python tools/summarize_ncu.py <tag>
TAG=${1:-r4z}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi_$TAG.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/test_all_$TAG.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/test_all_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref exit $?"
python - <<PY
import json
for f in ("bench_$TAG.json", "bench_ref_$TAG.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "parity", d.get("parity_ok"))
        if "kernels" in d:
            print({k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "parse failed", e)
PY
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launch list exit $?"
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_voice_audio -s 4 -c 1 -o gpurun_out/prof_voice_audio_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu audio exit $?"
timeout 600 $NCU -k regex:k_pqmf_analysis -s 4 -c 1 -o gpurun_out/prof_k_pqmf_analysis_$TAG $CMD > gpurun_out/ncu_full_k_pqmf_analysis_$TAG.log 2>&1; echo "ncu analysis exit $?"
timeout 300 $NCU -k regex:k_pqmf_synthesis -s 2 -c 1 -o gpurun_out/prof_k_pqmf_synthesis_n3_$TAG python tools/prof_pqmf.py > gpurun_out/ncu_syn3_$TAG.log 2>&1; echo "ncu synthesis N=3 exit $?"
timeout 300 $NCU -k regex:k_pqmf_synthesis -s 5 -c 1 -o gpurun_out/prof_k_pqmf_synthesis_n16_$TAG python tools/prof_pqmf.py > gpurun_out/ncu_syn16_$TAG.log 2>&1; echo "ncu synthesis N=16 exit $?"
timeout 600 python tools/bench_configs.py --skip-long > gpurun_out/configs_$TAG.jsonl 2> gpurun_out/configs_$TAG.err; echo "configs exit $?"
cut -c1-260 gpurun_out/configs_$TAG.jsonl | head -12
