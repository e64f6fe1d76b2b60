#!/bin/bash
# One GPU-box session: smoke, GPU parity tests (one process per file), bench, ncu launch list + one full capture.
# Usage: gpurun --timeout 1800 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi_$TAG.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_$TAG.log
for f in pqmf vicreg voice e2e; do
  timeout 1200 python -m pytest tests/test_gpu_$f.py -m gpu -q -s > gpurun_out/test_${f}_$TAG.log 2>&1
  echo "test_$f exit $?"; tail -4 gpurun_out/test_${f}_$TAG.log
done
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref exit $?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_voice_audio -s 4 -c 1 -o gpurun_out/prof_voice_audio_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
for K in k_pqmf_analysis k_gram_tc k_voice_control; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_full_${K}_$TAG.log 2>&1
  echo "ncu $K exit $?"
done
# BASELINE config 3 kernels (not on the bench.py step): PQMF synthesis N=3 / N=16, analysis N=16
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:k_pqmf_synthesis -s 2 -c 1 -o gpurun_out/prof_k_pqmf_synthesis_n3_$TAG python tools/prof_pqmf.py > gpurun_out/ncu_syn3_$TAG.log 2>&1; echo "ncu synthesis N=3 exit $?"
timeout 300 $NCU -k regex:k_pqmf_synthesis -s 5 -c 1 -o gpurun_out/prof_k_pqmf_synthesis_n16_$TAG python tools/prof_pqmf.py > gpurun_out/ncu_syn16_$TAG.log 2>&1; echo "ncu synthesis N=16 exit $?"
timeout 300 $NCU -k regex:k_pqmf_analysis -s 1 -c 1 -o gpurun_out/prof_k_pqmf_analysis_n16_$TAG python tools/prof_pqmf_analysis.py > gpurun_out/ncu_ana16_$TAG.log 2>&1; echo "ncu analysis N=16 exit $?"
ls -la gpurun_out | head -40
timeout 600 python tools/bench_configs.py > gpurun_out/configs_$TAG.jsonl 2> gpurun_out/configs_$TAG.err; echo "configs exit $?"
