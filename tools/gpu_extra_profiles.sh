#!/bin/bash
# ncu --set full captures of the kernels gpu_round.sh does not cover (one launch each):
#   k_voice_adsr, k_abs_avg_pool (front-end step), k_pqmf_synthesis N=3 / N=16, loss kernels at B=8192.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_extra_profiles.sh <tag>'; then python tools/summarize_ncu.py <tag>x
TAG=${1:-r02}x
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
STEP="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph"
for K in k_voice_adsr k_abs_avg_pool k_voice_schedule k_seed_params; do
  timeout 600 $NCU -k regex:$K -s 4 -c 1 -o gpurun_out/prof_${K}_$TAG $STEP > gpurun_out/ncu_${K}_$TAG.log 2>&1; echo "ncu $K exit $?"
done
timeout 600 $NCU -k regex:k_pqmf_synthesis -s 2 -c 1 -o gpurun_out/prof_k_pqmf_synthesis_n3_$TAG python tools/prof_pqmf.py > gpurun_out/ncu_syn3_$TAG.log 2>&1; echo "ncu synthesis N=3 exit $?"
timeout 600 $NCU -k regex:k_pqmf_synthesis -s 5 -c 1 -o gpurun_out/prof_k_pqmf_synthesis_n16_$TAG python tools/prof_pqmf.py > gpurun_out/ncu_syn16_$TAG.log 2>&1; echo "ncu synthesis N=16 exit $?"
for K in k_center_pack k_gram_tc k_finalize k_colsum_v4; do
  timeout 600 $NCU -k regex:$K -s 30 -c 1 -o gpurun_out/prof_${K}_b8192_$TAG python tools/time_vicreg.py > gpurun_out/ncu_${K}_b8192_$TAG.log 2>&1; echo "ncu $K B=8192 exit $?"
done
