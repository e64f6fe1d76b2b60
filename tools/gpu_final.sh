#!/bin/bash
# Short closing session: smoke, GPU parity tests, bench (both arms), ncu launch list.  No full captures.
# Usage: gpurun --timeout 900 -- 'bash tools/gpu_final.sh <tag>'
TAG=${1:-fin}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/test_all_$TAG.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/test_all_$TAG.log
timeout 400 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref exit $?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launch list exit $?"
