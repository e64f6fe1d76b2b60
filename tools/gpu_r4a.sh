#!/bin/bash
# Packed N=3 synthesis: parity (bit-identity against the scalar kernel, goldens) and timing.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r4a}
timeout 600 python -m pytest tests/test_gpu_pqmf.py -m gpu -q -x > gpurun_out/test_pqmf_$TAG.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/test_pqmf_$TAG.log
timeout 300 python tools/sweep_pqmf_synth_n3.py > gpurun_out/sweep_synth_$TAG.log 2>&1; echo "sweep exit $?"; cat gpurun_out/sweep_synth_$TAG.log
