#!/bin/bash
# Pointer-walk loads in the cosine-modulated synthesis kernels: PQMF parity tests + timing of every band count.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-r4i}
timeout 600 python -m pytest tests/test_gpu_pqmf.py -m gpu -q -x > gpurun_out/test_pqmf_$TAG.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/test_pqmf_$TAG.log
timeout 300 python tools/sweep_pqmf_synth_all.py > gpurun_out/sweep_synth_all_$TAG.log 2>&1; echo "sweep exit $?"; cat gpurun_out/sweep_synth_all_$TAG.log
