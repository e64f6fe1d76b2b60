#!/bin/bash
# Multi-GPU session: gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi.sh N tag'
N=${1:-2}; TAG=${2:-r01m}
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/multi_check_$TAG.log 2>&1; echo "multi check exit $?"; grep -E "PASS|FAIL|Error|error" gpurun_out/multi_check_$TAG.log | head
for G in 1 $N; do
  if [ $G -eq 1 ]; then timeout 300 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/bench_g1_$TAG.json 2> gpurun_out/bench_g1_$TAG.err
  else NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G > gpurun_out/bench_g${G}_$TAG.json 2> gpurun_out/bench_g${G}_$TAG.err; fi
  echo "bench g$G exit $?"; python -c "
import json,sys
for l in open('gpurun_out/bench_g${G}_$TAG.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('n_gpus',d['n_gpus'],'value',round(d['value']),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']))
"
done
tail -3 gpurun_out/bench_g${N}_$TAG.err
