#!/bin/bash
# multi-GPU check of the final build: gpurun --gpus N -- 'bash tools/gpu_r3m.sh N'
N=${1:-2}; TAG=r4w$N
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/test_multi_$TAG.log 2>&1; echo "test_multi exit $?"; tail -2 gpurun_out/test_multi_$TAG.log
for G in ${GS:-1 $N}; do
  if [ $G -eq 1 ]; then timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-noise-variant > gpurun_out/bench_g1_$TAG.json 2> gpurun_out/bench_g1_$TAG.err
  else NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --no-cpu-baseline --no-noise-variant > gpurun_out/bench_g${G}_$TAG.json 2> gpurun_out/bench_g${G}_$TAG.err; fi
  echo "bench g$G exit $?"; python - <<PY
import json
for l in open('gpurun_out/bench_g${G}_$TAG.json'):
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l)
        print('n_gpus', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), 'parity_ok', d.get('parity_ok'))
        print('  ', d['parity']['exchange'] if d.get('parity') else None)
PY
done
tail -3 gpurun_out/bench_g${N}_$TAG.err
