#!/bin/bash
# Round 2, session h (8 GPUs): exchange routes across processes, headline config at N=8, config 5 (30 s x 512/GPU) at N=8 and N=4.
TAG=${1:-r2h}
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -k across_processes > gpurun_out/test_multi_$TAG.log 2>&1; echo "test_multi exit $?"; grep -E "PASS|FAIL|passed|failed|Error" gpurun_out/test_multi_$TAG.log | tail -20
run() {  # run <gpus> <name> <extra args...>
  local G=$1; local NAME=$2; shift 2
  NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G "$@" > gpurun_out/$NAME.json 2> gpurun_out/$NAME.err
  echo "$NAME exit $?"; python - <<PY
import json
for l in open('gpurun_out/$NAME.json'):
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l)
        print('$NAME', 'n_gpus', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), 'parity_ok', d.get('parity_ok'))
        print('  ', {k: round(v['ms_per_launch'], 4) for k, v in d['kernels'].items() if 'vicreg' in k})
        print('  ', json.dumps((d.get('parity') or {}).get('exchange')))
PY
  tail -2 gpurun_out/$NAME.err
}
run 8 bench_g8_stats_$TAG --no-noise-variant --no-cpu-baseline
run 8 bench_c5_g8_$TAG --seconds 30 --batch-per-gpu 512 --steps 20 --warmup 3 --no-noise-variant
run 4 bench_c5_g4_$TAG --seconds 30 --batch-per-gpu 512 --steps 20 --warmup 3 --no-noise-variant
run 8 bench_g8_nccl_$TAG --gather nccl --no-noise-variant --no-cpu-baseline --steps 50
