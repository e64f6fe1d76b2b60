#!/bin/bash
# stats exchange after the one-fence-per-CTA change: emulation test (1 GPU of the box), then N ranks
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/test_multi_r2i.log 2>&1; echo "test_multi exit $?"; tail -2 gpurun_out/test_multi_r2i.log
for G in 1 $N; do
  if [ $G -eq 1 ]; then timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-noise-variant > gpurun_out/bench_g1_r2i.json 2> gpurun_out/bench_g1_r2i.err
  else NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --no-cpu-baseline --no-noise-variant > gpurun_out/bench_g${G}_r2i.json 2> gpurun_out/bench_g${G}_r2i.err; fi
  echo "bench g$G exit $?"; python - <<PY
import json
for l in open('gpurun_out/bench_g${G}_r2i.json'):
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l)
        print('n_gpus', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), 'parity_ok', d.get('parity_ok'))
        print('  ', {k: round(v['ms_per_launch'], 4) for k, v in d['kernels'].items() if 'vicreg' in k})
PY
done
