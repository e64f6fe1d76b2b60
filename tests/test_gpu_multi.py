"""GPU: the embedding exchange of the loss -- N ranks must reproduce the single-process
loss on the rank-ordered concatenation (the semantics of the reference's FullGatherLayer, vicreg.py:38-39,79-95),
forward and backward.

  * one GPU is enough for the statistics exchange: W emulated ranks (own buffers, workspaces and streams) run the real
    kernels concurrently on cuda:0 and wait on each other's flags (tools/stats_emulate.py, in a subprocess);
  * with >= 2 GPUs the same check runs across processes through every route (torch.distributed NCCL, libias_comm,
    fused peer-read gather, statistics exchange) plus the sharded seeding (tools/multi_gpu_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,b_local,D", [(1, 256, 256), (2, 256, 256), (4, 1024, 256), (3, 100, 200)])
def test_statistics_exchange_emulated_ranks_on_one_gpu(cuda_device, world, b_local, D):
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stats_emulate.py"), str(world), str(b_local),
                           str(D)], capture_output=True, text=True, timeout=600)
    print(proc.stdout[-3000:])
    print(proc.stderr[-2000:])
    assert proc.returncode == 0 and "stats exchange emulation PASS" in proc.stdout


def test_all_exchange_routes_across_processes(cuda_device):
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs (the one-GPU emulation above covers the statistics exchange kernels)")
    world = 2 if n < 4 else 4
    port = 29600 + os.getpid() % 1000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(proc.stdout[-4000:])
    print(proc.stderr[-2000:])
    assert proc.returncode == 0 and "FAIL" not in proc.stdout and "PASS" in proc.stdout
