"""CPU, world_size 2 over gloo: the embedding all-gather that replaces the reference's dead FullGatherLayer
(vicreg.py:79-95, call site vicreg.py:38-39).  N-rank result == 1-rank result on the rank-ordered concatenation,
forward and backward.  The loss arithmetic is stood in by the torch restatement of the oracle (the CUDA kernels need
a GPU); what is under test is the gather layer and the local-row bookkeeping of ias_b200.vicreg.vicreg_loss."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT

WORLD = 2


def _oracle_loss_fn():
    """torch-autograd stand-in with the signature of _VicregLossFn.apply (oracle arithmetic, CPU)."""
    import torch.nn.functional as F

    def apply(x, y, row0, b_local, cfg_batch, embeddim, sim, stdc, covc, holder):
        repr_loss = F.mse_loss(x[row0:row0 + b_local], y[row0:row0 + b_local])
        xc = x - x.mean(dim=0)
        yc = y - y.mean(dim=0)
        std_x = torch.sqrt(xc.var(dim=0) + 0.0001)
        std_y = torch.sqrt(yc.var(dim=0) + 0.0001)
        std_loss = torch.mean(F.relu(1 - std_x)) / 2 + torch.mean(F.relu(1 - std_y)) / 2
        cov_x = (xc.T @ xc) / (cfg_batch - 1)
        cov_y = (yc.T @ yc) / (cfg_batch - 1)
        off = lambda m: m - torch.diag(torch.diag(m))  # noqa: E731
        cov_loss = off(cov_x).pow(2).sum() / embeddim + off(cov_y).pow(2).sum() / embeddim
        return sim * repr_loss + stdc * std_loss + covc * cov_loss, repr_loss, std_loss, cov_loss

    return apply


def _worker(rank, port, tmp):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        import ias_b200
        from ias_b200 import vicreg as V
        from oracle import make_golden as MG

        B_local, D = 16, 32
        x_all, y_all = MG.vicreg_inputs(WORLD * B_local, D, "correlated", seed=9)
        x_all, y_all = x_all.double(), y_all.double()
        x = x_all[rank * B_local:(rank + 1) * B_local].clone().requires_grad_(True)
        y = y_all[rank * B_local:(rank + 1) * B_local].clone().requires_grad_(True)

        # 1. the layer itself: forward = rank-ordered tuple, backward = sum over ranks of the incoming grads, own slice
        parts = ias_b200.FullGatherLayer.apply(x)
        assert isinstance(parts, tuple) and len(parts) == WORLD
        gathered = torch.cat(parts, dim=0)
        assert torch.equal(gathered.detach(), x_all)
        w = torch.arange(WORLD * B_local * D, dtype=torch.float64).view(WORLD * B_local, D) * (rank + 1)
        (gathered * w).sum().backward()
        total_w = sum(torch.arange(WORLD * B_local * D, dtype=torch.float64).view(WORLD * B_local, D) * (r + 1)
                      for r in range(WORLD))
        assert torch.equal(x.grad, total_w[rank * B_local:(rank + 1) * B_local])
        x.grad = None

        # 2. vicreg_loss with the gather: global std/cov terms, local invariance term, gradients through the gather
        V._VicregLossFn.apply = staticmethod(_oracle_loss_fn())
        out = V.vicreg_loss(x, y, WORLD * B_local, D, 25.0, 25.0, 1.0, gather=True)
        out[0].backward()
        res = dict(loss4=[float(o) for o in out], gx=x.grad.numpy(), gy=y.grad.numpy())
        np.savez(os.path.join(tmp, f"rank{rank}.npz"), **res)
    finally:
        dist.destroy_process_group()


def test_full_gather_layer_two_ranks(tmp_path, built_lib):
    from oracle import make_golden as MG
    from oracle import vicreg as OV

    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    B_local, D = 16, 32
    x_all, y_all = MG.vicreg_inputs(WORLD * B_local, D, "correlated", seed=9)
    xa, ya = x_all.double().numpy(), y_all.double().numpy()
    full = OV.loss(xa, ya, WORLD * B_local, D)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(WORLD)]
    for k in range(WORLD):
        want = OV.loss(xa, ya, WORLD * B_local, D, local_rows=slice(k * B_local, (k + 1) * B_local))
        assert np.allclose(r[k]["loss4"], want, rtol=1e-10)
        assert abs(r[k]["loss4"][2] - full[2]) < 1e-12 and abs(r[k]["loss4"][3] - full[3]) < 1e-12
    # mean over ranks of the invariance term == single-process value (equal shards)
    assert abs(np.mean([r[k]["loss4"][1] for k in range(WORLD)]) - full[1]) < 1e-12
    # DDP semantics: every rank back-propagates its own loss; the gather's backward sums the ranks' gradients.
    # Sum over ranks of loss_r = W * (std + cov terms) + sum_r repr_r, so the per-rank gradient equals W times the
    # single-process gradient of the std/cov terms plus W times the repr gradient restricted to the own rows
    # (repr_r averages over B_local rows, the single-process term over W*B_local).
    gx_full, gy_full = OV.loss_grad(xa, ya, WORLD * B_local, D)
    for k in range(WORLD):
        sl = slice(k * B_local, (k + 1) * B_local)
        assert np.allclose(r[k]["gx"], WORLD * gx_full[sl], rtol=1e-8, atol=1e-12)
        assert np.allclose(r[k]["gy"], WORLD * gy_full[sl], rtol=1e-8, atol=1e-12)
