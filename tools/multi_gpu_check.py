"""torchrun --nproc-per-node W tools/multi_gpu_check.py : N-rank loss == oracle on the rank-ordered concatenation.

Checks the embedding exchange on real GPUs through every route (torch.distributed NCCL, the C-ABI
communicator in libias_comm.so, the fused peer-read gather and the statistics exchange), forward and backward, plus
the sharded front end: rank r renders sound ids [(step*W + r)*B_local, ...).
Rank 0 prints PASS/FAIL lines and exits non-zero on failure."""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import harness  # noqa: E402
import ias_b200  # noqa: E402
from oracle import make_golden as MG  # noqa: E402
from oracle import vicreg as OV  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    B_local, D = 256, 256
    x_all, y_all = MG.vicreg_inputs(world * B_local, D, "correlated", seed=21)
    sl = slice(rank * B_local, (rank + 1) * B_local)
    want = np.array(OV.loss(x_all.numpy(), y_all.numpy(), world * B_local, D, local_rows=sl))
    for route in ("torch.distributed", "libias_comm"):
        if route == "libias_comm":
            comm = ias_b200.Communicator()
            ias_b200.use_communicator(comm)
        with torch.no_grad():
            out = ias_b200.vicreg_loss(x_all[sl].to(dev), y_all[sl].to(dev), world * B_local, D, 25.0, 25.0, 1.0)
        got = np.array([float(o) for o in out])
        rel = np.abs(got - want) / np.abs(want)
        good = bool(np.all(rel <= 1e-4))
        flags = torch.tensor([1 if good else 0], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"{'PASS' if flags.item() else 'FAIL'} gather+loss via {route}: rank0 terms {got} want {want} rel {rel}")
        ok = ok and bool(flags.item())
    if True:
        # reduce-scatter semantics of the backward through the communicator
        g = torch.full((world, 8, 4), float(rank + 1), device=dev)
        own = ias_b200.dist.active().reduce_scatter(g)
        expect = float(sum(range(1, world + 1)))
        good = bool(torch.all(own == expect))
        if rank == 0:
            print(f"{'PASS' if good else 'FAIL'} reduce_scatter via libias_comm: {own.flatten()[0].item()} == {expect}")
        ok = ok and good
        ias_b200.use_communicator(None)
    # fused gather: the statistics kernel reads every rank's embeddings over NVLink itself (no collective launch)
    ex = ias_b200.EmbeddingExchange(B_local, D, dev)
    ias_b200.use_fused_gather(ex)
    for it in range(3):  # several rounds: exercises the reuse barriers of the exchange buffer
        xs = (x_all[sl] * (1.0 + 0.25 * it)).to(dev).requires_grad_(True)
        ys = (y_all[sl] - 0.5 * it).to(dev).requires_grad_(True)
        out = ias_b200.vicreg_loss(xs, ys, world * B_local, D, 25.0, 25.0, 1.0)
        out[0].backward()
        xa, ya = (x_all * (1.0 + 0.25 * it)).numpy(), (y_all - 0.5 * it).numpy()
        want_it = np.array(OV.loss(xa, ya, world * B_local, D, local_rows=sl))
        got = np.array([float(o) for o in out])
        gx_full, gy_full = OV.loss_grad(xa, ya, world * B_local, D)  # grads of the single-process loss
        # sum over ranks of d L_r / d x_own = world * (single-process gradient) on the own rows (see test_dist_gloo)
        egx = np.abs(xs.grad.cpu().numpy() - world * gx_full[sl]).max() / np.abs(world * gx_full[sl]).max()
        egy = np.abs(ys.grad.cpu().numpy() - world * gy_full[sl]).max() / np.abs(world * gy_full[sl]).max()
        rel = np.abs(got - want_it) / np.abs(want_it)
        good = bool(np.all(rel <= 1e-4)) and egx <= 1e-4 and egy <= 1e-4
        flags = torch.tensor([1 if good else 0], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"{'PASS' if flags.item() else 'FAIL'} fused gather round {it}: rel {rel} grad rel {egx:.2e} {egy:.2e}")
        ok = ok and bool(flags.item())
    ias_b200.use_fused_gather(None)
    # statistics exchange (default route of bench.py): every rank reduces its own rows, pushes a summary to its peers
    # from inside the loss kernels and combines the W summaries with the pooled formulas
    sx = ias_b200.StatsExchange(D, dev)
    ias_b200.use_fused_gather(sx)
    for it in range(4):  # several rounds: both parities of the double-buffered inbox, twice
        xs = (x_all[sl] * (1.0 + 0.25 * it) + 0.3 * rank).to(dev).requires_grad_(True)   # rank-dependent means
        ys = (y_all[sl] - 0.5 * it).to(dev).requires_grad_(True)
        out = ias_b200.vicreg_loss(xs, ys, world * B_local, D, 25.0, 25.0, 1.0)
        if it == 2:  # a second (diagnostic) forward before the backward of the first must not disturb it
            with torch.no_grad():
                ias_b200.vicreg_loss(ys.detach() * 2.0, xs.detach(), world * B_local, D, 25.0, 25.0, 1.0)
        out[0].backward()
        shift = torch.arange(world, dtype=torch.float32).repeat_interleave(B_local)[:, None] * 0.3
        xa, ya = (x_all * (1.0 + 0.25 * it) + shift).numpy(), (y_all - 0.5 * it).numpy()
        want_it = np.array(OV.loss(xa, ya, world * B_local, D, local_rows=sl))
        got = np.array([float(o) for o in out])
        gx_full, gy_full = OV.loss_grad(xa, ya, world * B_local, D)
        egx = np.abs(xs.grad.cpu().numpy() - world * gx_full[sl]).max() / np.abs(world * gx_full[sl]).max()
        egy = np.abs(ys.grad.cpu().numpy() - world * gy_full[sl]).max() / np.abs(world * gy_full[sl]).max()
        rel = np.abs(got - want_it) / np.abs(want_it)
        good = bool(np.all(rel <= 1e-4)) and egx <= 1e-4 and egy <= 1e-4
        flags = torch.tensor([1 if good else 0], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"{'PASS' if flags.item() else 'FAIL'} statistics exchange round {it}: rel {rel} grad rel {egx:.2e} {egy:.2e}")
        ok = ok and bool(flags.item())
    ias_b200.use_fused_gather(None)
    # sharded front end: this rank's sounds equal the same ids rendered by a single process
    cfg = ias_b200.SynthConfig(batch_size=64, reproducible=True, buffer_size_seconds=0.5)
    voice = ias_b200.Voice(cfg).to(dev)
    step = 3
    audio, params, _ = voice(step * world + rank)
    from oracle import voice as OVc
    ref_params = OVc.sorted_to_registration(OVc.seeded_params(step * world + rank, 64))
    good = torch.equal(params.cpu(), ref_params)
    ids_first = (step * world + rank) * 64
    flags = torch.tensor([1 if good else 0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{'PASS' if flags.item() else 'FAIL'} sharded seeding: rank {rank} renders ids from {ids_first}")
    ok = ok and bool(flags.item())
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
