"""python tools/stats_emulate.py [W] -- the statistics exchange (ias_vicreg_loss_stats) with W emulated ranks on ONE GPU.

Every emulated rank has its own exchange buffer (plain device memory: peer-accessible by construction), workspace and
CUDA stream; all ranks publish, then all ranks combine (ias_vicreg_loss_stats_stages), so each rank's combine kernel
reads the summaries and flags the other ranks' publish kernels wrote into its inbox.  Checks, for several steps (both inbox parities) and rank-dependent means:
every rank's four loss terms == oracle on the rank-ordered concatenation (<= 1e-4 relative), std/cov terms
bit-identical across ranks, backward == world x the single-process gradient on the own rows (vicreg.py:92-95).
Runs in its own process (tests/test_gpu_multi.py) so that a timeout trap cannot poison the caller's CUDA context."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import harness  # noqa: E402,F401  (sets sys.path)
import ias_b200  # noqa: E402
from ias_b200 import _lib  # noqa: E402
from ias_b200.vicreg import aligned_workspace  # noqa: E402
from oracle import make_golden as MG  # noqa: E402
from oracle import vicreg as OV  # noqa: E402


def main():
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    B_local = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    D = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    dev = torch.device("cuda:0")
    lib = ias_b200.lib()
    nbytes = lib.ias_vicreg_stats_buffer_bytes(W, D)
    bufs = []
    for _ in range(W):
        b = torch.zeros(nbytes // 4, dtype=torch.float32, device=dev)
        b.view(torch.int32)[32] = 1
        bufs.append(b)
    ptrs = (ctypes.c_void_p * W)(*[b.data_ptr() for b in bufs])
    wss = [aligned_workspace(lib.ias_vicreg_workspace_bytes(B_local, D), dev) for _ in range(W)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(W)]
    outs = [torch.empty(4, device=dev) for _ in range(W)]
    x_all, y_all = MG.vicreg_inputs(W * B_local, D, "correlated", seed=33)
    shift = torch.arange(W, dtype=torch.float32).repeat_interleave(B_local)[:, None] * 0.3
    torch.cuda.synchronize()
    ok = True
    for it in range(4):
        xa = x_all * (1.0 + 0.25 * it) + shift
        ya = y_all - 0.5 * it
        xs = [xa[q * B_local:(q + 1) * B_local].contiguous().to(dev) for q in range(W)]
        ys = [ya[q * B_local:(q + 1) * B_local].contiguous().to(dev) for q in range(W)]
        torch.cuda.synchronize()
        # two sweeps (publish on every rank, then combine on every rank): ranks that share a GPU cannot be relied on to
        # run concurrently -- a combine kernel spinning on every SM keeps the tcgen05 Gram kernel of a later rank (which
        # needs the SMs re-partitioned for 197 KB of shared memory) from starting.  Across real GPUs the single call is
        # used (tools/multi_gpu_check.py).  Each rank still runs on its own stream, so the sweeps interleave freely.
        for stage in (1, 2):
            for q in range(W):
                with torch.cuda.stream(streams[q]):
                    rc = lib.ias_vicreg_loss_stats_stages(
                        _lib.ptr(xs[q]), _lib.ptr(ys[q]), ptrs, W, q, B_local, W * B_local, D, D, 25.0, 25.0, 1.0,
                        _lib.ptr(outs[q]), _lib.ptr(wss[q]), wss[q].numel() * 4, stage,
                        ctypes.c_void_p(streams[q].cuda_stream))
                    _lib.check(rc, "ias_vicreg_loss_stats_stages")
        torch.cuda.synchronize()
        got = np.stack([o.cpu().numpy() for o in outs])
        gx_full, gy_full = OV.loss_grad(xa.numpy(), ya.numpy(), W * B_local, D)
        gout = torch.tensor([1.0, 0.0, 0.0, 0.0], device=dev)
        for q in range(W):
            sl = slice(q * B_local, (q + 1) * B_local)
            want = np.array(OV.loss(xa.numpy(), ya.numpy(), W * B_local, D, local_rows=sl))
            rel = np.abs(got[q] - want) / np.abs(want)
            gx, gy = torch.empty_like(xs[q]), torch.empty_like(ys[q])
            rc = lib.ias_vicreg_loss_stats_backward(_lib.ptr(xs[q]), _lib.ptr(ys[q]), W, B_local, W * B_local, D, D, 25.0,
                                                    25.0, 1.0, _lib.ptr(gout), _lib.ptr(gx), _lib.ptr(gy),
                                                    _lib.ptr(wss[q]), wss[q].numel() * 4,
                                                    _lib.current_stream(dev))
            _lib.check(rc, "ias_vicreg_loss_stats_backward")
            egx = np.abs(gx.cpu().numpy() - W * gx_full[sl]).max() / np.abs(W * gx_full[sl]).max()
            egy = np.abs(gy.cpu().numpy() - W * gy_full[sl]).max() / np.abs(W * gy_full[sl]).max()
            good = bool(np.all(rel <= 1e-4)) and egx <= 1e-4 and egy <= 1e-4
            ok = ok and good
            if not good or q == 0:
                print(f"{'PASS' if good else 'FAIL'} step {it} rank {q}/{W}: rel {rel} grad rel {egx:.2e} {egy:.2e}")
        same = bool(np.all(got[:, 2:] == got[0, 2:]))
        print(f"{'PASS' if same else 'FAIL'} step {it}: std/cov terms bit-identical on all {W} ranks")
        ok = ok and same
    steps = [int(b.view(torch.int32)[32].item()) for b in bufs]
    print("step counters", steps)
    ok = ok and all(s == 5 for s in steps)
    print("stats exchange emulation", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
