"""Pooled N = 3 analysis (the kernel on the bench step) and the plain one, 1024 x 4 s, device-timed.
Usage on the GPU box: [IAS_B200_LIB=...] python tools/sweep_pqmf_pooled.py"""
import os, sys, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import harness, ias_b200  # noqa: E401,F401
dev = torch.device("cuda:0")
B, T = 1024, 176400
x = (torch.rand((B, 1, T), device=dev) * 2 - 1)
scale = torch.rand(B, device=dev) * 0.5 + 0.5


def timed(fn, iters=30):
    out = None
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


m = ias_b200.PQMF(N=3).to(dev)
for rep in range(2):
    ms, (bands, feat) = timed(lambda: m.analysis_pooled(x, 256, row_scale=scale))
    ref = torch.nn.functional.adaptive_avg_pool1d(bands.abs().reshape(B, 1, -1), 256).squeeze(1)
    err = float((feat - ref).abs().max() / ref.abs().max())
    ms2, _ = timed(lambda: m.analysis(x, row_scale=scale))
    print(f"pooled analysis (+ finalize + memset) {ms:.4f} ms, plain {ms2:.4f} ms, pooled features vs torch pooling {err:.2e}", flush=True)
