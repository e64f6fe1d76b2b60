import os, sys, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import harness, ias_b200
dev = torch.device("cuda:0")
B, T = 1024, 176400
x = (torch.rand((B, 1, T), device=dev) * 2 - 1)
def timed(fn, iters=10):
    out = None
    for _ in range(3): out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out
for N, qs in ((3, (4, 8, 0)), (4, (4, 8, 0))):
    m = ias_b200.PQMF(N=N).to(dev)
    z = m.analysis(x)
    ref = None
    for q in qs:
        os.environ["IAS_PQMF_SYNTH_Q"] = str(q)
        if q == 0:
            os.environ["IAS_PQMF_SYNTH_DIRECT"] = "1"  # the direct-form kernel
        else:
            os.environ.pop("IAS_PQMF_SYNTH_DIRECT", None)
        ms, y = timed(lambda: m.synthesis(z))
        if ref is None: ref = y.clone()
        print(f"N={N} Q={q}: synthesis {ms:.4f} ms, max diff vs first {float((y-ref).abs().max()):.2e}", flush=True)
m = ias_b200.PQMF(N=16).to(dev)
z = m.analysis(x)
ms, y = timed(lambda: m.synthesis(z))
print(f"N=16 cosine-modulated: {ms:.4f} ms", flush=True)
