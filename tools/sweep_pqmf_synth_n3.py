"""N = 3 synthesis: packed (k_pqmf_synthesis_n3p) against the scalar cosine-modulated kernel and the direct form,
1024 x 4 s, device-timed.  Usage on the GPU box: python tools/sweep_pqmf_synth_n3.py"""
import os, sys, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import harness, ias_b200  # noqa: E401,F401
dev = torch.device("cuda:0")
B, T = 1024, 176400
x = (torch.rand((B, 1, T), device=dev) * 2 - 1)


def timed(fn, iters=20):
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
z = m.analysis(x)
ref = None
CASES = (("1", "8"), ("1", "88"), ("1", "16"), ("1", "4"), ("0", "8"))
if os.environ.get("IAS_SWEEP_ONLY"):  # A/B of variant libraries: the default shape and the scalar kernel only
    CASES = (("1", os.environ["IAS_SWEEP_ONLY"]), ("0", "8"))
for rep in range(2):
    for packed, q in CASES:
        os.environ["IAS_PQMF_SYNTH_PACKED"] = packed
        os.environ["IAS_PQMF_SYNTH_Q"] = q
        ms, y = timed(lambda: m.synthesis(z))
        if ref is None:
            ref = y.clone()
        gbs = 8.0 * T * B / (ms * 1e-3) / 1e9
        print(f"N=3 packed={packed} Q={q}: synthesis {ms:.4f} ms = {gbs:.0f} GB/s, max diff vs first "
              f"{float((y - ref).abs().max()):.2e}", flush=True)
os.environ.pop("IAS_PQMF_SYNTH_PACKED")
os.environ.pop("IAS_PQMF_SYNTH_Q")
ms, zz = timed(lambda: m.analysis(x))
print(f"N=3 analysis {ms:.4f} ms", flush=True)
