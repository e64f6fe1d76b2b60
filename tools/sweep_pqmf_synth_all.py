"""Synthesis of every specialised band count, 1024 x 4 s, device-timed.  Usage: python tools/sweep_pqmf_synth_all.py"""
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


for packed, q in (("1", "8"), ("1", "4"), ("0", "4"), ("0", "8")):  # N = 4: packed kernel against the scalar one
    os.environ["IAS_PQMF_SYNTH_PACKED"], os.environ["IAS_PQMF_SYNTH_Q"] = packed, q
    m = ias_b200.PQMF(N=4).to(dev)
    z = m.analysis(x)
    mss, y = timed(lambda: m.synthesis(z))
    print(f"N=4 packed={packed} Q={q}: synthesis {mss:.4f} ms ({8.0 * T * B / 1e6 / mss:.0f} GB/s)", flush=True)
os.environ.pop("IAS_PQMF_SYNTH_PACKED")
os.environ.pop("IAS_PQMF_SYNTH_Q")
for rep in range(2):  # N = 8, 16: two steps per thread (cm2) against one (cm)
    for N in (8, 16):
        for cm2 in ("1", "0"):
            os.environ["IAS_PQMF_SYNTH_CM2"] = cm2
            m = ias_b200.PQMF(N=N).to(dev)
            z = m.analysis(x)
            mss, y = timed(lambda: m.synthesis(z))
            print(f"N={N} cm2={cm2}: synthesis {mss:.4f} ms ({8.0 * T * B / 1e6 / mss:.0f} GB/s)", flush=True)
os.environ.pop("IAS_PQMF_SYNTH_CM2")
for rep in range(2):
    for N in (2, 3, 4, 8, 16):
        m = ias_b200.PQMF(N=N).to(dev)
        msa, z = timed(lambda: m.analysis(x))
        mss, y = timed(lambda: m.synthesis(z))
        gb = 8.0 * T * B / 1e6
        print(f"N={N}: analysis {msa:.4f} ms ({gb / msa:.0f} GB/s), synthesis {mss:.4f} ms ({gb / mss:.0f} GB/s)", flush=True)
