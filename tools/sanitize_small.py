"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
Voice (seed, ADSR, control, schedule, audio incl. the normalise pass), PQMF analysis (plain, image, pooled) and
synthesis for N = 3 and 16, VICReg forward + backward (tcgen05 Gram), and the statistics exchange with two emulated
ranks (publish on both, then combine on both)."""
import ctypes
import os
import sys
import types

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import harness  # noqa: E402,F401
import ias_b200  # noqa: E402
from ias_b200 import _lib  # noqa: E402
from ias_b200.vicreg import aligned_workspace  # noqa: E402

dev = torch.device("cuda:0")
lib = ias_b200.lib()
for repro in (True, False):
    cfg = ias_b200.SynthConfig(batch_size=32, reproducible=repro, sample_rate=44100, buffer_size_seconds=0.25)
    voice = ias_b200.Voice(synthconfig=cfg).to(dev)
    audio, params, is_train = voice(3)
    voice.prepare(4)
    audio, params, is_train = voice(4, prepared=True)
for N in (3, 16):
    gram = ias_b200.PQMF(N=N).to(dev)
    z = gram(audio.unsqueeze(1))
    y = gram.synthesis(z)
long_in = (torch.rand((2, 1, 176400), generator=torch.Generator().manual_seed(7)) * 2 - 1).to(dev)
gram3 = ias_b200.PQMF(N=3).to(dev)
bands, feat = gram3.analysis_pooled(long_in, 256)
img = gram3.analysis_image(long_in, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225], image_shape=(240, 245))
vcfg = types.SimpleNamespace(dim=256, embeddim=256, vicreg=types.SimpleNamespace(
    mlp="8-8-%d", batch_size=256, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0))
vic = ias_b200.VICReg(vcfg, torch.nn.Identity(), torch.nn.Identity())
x = torch.randn(256, 256, device=dev, requires_grad=True)
yv = torch.randn(256, 256, device=dev)
out = vic.loss(x, yv)
out[0].backward()
W, Bl, D = 2, 128, 256
nbytes = lib.ias_vicreg_stats_buffer_bytes(W, D)
bufs = []
for _ in range(W):
    b = torch.zeros(nbytes // 4, dtype=torch.float32, device=dev)
    b.view(torch.int32)[32] = 1
    bufs.append(b)
ptrs = (ctypes.c_void_p * W)(*[b.data_ptr() for b in bufs])
wss = [aligned_workspace(lib.ias_vicreg_workspace_bytes(Bl, D), dev) for _ in range(W)]
outs = [torch.empty(4, device=dev) for _ in range(W)]
xs = [torch.randn(Bl, D, device=dev) for _ in range(W)]
ys = [torch.randn(Bl, D, device=dev) for _ in range(W)]
for step in range(2):
    for stage in (1, 2):
        for q in range(W):
            _lib.check(lib.ias_vicreg_loss_stats_stages(_lib.ptr(xs[q]), _lib.ptr(ys[q]), ptrs, W, q, Bl, W * Bl, D, D, 25.0,
                                                        25.0, 1.0, _lib.ptr(outs[q]), _lib.ptr(wss[q]), wss[q].numel() * 4,
                                                        stage, _lib.current_stream(dev)))
torch.cuda.synchronize()
print("sanitize_small done", [float(o[0]) for o in outs], float(out[0]), float(feat.sum()), float(y.abs().max()))
