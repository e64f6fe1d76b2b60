import os, sys, ctypes, types, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import harness, ias_b200
from ias_b200 import _lib
dev = torch.device("cuda:0")
lib = ias_b200.lib()
for B in (1024, 8192):
    x = torch.randn(B, 256, device=dev); y = x + 0.1 * torch.randn(B, 256, device=dev)
    cfg = types.SimpleNamespace(dim=256, embeddim=256, vicreg=types.SimpleNamespace(mlp="8-8-%d", batch_size=B, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0))
    vic = ias_b200.VICReg(cfg, torch.nn.Identity(), torch.nn.Identity())
    with torch.no_grad():
        for _ in range(5): vic.loss(x, y)
        torch.cuda.synchronize()
        lib.ias_prof_reset(); lib.ias_prof_enable(1)
        for _ in range(20): out = vic.loss(x, y)
        torch.cuda.synchronize(); lib.ias_prof_enable(0)
    k = {}
    for i in range(lib.ias_prof_kernel_count()):
        tot, n = ctypes.c_double(0), ctypes.c_longlong(0)
        lib.ias_prof_read(i, ctypes.byref(tot), ctypes.byref(n))
        if n.value: k[lib.ias_prof_kernel_name(i).decode()] = round(tot.value / n.value * 1e3, 1)
    print(B, k, "sum", round(sum(k.values()), 1), "us")
