"""GPU parity: VICReg loss kernels (tcgen05 Gram included) through ias_b200.VICReg / the C ABI, against the oracle
and the goldens frozen from the reference's vicreg.VICReg.loss.  Tolerance (north star): 1e-4 relative per term."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import make_golden as MG
from oracle import vicreg as OV

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _cfg(D, E, cfgB):
    return types.SimpleNamespace(dim=D, embeddim=E, vicreg=types.SimpleNamespace(
        mlp="8-8-%d", batch_size=cfgB, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0))


def _ws(B, D, dev):
    from ias_b200.vicreg import _Workspace

    return _Workspace().workspace(B, D, dev)


@pytest.mark.parametrize("B,D", [(8192, 256), (1000, 128), (100, 256), (32, 128), (4096, 512)])
def test_tcgen05_gram_vs_cuda_core_and_float64(cuda_device, B, D):
    import ias_b200
    from ias_b200 import _lib

    lib = ias_b200.lib()
    g = torch.Generator().manual_seed(B + D)
    x = (torch.randn((B, D), generator=g) * torch.rand((1, D), generator=g) * 3 + torch.randn((1, D), generator=g)).to(
        cuda_device)
    ws = _ws(B, D, cuda_device)
    st = _lib.current_stream(cuda_device)
    gram_tc = torch.empty((2, D, D), device=cuda_device)
    _lib.check(lib.ias_vicreg_gram_tc(_lib.ptr(x), B, D, _lib.ptr(gram_tc), _lib.ptr(ws), ws.numel() * 4, st))
    gram_cc = torch.empty((D, D), device=cuda_device)
    _lib.check(lib.ias_vicreg_gram_reference(_lib.ptr(x), B, D, _lib.ptr(gram_cc), _lib.ptr(ws), ws.numel() * 4, st))
    torch.cuda.synchronize()
    xd = x.double().cpu()
    xd = xd - xd.mean(dim=0)
    ref = (xd.T @ xd).numpy()
    scale = np.abs(ref).max()
    err_tc = np.abs(gram_tc[0].cpu().numpy() - ref).max() / scale
    err_cc = np.abs(gram_cc.cpu().numpy() - ref).max() / scale
    print(f"gram B={B} D={D}: tcgen05 3xTF32 rel err {err_tc:.2e}, CUDA-core fp32 rel err {err_cc:.2e}")
    assert err_cc <= 1e-5
    assert err_tc <= 1e-5  # 3xTF32 split keeps fp32-class accuracy (single-pass TF32 would be ~1e-3)
    assert torch.equal(gram_tc[0], gram_tc[1])


@pytest.mark.parametrize("name,B,D,kind,cfgB,E", MG.VICREG_CASES)
def test_loss_vs_reference_goldens(cuda_device, name, B, D, kind, cfgB, E):
    import ias_b200

    gold = np.load(os.path.join(GOLDEN, "vicreg_cases.npz"))[f"{name}_loss4"]
    x, y = MG.vicreg_inputs(B, D, kind)
    m = ias_b200.VICReg(_cfg(D, E, cfgB), torch.nn.Identity(), torch.nn.Identity())
    x0, y0 = x.clone(), y.clone()
    with torch.no_grad():
        out = m.loss(x.to(cuda_device), y.to(cuda_device))
    got = np.array([float(o) for o in out])
    assert all(o.dim() == 0 for o in out)
    rel = np.abs(got - gold) / np.abs(gold)
    print(name, got, gold, rel)
    assert np.all(rel <= TOL)
    assert torch.equal(x, x0) and torch.equal(y, y0)  # inputs not mutated
    # and the oracle on the same inputs
    ora = np.array(OV.loss(x.numpy(), y.numpy(), cfgB, E))
    assert np.all(np.abs(got - ora) <= TOL * np.abs(ora))


def test_loss_is_deterministic_and_handles_local_rows(cuda_device):
    from ias_b200 import _lib
    import ias_b200

    lib = ias_b200.lib()
    B, D = 2048, 256
    x, y = MG.vicreg_inputs(B, D, "correlated", seed=4)
    xd, yd = x.to(cuda_device), y.to(cuda_device)
    ws = _ws(B, D, cuda_device)
    outs = []
    for row0, bl in [(0, B), (512, 256), (0, B)]:
        out4 = torch.empty(4, device=cuda_device)
        _lib.check(lib.ias_vicreg_loss(_lib.ptr(xd), _lib.ptr(yd), B, row0, bl, B, D, D, 25.0, 25.0, 1.0,
                                       _lib.ptr(out4), _lib.ptr(ws), ws.numel() * 4,
                                       _lib.current_stream(cuda_device)))
        outs.append(out4.cpu().numpy().astype(np.float64))
    assert np.array_equal(outs[0], outs[2])  # bitwise reproducible (no atomics anywhere in the pipeline)
    ref_local = np.array(OV.loss(x.numpy(), y.numpy(), B, D, local_rows=slice(512, 768)))
    assert np.all(np.abs(outs[1] - ref_local) <= TOL * np.abs(ref_local))
    # std/cov terms are global: identical whatever the local range
    assert outs[0][2] == outs[1][2] and outs[0][3] == outs[1][3]


@pytest.mark.parametrize("name,B,D,kind,cfgB,E", MG.VICREG_CASES)
def test_backward_vs_reference_goldens(cuda_device, name, B, D, kind, cfgB, E):
    """d loss / d x, d y through the autograd Function (tcgen05 backward GEMM) vs torch autograd on the reference."""
    import ias_b200

    gold = np.load(os.path.join(GOLDEN, "vicreg_cases.npz"))
    x, y = MG.vicreg_inputs(B, D, kind)
    xd = x.to(cuda_device).requires_grad_(True)
    yd = y.to(cuda_device).requires_grad_(True)
    m = ias_b200.VICReg(_cfg(D, E, cfgB), torch.nn.Identity(), torch.nn.Identity())
    out = m.loss(xd, yd)
    out[0].backward()
    gx, gy = xd.grad.cpu().numpy(), yd.grad.cpu().numpy()
    if B <= 128:
        rx, ry = gold[f"{name}_gx"], gold[f"{name}_gy"]
    else:
        gx, gy, rx, ry = gx[::64], gy[::64], gold[f"{name}_gx_sub"], gold[f"{name}_gy_sub"]
    ex = np.abs(gx - rx).max() / np.abs(rx).max()
    ey = np.abs(gy - ry).max() / np.abs(ry).max()
    print(f"{name}: grad rel err x {ex:.2e} y {ey:.2e}")
    assert ex <= 1e-4 and ey <= 1e-4


def test_backward_of_individual_terms_and_local_rows(cuda_device):
    from ias_b200.vicreg import vicreg_loss

    B, D = 512, 256
    x, y = MG.vicreg_inputs(B, D, "correlated", seed=8)
    xd = x.to(cuda_device).requires_grad_(True)
    yd = y.to(cuda_device).requires_grad_(True)
    out = vicreg_loss(xd, yd, B, D, 25.0, 25.0, 1.0)
    (2.0 * out[1] + 3.0 * out[3] + 0.5 * out[2]).backward()
    gx, gy = OV.loss_grad(x.numpy(), y.numpy(), B, D, sim_coeff=2.0, std_coeff=0.5, cov_coeff=3.0)
    assert np.abs(xd.grad.cpu().numpy() - gx).max() <= 1e-4 * np.abs(gx).max()
    assert np.abs(yd.grad.cpu().numpy() - gy).max() <= 1e-4 * np.abs(gy).max()
    # local rows: the invariance gradient lives only on [row0, row0 + b_local) and is scaled by 1 / b_local
    from ias_b200 import _lib
    import ias_b200

    lib = ias_b200.lib()
    ws = _ws(B, D, cuda_device)
    out4 = torch.empty(4, device=cuda_device)
    xc, yc = x.to(cuda_device), y.to(cuda_device)
    st = _lib.current_stream(cuda_device)
    _lib.check(lib.ias_vicreg_loss(_lib.ptr(xc), _lib.ptr(yc), B, 128, 64, B, D, D, 25.0, 25.0, 1.0, _lib.ptr(out4),
                                   _lib.ptr(ws), ws.numel() * 4, st))
    gout = torch.tensor([0.0, 1.0, 0.0, 0.0], device=cuda_device)
    gxd, gyd = torch.empty_like(xc), torch.empty_like(yc)
    _lib.check(lib.ias_vicreg_loss_backward(_lib.ptr(xc), _lib.ptr(yc), B, 128, 64, B, D, D, 25.0, 25.0, 1.0,
                                            _lib.ptr(gout), _lib.ptr(gxd), _lib.ptr(gyd), _lib.ptr(ws), ws.numel() * 4,
                                            st))
    want = torch.zeros_like(x)
    want[128:192] = 2.0 * (x[128:192] - y[128:192]) / (64 * D)
    assert float((gxd.cpu() - want).abs().max()) <= 1e-6 * float(want.abs().max())
    assert float((gyd.cpu() + want).abs().max()) <= 1e-6 * float(want.abs().max())


def test_two_forwards_before_the_first_backward_keep_their_own_state(cuda_device):
    """A second VICReg.loss() on the same module before the first backward (two view pairs, a diagnostic loss on another
    batch, a different B) must not disturb the first one's gradients: the C ABI keeps the backward's inputs in the
    workspace of the forward call, so every differentiable forward owns its workspace (ias_b200/vicreg.py)."""
    import ias_b200

    D = 256
    m = ias_b200.VICReg(_cfg(D, D, 512), torch.nn.Identity(), torch.nn.Identity())

    def grads(x, y, between=None):
        xd = x.to(cuda_device).requires_grad_(True)
        yd = y.to(cuda_device).requires_grad_(True)
        out = m.loss(xd, yd)
        if between is not None:
            between()
        out[0].backward()
        return xd.grad.clone(), yd.grad.clone(), torch.stack([o.detach() for o in out])

    x1, y1 = MG.vicreg_inputs(512, D, "correlated", seed=5)
    x2, y2 = MG.vicreg_inputs(384, D, "correlated", seed=6)
    want = grads(x1, y1)
    held = []

    def other_grad_forward():  # differentiable forward on another batch and another B, left pending
        a = (x2 * 3.0).to(cuda_device).requires_grad_(True)
        held.append((a, m.loss(a, y2.to(cuda_device))))

    def other_nograd_forward():
        with torch.no_grad():
            m.loss((y1 * 2.0).to(cuda_device), x1.to(cuda_device))

    for between in (other_grad_forward, other_nograd_forward):
        got = grads(x1, y1, between)
        for a, b in zip(got, want):
            assert torch.equal(a, b)
    # the pending forward still back-propagates its own state afterwards
    a, out = held[0]
    out[0].backward()
    gx, _ = OV.loss_grad((x2 * 3.0).numpy(), y2.numpy(), 512, D)
    assert np.abs(a.grad.cpu().numpy() - gx).max() <= 1e-4 * np.abs(gx).max()


def test_bad_arguments_fail_loudly(cuda_device):
    from ias_b200 import _lib
    import ias_b200

    lib = ias_b200.lib()
    x = torch.zeros((64, 64), device=cuda_device)
    out4 = torch.empty(4, device=cuda_device)
    rc = lib.ias_vicreg_loss(_lib.ptr(x), _lib.ptr(x), 64, 0, 64, 64, 64, 64, 25.0, 25.0, 1.0, _lib.ptr(out4), None, 0,
                             _lib.current_stream(cuda_device))
    assert rc == 4 and b"workspace" in lib.ias_last_error()
    ws = _ws(64, 64, cuda_device)
    rc = lib.ias_vicreg_loss(_lib.ptr(x), _lib.ptr(x), 64, 32, 64, 64, 64, 64, 25.0, 25.0, 1.0, _lib.ptr(out4),
                             _lib.ptr(ws), ws.numel() * 4, _lib.current_stream(cuda_device))
    assert rc == 1
