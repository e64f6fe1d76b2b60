"""ORACLE tooling: freeze golden vectors from the *reference itself* (run in the build container only).

``/root/reference`` does not travel to the GPU box, so the outputs of the reference's own ``pqmf.PQMF`` and
``vicreg.VICReg.loss`` on seeded inputs are committed under ``tests/golden/``.  Inputs are regenerated from
their seeds by the tests (``torch.Generator`` CPU streams are machine independent); only outputs, subsamples
and checksums are stored so the fixtures stay small.

    python oracle/make_golden.py            # needs /root/reference

The Voice stage has no reference implementation to run (torchsynth is absent: SURVEY F1), so no synth golden is
written here -- see oracle/voice.py ("parity unpinned").
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("IAS_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

PQMF_CASES = [  # (name, N, cutoff, B, T)
    ("n3_t4096", 3, 0.15, 2, 4096),
    ("n3_t4001", 3, 0.15, 2, 4001),     # ragged: T % N != 0
    ("n4_t4096", 4, 0.15, 2, 4096),
    ("n16_t4099", 16, 0.15, 2, 4099),
    ("n16c003_t4096", 16, 0.03, 1, 4096),
    ("n3_t63", 3, 0.15, 1, 63),
    ("n3_t1", 3, 0.15, 1, 1),
    ("n2_t1000", 2, 0.15, 1, 1000),
    ("n8_t3000", 8, 0.15, 1, 3000),
]
PQMF_FULL = [("n3_full", 3, 0.15, 2, 176400), ("n16_full", 16, 0.15, 2, 176400)]
SUB = 97  # stride of the stored subsample for full-length cases


def pqmf_input(B: int, T: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand((B, 1, T), generator=g) * 2 - 1


def vicreg_inputs(B: int, D: int, kind: str, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((B, D), generator=g)
    if kind == "randn":
        y = x + 0.1 * torch.randn((B, D), generator=g)
    elif kind == "correlated":  # low-rank + offset: large means, strong off-diagonal covariance, some std < 1
        basis = torch.randn((8, D), generator=g)
        x = torch.randn((B, 8), generator=g) @ basis * 0.3 + 0.05 * x + 3.0
        y = x * 0.9 + 0.2 * torch.randn((B, D), generator=g) - 1.0
    else:
        raise ValueError(kind)
    return x.contiguous(), y.contiguous()


VICREG_CASES = [  # (name, B, D, kind, cfg_batch_size, embeddim)
    ("b128_d256_randn", 128, 256, "randn", 128, 256),
    ("b128_d256_corr", 128, 256, "correlated", 128, 256),
    ("b8192_d256_randn", 8192, 256, "randn", 8192, 256),
    ("b8192_d256_corr", 8192, 256, "correlated", 8192, 256),
    ("b100_d64_cfg16", 100, 64, "correlated", 16, 8192),   # cfg batch size / embeddim differ from the shapes
    ("b2_d32", 2, 32, "randn", 2, 32),
]


def main() -> None:
    sys.path.insert(0, REF)
    from pqmf import PQMF  # noqa: the reference module
    import vicreg as ref_vicreg  # noqa: the reference module

    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)

    blob = {}
    for N, cutoff in [(3, 0.15), (4, 0.15), (16, 0.15), (16, 0.03), (2, 0.15), (8, 0.15)]:
        m = PQMF(N=N, cutoff=cutoff)
        blob[f"H_n{N}_c{cutoff}"] = m.H[:, 0, :].numpy()
        blob[f"G_n{N}_c{cutoff}"] = m.G[0, :, :].numpy()
    np.savez_compressed(os.path.join(OUT, "pqmf_filters.npz"), **blob)

    blob = {}
    for name, N, cutoff, B, T in PQMF_CASES:
        m = PQMF(N=N, cutoff=cutoff)
        x = pqmf_input(B, T)
        z = m.analysis(x)
        y = m.synthesis(z)
        blob[f"{name}_analysis"] = z.numpy()
        blob[f"{name}_synthesis"] = y.numpy()
    for name, N, cutoff, B, T in PQMF_FULL:
        m = PQMF(N=N, cutoff=cutoff)
        x = pqmf_input(B, T)
        z = m.analysis(x)
        y = m.synthesis(z)
        blob[f"{name}_analysis_sub"] = z.numpy()[:, :, ::SUB].copy()
        blob[f"{name}_analysis_sum"] = z.double().sum(dim=2).numpy()
        blob[f"{name}_analysis_abssum"] = z.double().abs().sum(dim=2).numpy()
        blob[f"{name}_synthesis_sub"] = y.numpy()[:, :, ::SUB].copy()
        blob[f"{name}_synthesis_sum"] = y.double().sum(dim=2).numpy()
        blob[f"{name}_synthesis_abssum"] = y.double().abs().sum(dim=2).numpy()
    np.savez_compressed(os.path.join(OUT, "pqmf_cases.npz"), **blob)

    blob = {}
    for name, B, D, kind, cfgB, E in VICREG_CASES:
        cfg = types.SimpleNamespace(
            dim=D, embeddim=E,
            vicreg=types.SimpleNamespace(mlp="8-8-%d", batch_size=cfgB, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0),
        )
        mod = ref_vicreg.VICReg(cfg, torch.nn.Identity(), torch.nn.Identity())
        x, y = vicreg_inputs(B, D, kind)
        xg = x.clone().requires_grad_(True)
        yg = y.clone().requires_grad_(True)
        out = mod.loss(xg, yg)
        out[0].backward()
        blob[f"{name}_loss4"] = np.array([float(o.detach()) for o in out], dtype=np.float64)
        if B <= 128:
            blob[f"{name}_gx"] = xg.grad.numpy()
            blob[f"{name}_gy"] = yg.grad.numpy()
        else:
            blob[f"{name}_gx_sub"] = xg.grad.numpy()[::64].copy()
            blob[f"{name}_gy_sub"] = yg.grad.numpy()[::64].copy()
    blob["off_diagonal_4x4"] = ref_vicreg.off_diagonal(torch.arange(16).view(4, 4)).numpy()
    np.savez_compressed(os.path.join(OUT, "vicreg_cases.npz"), **blob)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
