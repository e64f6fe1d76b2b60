"""GPU parity: PQMF kernels (through the ias_b200.PQMF binding -> C ABI) against the oracle and the reference goldens.
Tolerance (north star): max|a-b| / max|b| <= 1e-5."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import make_golden as MG
from oracle import pqmf as OP

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def cases():
    return np.load(os.path.join(GOLDEN, "pqmf_cases.npz"))


def _mod(N, cutoff, dev, **kw):
    import ias_b200

    return ias_b200.PQMF(N=N, cutoff=cutoff, **kw).to(dev)


@pytest.mark.parametrize("name,N,cutoff,B,T", MG.PQMF_CASES)
def test_analysis_and_synthesis_vs_reference_goldens(cuda_device, cases, name, N, cutoff, B, T):
    m = _mod(N, cutoff, cuda_device)
    x = MG.pqmf_input(B, T).to(cuda_device)
    z = m(x)
    ref_z = cases[f"{name}_analysis"]
    assert tuple(z.shape) == ref_z.shape
    assert OP.rel_err(z.cpu().numpy(), ref_z) <= TOL
    y = m.synthesis(torch.from_numpy(ref_z).to(cuda_device))
    ref_y = cases[f"{name}_synthesis"]
    assert tuple(y.shape) == ref_y.shape
    assert OP.rel_err(y.cpu().numpy(), ref_y) <= TOL
    # and against the oracle restatement on the same seeded input
    H, G = OP.design(N, cutoff=cutoff)
    assert OP.rel_err(z.cpu().numpy(), OP.analysis(x.cpu().numpy()[:, 0, :], H, N)) <= TOL


@pytest.mark.parametrize("name,N", [("n3_full", 3), ("n16_full", 16)])
def test_full_length_vs_reference_subsample(cuda_device, cases, name, N):
    m = _mod(N, 0.15, cuda_device)
    x = MG.pqmf_input(2, 176400).to(cuda_device)
    z = m.analysis(x)
    assert z.shape == (2, N, (176400 - 1) // N + 1)
    assert OP.rel_err(z[:, :, ::MG.SUB].cpu().numpy(), cases[f"{name}_analysis_sub"]) <= TOL
    assert np.allclose(z.double().sum(dim=2).cpu().numpy(), cases[f"{name}_analysis_sum"], rtol=0, atol=5e-3)
    assert np.allclose(z.double().abs().sum(dim=2).cpu().numpy(), cases[f"{name}_analysis_abssum"], rtol=1e-6)
    y = m.synthesis(z)
    assert y.shape == (2, 1, 176400)
    # the golden synthesis ran on the reference's own analysis output; ours differs from that by <= 1e-5 relative
    assert OP.rel_err(y[:, :, ::MG.SUB].cpu().numpy(), cases[f"{name}_synthesis_sub"]) <= 5 * TOL
    assert np.allclose(y.double().abs().sum(dim=2).cpu().numpy(), cases[f"{name}_synthesis_abssum"], rtol=1e-5)


@pytest.mark.parametrize("N,taps", [(5, 62), (3, 30), (7, 41), (4, 63)])
def test_generic_shapes_vs_oracle(cuda_device, N, taps):
    """Shapes without a specialised kernel (any N, any tap count, even K) take the generic path."""
    import ias_b200

    m = ias_b200.PQMF(N=N, taps=taps).to(cuda_device)
    H = m.H[:, 0, :].cpu().numpy()
    G = m.G[0].cpu().numpy()
    x = MG.pqmf_input(2, 3001, seed=3).to(cuda_device)
    z = m(x)
    K = taps + 1
    if K % 2 == 1:  # the numpy oracle covers odd K (the reference's default); even K is checked by a direct sum
        assert OP.rel_err(z.cpu().numpy(), OP.analysis(x.cpu().numpy()[:, 0, :], H, N)) <= TOL
        y = m.synthesis(z)
        assert OP.rel_err(y.cpu().numpy()[:, 0, :], OP.synthesis(z.cpu().numpy(), G, N)) <= TOL
    else:
        ref = torch.nn.functional.conv1d(x.cpu().double(), m.H.cpu().double(), padding=taps // 2, stride=N)
        assert tuple(z.shape) == tuple(ref.shape)
        assert OP.rel_err(z.cpu().numpy(), ref.float().numpy()) <= TOL


def test_baseline_size_properties(cuda_device):
    """Config 3 size (1024 x 4 s, N = 16 and 3): size-independent properties + oracle on a row subset."""
    B, T = 1024, 176400
    g = torch.Generator(device="cpu").manual_seed(11)
    x = (torch.rand((B, 1, T), generator=g) * 2 - 1).to(cuda_device)
    for N in (16, 3):
        m = _mod(N, 0.15, cuda_device)
        z = m(x)
        # rows are independent: a row filtered alone is bit-identical to the same row inside the batch
        for r in (0, 517, 1023):
            assert torch.equal(m(x[r:r + 1]), z[r:r + 1])
        # linearity
        a, b2 = 0.75, -1.5
        z2 = m(a * x[:8] + b2 * x[8:16])
        lin = a * z[:8] + b2 * z[8:16]
        assert float((z2 - lin).abs().max() / lin.abs().max()) <= TOL
        # row_scale folds a per-row gain in (normalize_if_clipping folded into the consumer)
        s = torch.rand(B, device=cuda_device) + 0.5
        zs = m.analysis(x, row_scale=s)
        assert float((zs - z * s[:, None, None]).abs().max() / z.abs().max()) <= TOL
        # oracle on two rows, and the reference's own reconstruction behaviour (not small!)
        H, G = OP.design(N)
        rows = [3, 1000]
        assert OP.rel_err(z[rows].cpu().numpy(), OP.analysis(x[rows, 0].cpu().numpy(), H, N)) <= TOL
        y = m.synthesis(z)
        assert y.shape == (B, 1, T)
        assert OP.rel_err(y[rows, 0].cpu().numpy(), OP.synthesis(z[rows].cpu().numpy(), G, N)) <= TOL


def test_long_clip_and_ragged_lengths(cuda_device):
    m = _mod(16, 0.15, cuda_device)
    x = MG.pqmf_input(2, 1323000, seed=5).to(cuda_device)  # 30 s
    z = m(x)
    assert z.shape == (2, 16, 82688)
    y = m.synthesis(z)
    assert y.shape == (2, 1, 1323008)
    H, G = OP.design(16)
    assert OP.rel_err(z[:1].cpu().numpy(), OP.analysis(x[:1, 0].cpu().numpy(), H, 16)) <= TOL
    assert OP.rel_err(y[:1, 0].cpu().numpy(), OP.synthesis(z[:1].cpu().numpy(), G, 16)) <= TOL
    m3 = _mod(3, 0.15, cuda_device)
    for T in (1, 2, 31, 32, 62, 63, 64, 95, 1025, 3073, 6145):
        xs = MG.pqmf_input(3, T, seed=T).to(cuda_device)
        zs = m3(xs)
        H3, G3 = OP.design(3)
        ref = OP.analysis(xs[:, 0].cpu().numpy(), H3, 3)
        assert zs.shape == ref.shape
        assert np.abs(zs.cpu().numpy() - ref).max() <= TOL * max(np.abs(ref).max(), 1e-3)
        ys = m3.synthesis(zs)
        refy = OP.synthesis(zs.cpu().numpy(), G3, 3)
        assert np.abs(ys[:, 0].cpu().numpy() - refy).max() <= TOL * max(np.abs(refy).max(), 1e-3)


@pytest.mark.parametrize("N", [2, 3, 4, 8, 16])
def test_polyphase_and_direct_forms_agree(cuda_device, N):
    """Designed filter -> polyphase kernel; same taps with the fast path disabled -> direct kernel; both vs oracle."""
    m = _mod(N, 0.15, cuda_device)
    x = MG.pqmf_input(3, 20011, seed=N).to(cuda_device)
    zp = m(x)
    m.polyphase = False
    zd = m(x)
    H, _ = OP.design(N)
    ref = OP.analysis(x[:, 0].cpu().numpy(), H, N)
    assert OP.rel_err(zp.cpu().numpy(), ref) <= TOL and OP.rel_err(zd.cpu().numpy(), ref) <= TOL
    assert OP.rel_err(zp.cpu().numpy(), zd.cpu().numpy()) <= 2e-6
    # synthesis: cosine-modulated kernel (N = 8, 16) against the direct form and the oracle, ragged band lengths
    _, G = OP.design(N)
    for L in (1, 2, 3, 5, 124, 125, 126, 250, 1251, 4099):
        zz = MG.pqmf_input(3, N * L, seed=100 + L).reshape(3, N, L).to(cuda_device)
        m.polyphase = True
        yp = m.synthesis(zz)
        m.polyphase = False
        yd = m.synthesis(zz)
        refy = OP.synthesis(zz.cpu().numpy(), G, N)
        assert yp.shape == (3, 1, L * N)
        scale = max(np.abs(refy).max(), 1e-3)
        assert np.abs(yp[:, 0].cpu().numpy() - refy).max() <= TOL * scale
        assert np.abs(yd[:, 0].cpu().numpy() - refy).max() <= TOL * scale


def test_checkpoint_loaded_filter_is_used(cuda_device):
    """H/G are persistent buffers (state-dict keys gram.H ...): values loaded from a checkpoint must be the taps used."""
    import ias_b200

    m = ias_b200.PQMF(N=3).to(cuda_device)
    x = MG.pqmf_input(1, 4096).to(cuda_device)
    z0 = m(x)
    sd = m.state_dict()
    sd["H"] = sd["H"] * 2.0
    m.load_state_dict(sd)
    assert float((m(x) - 2 * z0).abs().max()) <= 1e-6 * float(z0.abs().max())


def test_image_epilogue_matches_reshape_plus_normalize(cuda_device):
    """AudioEmbedding._preprocess (audioembed.py:36-49): gram(audio).reshape(-1,3,240,245) then Normalize(mean, std)
    (vicreg_audio_params.py:60-62).  The fused epilogue applies torch's own fp32 sub/div, so it is bit-identical to
    running them on the unfused bands."""
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    m = _mod(3, 0.15, cuda_device)
    x = MG.pqmf_input(4, 176400).to(cuda_device)
    z = m(x).reshape(-1, 3, 240, 245)
    want = (z - torch.tensor(mean, device=cuda_device).view(1, 3, 1, 1)) / torch.tensor(std, device=cuda_device).view(1, 3, 1, 1)
    got = m.analysis_image(x, mean, std, image_shape=(240, 245))
    assert got.shape == (4, 3, 240, 245)
    assert torch.equal(got, want)
    # a shape without a specialised kernel (N=5) goes through the generic kernel's epilogue
    m5 = _mod(5, 0.1, cuda_device)
    x5 = MG.pqmf_input(2, 5000).to(cuda_device)
    mean5, std5 = [0.1, -0.2, 0.3, 0.0, 1.5], [0.5, 2.0, 0.25, 1.0, 3.0]
    z5 = m5(x5)
    want5 = (z5 - torch.tensor(mean5, device=cuda_device).view(1, 5, 1)) / torch.tensor(std5, device=cuda_device).view(1, 5, 1)
    assert torch.equal(m5.analysis_image(x5, mean5, std5), want5)
    with pytest.raises(ValueError):
        m.analysis_image(x, mean[:2], std[:2])


@pytest.mark.parametrize("N,B,T,P", [(3, 5, 176400, 256), (16, 3, 176400, 256), (3, 2, 100003, 64), (8, 2, 44100, 16),
                                       (3, 1024, 176400, 256)])
def test_pooled_epilogue_matches_pooling_the_bands(cuda_device, N, B, T, P):
    """Harness bridge: adaptive_avg_pool1d(|bands|.reshape(B,1,N*L), P) fused into the analysis kernel.
    Bands are bit-identical to the unfused call; features agree with torch's pooling of those bands (fp32 sums in a
    different order: 1e-5 relative) and are reproducible run to run."""
    import harness

    m = _mod(N, 0.15, cuda_device)
    g = torch.Generator(device="cpu").manual_seed(N * 1000 + P)
    x = (torch.rand((B, 1, T), generator=g) * 2 - 1).to(cuda_device)
    x[0, 0, : T // 3] = 0.0  # a silent stretch: bins that are exactly zero
    z = m(x)
    z2, feat = m.analysis_pooled(x, P)
    assert torch.equal(z, z2)
    rows = slice(0, min(B, 16))
    want = torch.nn.functional.adaptive_avg_pool1d(z[rows].double().abs().reshape(z[rows].shape[0], 1, -1).cpu(), P).squeeze(1)
    got = feat[rows].double().cpu()
    assert float((got - want).abs().max() / want.abs().max()) <= TOL
    _, feat_again = m.analysis_pooled(x, P)
    assert torch.equal(feat, feat_again)
    # the stand-alone pooling kernel of the harness agrees too
    if P == harness.EMBED_DIM:
        wa, wp = harness.bridge_weights(cuda_device)
        params = torch.zeros((B, harness.NPARAMS), device=cuda_device)
        xa, _ = harness.bridge(z, params, wa, wp)
        _, xb, _ = harness.analysis_bridge(m, x, params, wa, wp)
        assert float((xa - xb).abs().max() / xa.abs().max()) <= TOL


def test_pooled_epilogue_refuses_shapes_it_cannot_cover(cuda_device):
    import ias_b200

    m4 = _mod(4, 0.15, cuda_device)  # 1024-step CTA tiles: wider than a 689-element bin
    x = MG.pqmf_input(1, 176400).to(cuda_device)
    with pytest.raises(ias_b200.IasError):
        m4.analysis_pooled(x, 256)
    m5 = _mod(5, 0.15, cuda_device)  # no specialised kernel
    with pytest.raises(ias_b200.IasError):
        m5.analysis_pooled(x, 16)


@pytest.mark.parametrize("N", [3, 16])
def test_foreign_cosine_modulated_bank_with_stale_factors_takes_the_direct_form(cuda_device, N):
    """C-ABI callers may pass any filter with a (proto, mod) factorisation.  The fast kernels rebuild the modulation
    from PQMF.__init__'s design, so a bank they do not describe -- here the textbook taps/2 centring of the reference's
    own TODO (pqmf.py:26), passed together with the factors of the default design -- must fall back to the direct form
    instead of silently filtering with the wrong taps (ias_b200.h, ias_pqmf_analysis / ias_pqmf_synthesis)."""
    import ias_b200
    from ias_b200 import _lib
    from ias_b200.pqmf import cosine_modulation_factors
    from scipy import signal as sig

    taps, K = 62, 63
    proto = sig.firwin(K, 0.15, window=("kaiser", 9.0))
    k = np.arange(N)[:, None]
    j = np.arange(K)[None, :]
    theta = (2 * k + 1) * (np.pi / (2 * N)) * (j - taps / 2)   # textbook centring, NOT the reference's (taps-1)/2
    phase = ((-1.0) ** k) * np.pi / 4
    H = (2 * proto * np.cos(theta + phase)).astype(np.float32)
    G = (2 * proto * np.cos(theta - phase)).astype(np.float32)
    g, c = cosine_modulation_factors(N, taps, 0.15, 9.0)        # factors of the DEFAULT design: stale for H, G
    x = MG.pqmf_input(2, 9000, seed=5).to(cuda_device)
    lib = ias_b200.lib()
    L = lib.ias_pqmf_out_len(9000, N, K)
    Hh, Gh = torch.from_numpy(H).contiguous(), torch.from_numpy(G).contiguous()
    gh, ch = torch.from_numpy(g).contiguous(), torch.from_numpy(c).contiguous()
    z = torch.empty((2, N, L), device=cuda_device)
    _lib.check(lib.ias_pqmf_analysis(_lib.ptr(x), _lib.ptr(Hh.to(cuda_device)), _lib.ptr(Hh), _lib.ptr(gh), _lib.ptr(ch),
                                     None, _lib.ptr(z), 2, 9000, N, K, _lib.current_stream(cuda_device)))
    assert OP.rel_err(z.cpu().numpy(), OP.analysis(x.cpu().numpy()[:, 0, :], H.astype(np.float64), N)) <= TOL
    y = torch.empty((2, L * N), device=cuda_device)
    _lib.check(lib.ias_pqmf_synthesis(_lib.ptr(z), _lib.ptr(Gh.to(cuda_device)), _lib.ptr(Gh), _lib.ptr(gh), _lib.ptr(y),
                                      2, L, N, K, _lib.current_stream(cuda_device)))
    assert OP.rel_err(y.cpu().numpy(), OP.synthesis(z.cpu().numpy(), G.astype(np.float64), N)) <= TOL


@pytest.mark.parametrize("N", [3, 4])
def test_packed_small_n_synthesis_is_bit_identical_to_the_scalar_kernel(cuda_device, monkeypatch, N):
    """k_pqmf_synthesis_n3p / _n4p (FFMA2 FIR phase, default for N = 3 and N = 4) run the same multiply-add chains in the
    same order as k_pqmf_synthesis_small<N>: equal bit for bit (torch.equal treats +-0 alike), interior and edge tiles,
    ragged lengths, every Q."""
    m = _mod(N, 0.15, cuda_device)
    _, G = OP.design(N)
    for B, L in ((2, 1), (2, 7), (3, 341), (2, 1024), (2, 1025), (2, 4099), (3, 58800), (1, 330750)):
        zz = MG.pqmf_input(B, N * L, seed=300 + L % 97).reshape(B, N, L).to(cuda_device)
        outs = {}
        for packed in ("1", "0"):
            for q in ("8", "4"):
                monkeypatch.setenv("IAS_PQMF_SYNTH_PACKED", packed)
                monkeypatch.setenv("IAS_PQMF_SYNTH_Q", q)
                outs[packed, q] = m.synthesis(zz)
        monkeypatch.delenv("IAS_PQMF_SYNTH_PACKED")
        if N == 3:
            monkeypatch.setenv("IAS_PQMF_SYNTH_Q", "16")
            assert torch.equal(outs["1", "8"], m.synthesis(zz))
        monkeypatch.delenv("IAS_PQMF_SYNTH_Q")
        assert torch.equal(outs["1", "8"], outs["0", "8"]) and torch.equal(outs["1", "4"], outs["0", "4"])
        assert torch.equal(outs["1", "8"], outs["1", "4"])
        assert torch.equal(outs["1", "8"], m.synthesis(zz))  # packed is what runs by default
        if L <= 4099:
            refy = OP.synthesis(zz.cpu().numpy(), G, N)
            assert np.abs(outs["1", "8"][:, 0].cpu().numpy() - refy).max() <= TOL * max(np.abs(refy).max(), 1e-3)


@pytest.mark.parametrize("N", [8, 16])
def test_two_step_cosine_modulated_synthesis_is_bit_identical_to_the_one_step_kernel(cuda_device, monkeypatch, N):
    """k_pqmf_synthesis_cm2 (two same-parity steps per thread, reversed row halves, FFMA2 FIR phase; default for
    N = 8, 16) applies the same taps in the same order per output as k_pqmf_synthesis_cm."""
    m = _mod(N, 0.15, cuda_device)
    _, G = OP.design(N)
    for B, L in ((2, 1), (2, 3), (3, 119), (2, 120), (2, 121), (3, 124), (2, 125), (2, 1251), (3, 11025), (1, 82688)):
        zz = MG.pqmf_input(B, N * L, seed=500 + L % 89).reshape(B, N, L).to(cuda_device)
        monkeypatch.setenv("IAS_PQMF_SYNTH_CM2", "0")
        y1 = m.synthesis(zz)
        monkeypatch.setenv("IAS_PQMF_SYNTH_CM2", "1")
        y2 = m.synthesis(zz)
        monkeypatch.delenv("IAS_PQMF_SYNTH_CM2")
        assert torch.equal(y1, y2) and torch.equal(y2, m.synthesis(zz))
        if L <= 1251:
            refy = OP.synthesis(zz.cpu().numpy(), G, N)
            assert np.abs(y2[:, 0].cpu().numpy() - refy).max() <= TOL * max(np.abs(refy).max(), 1e-3)
