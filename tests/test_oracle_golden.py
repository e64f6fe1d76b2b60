"""CPU: the oracles against the golden vectors frozen from the reference itself (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import make_golden as MG
from oracle import pqmf as P
from oracle import vicreg as VR
from oracle import voice as V


@pytest.fixture(scope="module")
def filters():
    return np.load(os.path.join(GOLDEN, "pqmf_filters.npz"))


@pytest.fixture(scope="module")
def cases():
    return np.load(os.path.join(GOLDEN, "pqmf_cases.npz"))


@pytest.mark.parametrize("N,cutoff", [(3, 0.15), (4, 0.15), (16, 0.15), (16, 0.03), (2, 0.15), (8, 0.15)])
def test_pqmf_design_bit_equal(filters, N, cutoff):
    H, G = P.design(N, cutoff=cutoff)
    assert np.array_equal(H, filters[f"H_n{N}_c{cutoff}"])
    assert np.array_equal(G, filters[f"G_n{N}_c{cutoff}"])


def test_pqmf_known_answers(filters):
    H = filters["H_n3_c0.15"]  # SURVEY 8(c) known-answer facts
    assert np.allclose(H[0, :4], [-1.4490e-05, -2.6802e-05, 2.3541e-19, 6.1349e-05], rtol=1e-4, atol=1e-22)
    assert np.allclose(H.sum(axis=1), [0.29623, -2.5e-06, -5.1e-07], atol=2e-6)
    assert abs(np.abs(H).max() - 0.30000341) < 1e-7


@pytest.mark.parametrize("name,N,cutoff,B,T", MG.PQMF_CASES)
def test_pqmf_oracle_vs_reference(cases, name, N, cutoff, B, T):
    H, G = P.design(N, cutoff=cutoff)
    x = MG.pqmf_input(B, T).numpy()[:, 0, :]
    z = P.analysis(x, H, N)
    assert z.shape == cases[f"{name}_analysis"].shape == (B, N, P.out_len(T, N))
    assert P.rel_err(z, cases[f"{name}_analysis"]) <= 1e-5
    y = P.synthesis(cases[f"{name}_analysis"], G, N)
    assert y.shape == cases[f"{name}_synthesis"][:, 0, :].shape
    assert P.rel_err(y, cases[f"{name}_synthesis"][:, 0, :]) <= 1e-5


def test_pqmf_oracle_full_length_subsample(cases):
    H, G = P.design(3)
    x = MG.pqmf_input(2, 176400).numpy()[:, 0, :]
    z = P.analysis(x, H, 3)
    assert P.rel_err(z[:, :, ::MG.SUB], cases["n3_full_analysis_sub"]) <= 1e-5
    assert np.allclose(z.astype(np.float64).sum(axis=2), cases["n3_full_analysis_sum"], rtol=0, atol=2e-3)


def test_pqmf_reference_is_not_perfect_reconstruction(cases):
    """the reference filter bank reconstructs poorly; parity means equal to *its* output, not small error."""
    x = MG.pqmf_input(2, 4096).numpy()[:, 0, :]
    y = cases["n3_t4096_synthesis"][:, 0, :4096]
    lag = 1
    num = np.sqrt(np.mean((y[:, lag:] - x[:, :-lag]) ** 2))
    assert 0.2 < num / np.sqrt(np.mean(x ** 2)) < 0.7


@pytest.fixture(scope="module")
def vcases():
    return np.load(os.path.join(GOLDEN, "vicreg_cases.npz"))


@pytest.mark.parametrize("name,B,D,kind,cfgB,E", MG.VICREG_CASES)
def test_vicreg_oracle_vs_reference(vcases, name, B, D, kind, cfgB, E):
    x, y = MG.vicreg_inputs(B, D, kind)
    got = np.array(VR.loss(x.numpy(), y.numpy(), cfgB, E))
    ref = vcases[f"{name}_loss4"]
    assert np.all(np.abs(got - ref) <= 1e-5 * np.abs(ref) + 1e-9)
    gx, gy = VR.loss_grad(x.numpy(), y.numpy(), cfgB, E)
    if B <= 128:
        rx, ry = vcases[f"{name}_gx"], vcases[f"{name}_gy"]
    else:
        gx, gy, rx, ry = gx[::64], gy[::64], vcases[f"{name}_gx_sub"], vcases[f"{name}_gy_sub"]
    assert np.abs(gx - rx).max() <= 1e-5 * np.abs(rx).max()
    assert np.abs(gy - ry).max() <= 1e-5 * np.abs(ry).max()


def test_off_diagonal_known_answer(vcases):
    assert np.array_equal(VR.off_diagonal(np.arange(16).reshape(4, 4)), vcases["off_diagonal_4x4"])
    assert list(vcases["off_diagonal_4x4"]) == [1, 2, 3, 4, 6, 7, 8, 9, 11, 12, 13, 14]


def test_vicreg_local_rows_and_gather_semantics():
    """N-rank result == 1-rank result on the rank-ordered concatenation; repr term stays local (vicreg.py:36-39)."""
    x, y = MG.vicreg_inputs(128, 64, "correlated")
    x, y = x.numpy(), y.numpy()
    shards = [x[r * 32:(r + 1) * 32] for r in range(4)]
    assert np.array_equal(VR.full_gather_forward(shards), x)
    full = VR.loss(x, y, 128, 64)
    parts = [VR.loss(x, y, 128, 64, local_rows=slice(r * 32, (r + 1) * 32)) for r in range(4)]
    for p in parts:
        assert abs(p[2] - full[2]) < 1e-12 and abs(p[3] - full[3]) < 1e-12
    assert abs(np.mean([p[1] for p in parts]) - full[1]) < 1e-12
    g = [np.random.default_rng(r).standard_normal((128, 64)) for r in range(4)]
    own = VR.full_gather_backward(g, 2, 32)
    assert np.allclose(own, sum(g)[64:96])


# ---- Voice oracle (parity unpinned: structural checks only) -----------------------------------------------------
def test_voice_parameter_inventory():
    assert V.NPARAMS == 78  # conf/config.yaml:27
    s = V.sorted_keys()
    assert s[:5] == [("adsr_1", n) for n in ("alpha", "attack", "decay", "release", "sustain")]
    assert s[10:12] == [("keyboard", "duration"), ("keyboard", "midi_f0")]
    assert s[12][0] == "lfo_1" and s[20][0] == "lfo_1_amp_adsr"  # '.' sorts before '_'
    assert s[48:51] == [("mixer", "noise"), ("mixer", "vco_1"), ("mixer", "vco_2")]
    assert s[71:] == [("vco_1", "initial_phase"), ("vco_1", "mod_depth"), ("vco_1", "tuning"),
                      ("vco_2", "initial_phase"), ("vco_2", "mod_depth"), ("vco_2", "shape"), ("vco_2", "tuning")]
    assert len(set(V.registration_keys())) == 78 and set(V.registration_keys()) == set(s)
    assert V.get_parameters_keys()[0] == ("adsr_1", "attack")


def test_mt19937_restatement_matches_torch_rand():
    for seed in (0, 1, 12345, 2 ** 31 + 7, 2 ** 32 + 5):
        g = torch.Generator().manual_seed(seed)
        assert np.array_equal(V.mt19937_uniform24(seed, 78), torch.rand(78, generator=g).numpy())
    u = V.seeded_params(3, 4)
    assert np.array_equal(u[2].numpy(), V.mt19937_uniform24(3 * 4 + 2, 78))


def test_voice_oracle_shapes_and_ranges():
    cfg = V.SynthConfigO(batch_size=32, buffer_size_seconds=0.5)
    audio, params, is_train = V.voice_forward(7, cfg)
    assert audio.shape == (32, cfg.buffer_size) and params.shape == (32, 78) and is_train.shape == (32,)
    assert audio.abs().max() <= 1.0 + 1e-6  # normalize_if_clipping
    assert torch.isfinite(audio).all()
    assert bool(is_train.all())  # ids 224..255 -> (id//32)%10 == 7
    assert not bool(V.is_train(9, 32).any())  # ids 288..319 -> block 9 is the held-out one
    assert (params >= 0).all() and (params < 1).all()


def test_voice_oracle_has_not_drifted():
    """oracle/voice.py against the committed fixture of its own output (oracle/make_voice_fixture.py): the Voice oracle
    is unpinned w.r.t. torchsynth, but it cannot change silently."""
    from oracle import make_voice_fixture as F

    want = np.load(os.path.join(GOLDEN, "voice_oracle.npz"))
    got = F.compute()
    assert list(got["params_sha"]) == list(want["params_sha"])      # MT19937 + 24-bit mantissa: host independent
    assert list(got["noise_sha"]) == list(want["noise_sha"])
    assert np.array_equal(got["is_train"], want["is_train"])
    assert np.abs(got["adsr"] - want["adsr"]).max() <= 2e-7
    assert np.abs(got["ctrl"] - want["ctrl"]).max() <= 5e-7
    err = np.abs(got["audio"] - want["audio"]).max(axis=1)
    assert np.median(err) <= 1e-5 and np.abs(got["peak"] - want["peak"]).max() <= 1e-3
