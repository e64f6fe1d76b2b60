"""CPU: the C-ABI library loads and exports every symbol include/ias_b200.h declares; host-side mirrors behave like
the reference surfaces (no compute calls: those need a GPU and live in the -m gpu tests)."""
import ctypes
import os
import re
import subprocess
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from oracle import voice as V


def _declared():
    text = open(os.path.join(ROOT, "include", "ias_b200.h")).read()
    return sorted(set(re.findall(r"IAS_API\s+[\w\s\*]+?\b(ias_\w+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    import ias_b200

    declared = _declared()
    assert "ias_voice_render" in declared and "ias_vicreg_loss" in declared and "ias_pqmf_analysis" in declared
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (ias_\w+)", out))
    core = [s for s in declared if not s.startswith("ias_comm_")]
    assert sorted(exported) == sorted(core)
    assert sorted(ias_b200._lib.CORE_SYMBOLS) == sorted(core)
    comm = [s for s in declared if s.startswith("ias_comm_")]
    assert sorted(set(ias_b200._lib.COMM_SYMBOLS) | {"ias_comm_last_error"}) == sorted(comm)
    if os.path.exists(ias_b200._lib.COMM_LIB_PATH):
        out = subprocess.run(["nm", "-D", "--defined-only", ias_b200._lib.COMM_LIB_PATH], capture_output=True,
                             text=True, check=True).stdout
        assert sorted(set(re.findall(r" T (ias_\w+)", out))) == sorted(comm)


def test_library_loads_and_reports_errors(built_lib):
    import ias_b200

    lib = ias_b200.lib()
    assert lib.ias_version() >= 100
    # argument validation happens before any CUDA call, so it is testable without a GPU
    rc = lib.ias_pqmf_analysis(None, None, None, None, None, None, None, 0, 10, 3, 63, None)
    assert rc == 1 and b"ias_pqmf_analysis" in lib.ias_last_error()
    rc = lib.ias_voice_render(None, None, 0, None, None, 1, 100, 10, 44100.0, 441.0, 1e-6, 1, None, None, None, 0, None)
    assert rc == 1
    rc = lib.ias_vicreg_loss(None, None, 0, 0, 0, 2, 0, 1, 25.0, 25.0, 1.0, None, None, 0, None)
    assert rc == 1
    assert lib.ias_pqmf_out_len(176400, 3, 63) == 58800
    assert lib.ias_pqmf_out_len(176400, 16, 63) == 11025
    assert lib.ias_pqmf_out_len(1323000, 16, 63) == 82688
    assert lib.ias_pqmf_out_len(1, 3, 63) == 1
    # records [B][C][16] + ctrl [B][5][C] + scratch [B][6][C] + constants [B][16] + schedule (2B + 4 ints)
    assert lib.ias_voice_workspace_bytes(1024, 176400, 1764) == 4 * (1024 * 27 * 1764 + 1024 * 16 + 2 * 1024 + 4)
    assert lib.ias_vicreg_workspace_bytes(8192, 256) > 2 * 2 * 8192 * 256 * 4


def test_new_entry_points_validate_arguments_without_a_gpu(built_lib):
    """ias_voice_seed_params_dev / ias_pqmf_analysis_image reject bad arguments before touching CUDA; the Python
    surfaces refuse CPU tensors (no fallback)."""
    import ias_b200

    lib = ias_b200.lib()
    assert lib.ias_voice_seed_params_dev(None, 32, None, None, None, None) == 1
    assert b"batch_idx_dev" in lib.ias_last_error()
    rc = lib.ias_pqmf_analysis_image(None, None, None, None, None, None, None, None, None, None, 4, 100, 3, 63, None)
    assert rc == 1 and b"mean/std" in lib.ias_last_error()
    gram = ias_b200.PQMF(N=3)
    with pytest.raises(ias_b200.IasError):
        gram.analysis_image(torch.zeros(2, 1, 1000), [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
    with pytest.raises(ValueError):
        gram.analysis_image(torch.zeros(2, 1, 1000), [0.5], [0.5])
    voice = ias_b200.Voice(synthconfig=ias_b200.SynthConfig(batch_size=32, reproducible=True, buffer_size_seconds=0.1))
    with pytest.raises(ias_b200.IasError):
        voice(0)  # parameters live on the CPU until .to(device): no CPU render path exists


def test_round3_entry_points_validate_arguments_without_a_gpu(built_lib):
    """ias_pqmf_analysis_pooled / ias_voice_render_stages / the prototype argument of ias_pqmf_synthesis: argument
    checks run before any CUDA call; the Python mirrors refuse CPU tensors."""
    import ias_b200

    lib = ias_b200.lib()
    rc = lib.ias_pqmf_analysis_pooled(None, None, None, None, None, None, None, None, 256, None, 0, 4, 1000, 3, 63, None)
    assert rc == 1 and b"ias_pqmf_analysis_pooled" in lib.ias_last_error()
    # workspace: two floats per (sound, band, CTA tile) of the smallest tile any kernel uses (256 steps)
    assert lib.ias_pqmf_pool_workspace_bytes(1024, 176400, 3, 63) == 1024 * 3 * ((58800 + 255) // 256) * 2 * 4
    assert lib.ias_pqmf_pool_workspace_bytes(0, 176400, 3, 63) == 0
    rc = lib.ias_voice_render_stages(None, None, 0, None, None, 1, 100, 10, 44100.0, 441.0, 1e-6, 1, None, None, None, 0,
                                     0, None)
    assert rc == 1 and b"stages" in lib.ias_last_error()
    rc = lib.ias_voice_render_stages(None, None, 0, None, None, 1, 1000, 10, 44100.0, 441.0, 1e-6, 1, None, None, None,
                                     0, 1, None)  # control stage only: params01 is still required
    assert rc == 1
    rc = lib.ias_pqmf_synthesis(None, None, None, None, None, 0, 10, 16, 63, None)
    assert rc == 1 and b"ias_pqmf_synthesis" in lib.ias_last_error()
    gram = ias_b200.PQMF(N=3)
    with pytest.raises(ias_b200.IasError):
        gram.analysis_pooled(torch.zeros(2, 1, 200000), 256)
    with pytest.raises(ValueError):
        gram.analysis_pooled(torch.zeros(2, 200000), 256)
    voice = ias_b200.Voice(synthconfig=ias_b200.SynthConfig(batch_size=32, reproducible=True, buffer_size_seconds=0.1))
    with pytest.raises(ias_b200.IasError):
        voice.prepare(0)
    with pytest.raises(RuntimeError):
        voice(0, prepared=True)  # nothing was prepared
    with pytest.raises(ValueError):
        voice.prepare(None)


def test_voice_tables_match_oracle(built_lib):
    import ias_b200

    lib = ias_b200.lib()
    reg = V.registration_keys()
    srt = {k: i for i, k in enumerate(V.sorted_keys())}
    for i, (m, n) in enumerate(reg):
        assert lib.ias_voice_param_name(i).decode() == f"{m}/{n}"
        assert lib.ias_voice_sorted_index(i) == srt[(m, n)]
    assert lib.ias_voice_param_name(78) is None and lib.ias_voice_sorted_index(-1) == -1


def test_voice_module_surface(built_lib):
    """The torchsynth API the reference calls (audio_to_params.py:238-257; vicreg_audio_params.py:86-94)."""
    import ias_b200

    cfg = ias_b200.SynthConfig(batch_size=32, reproducible=False, sample_rate=44100, buffer_size_seconds=4.0)
    assert (cfg.buffer_size, cfg.control_buffer_size, cfg.control_rate, cfg.eps, cfg.no_grad) == (176400, 1764, 441,
                                                                                                  1e-6, True)
    with pytest.raises(ValueError):
        ias_b200.SynthConfig(batch_size=33, reproducible=True)
    voice = ias_b200.Voice(synthconfig=cfg)
    assert voice.batch_size == 32
    keys = list(voice.get_parameters().keys())
    assert keys == V.get_parameters_keys() and len(keys) == 78
    assert [n for n, _ in voice.named_parameters()] == [f"{m}.torchparameters.{p}" for m, p in V.registration_keys()]
    sd = voice.state_dict()
    assert "noise.noise" in sd and sd["noise.noise"].shape == (32, 176400)
    assert "adsr_1.range" in sd and "vco_2.torchparameters.shape" in sd
    assert sum(k.endswith(tuple(p for _, p in V.registration_keys())) and ".torchparameters." in k for k in sd) == 78
    # noise table == torchsynth Noise(seed=13) restatement
    assert torch.equal(sd["noise.noise"][:2], V.noise_table(32, 176400)[:2])
    # set_parameter_0to1 writes through to the shared [78,B] block, params01() is registration order
    value = torch.linspace(0, 0.9, 32)
    voice.vco_1.set_parameter_0to1("tuning", value)
    row = V.registration_keys().index(("vco_1", "tuning"))
    assert torch.equal(voice.params01()[:, row], value)
    assert torch.allclose(voice.vco_1.p("tuning"), -24 + 48 * value)
    # a user assigning parameter.data (as torchsynth.randomize does) is picked up at the next tie
    voice.mixer.torchparameters["noise"].data = torch.full((32,), 0.25)
    assert torch.equal(voice.params01()[:, 77], torch.full((32,), 0.25))
    # freeze / unfreeze
    voice.freeze_parameters([("vco_1", "tuning"), ("mixer", "noise")])
    assert len(voice.get_parameters()) == 76 and len(voice.get_parameters(include_frozen=True)) == 78
    frozen = voice._frozen_rows()
    assert sum(frozen) == 2 and frozen[row] == 1 and frozen[77] == 1
    voice.unfreeze_all_parameters()
    assert len(voice.get_parameters()) == 78
    # state dict round trip keeps the views tied
    other = ias_b200.Voice(synthconfig=cfg)
    other.load_state_dict(sd)
    assert torch.equal(other.params01(), voice.params01())
    # symmetric range helpers invert each other
    u = torch.rand(100)
    for r in (voice.lfo_1.torchparameters["mod_depth"].parameter_range,   # symmetric, curve 0.5
              voice.adsr_1.torchparameters["attack"].parameter_range):    # plain, curve 0.5
        assert torch.allclose(r.to_0to1(r.from_0to1(u)), u, atol=2e-3)
    # no CPU fallback
    with pytest.raises(ias_b200.IasError):
        voice(0)
    reproducible = ias_b200.Voice(ias_b200.SynthConfig(batch_size=32, reproducible=True, buffer_size_seconds=0.1))
    with pytest.raises(ValueError):
        reproducible(None)


def test_pqmf_module_surface(built_lib):
    import ias_b200

    filters = np.load(os.path.join(GOLDEN, "pqmf_filters.npz"))
    for N, cutoff in [(3, 0.15), (4, 0.15), (16, 0.15), (16, 0.03)]:
        m = ias_b200.PQMF(N=N, cutoff=cutoff)
        assert (m.N, m.taps, m.cutoff, m.beta) == (N, 62, cutoff, 9.0)
        assert m.H.shape == (N, 1, 63) and m.G.shape == (1, N, 63) and m.updown_filter.shape == (N, N, N)
        assert np.array_equal(m.H[:, 0, :].numpy(), filters[f"H_n{N}_c{cutoff}"])  # bit-equal to the reference buffers
        assert np.array_equal(m.G[0].numpy(), filters[f"G_n{N}_c{cutoff}"])
        assert list(m.state_dict().keys()) == ["H", "G", "updown_filter"]
        assert float(m.updown_filter.sum()) == N and float(m.updown_filter[1, 1, 0]) == 1.0
    assert ias_b200.PQMF().N == 4
    # the polyphase factorisation offered to the kernel reproduces H to fp32 rounding
    from ias_b200.pqmf import cosine_modulation_factors, design_filters
    for N in (2, 3, 4, 8, 16):
        g, c = cosine_modulation_factors(N, 62, 0.15, 9.0)
        H, _ = design_filters(N, 62, 0.15, 9.0)
        rebuilt = g[None, :].astype(np.float64) * c[:, np.arange(63) % (2 * N)].astype(np.float64)
        assert np.abs(rebuilt - H).max() <= 2e-7 * np.abs(H).max()
    m = ias_b200.PQMF(N=3)
    assert m._taps("H")[2] is not None
    m.H.mul_(2.0)  # taps that are no longer the designed filter: only the direct form may run
    assert m._taps("H")[2] is None
    with pytest.raises(ias_b200.IasError):
        ias_b200.PQMF(N=3)(torch.zeros(1, 1, 100))
    with pytest.raises(ValueError):
        ias_b200.PQMF(N=3).analysis(torch.zeros(1, 100))


def test_vicreg_module_surface(built_lib):
    import ias_b200

    cfg = types.SimpleNamespace(dim=16, embeddim=32, vicreg=types.SimpleNamespace(
        mlp="24-24-%d", batch_size=8, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0))
    m = ias_b200.VICReg(cfg, torch.nn.Identity(), torch.nn.Identity())
    assert m.reprdim == 16 and m.embeddim == 32
    shapes = [tuple(p.shape) for p in m.projector.parameters()]
    assert shapes == [(24, 16), (24,), (24,), (24,), (24, 24), (24,), (24,), (24,), (32, 24)]
    x, y = m(torch.randn(8, 16), torch.randn(8, 16))
    assert x.shape == (8, 32) and y.shape == (8, 32)
    assert list(ias_b200.off_diagonal(torch.arange(16).view(4, 4))) == [1, 2, 3, 4, 6, 7, 8, 9, 11, 12, 13, 14]
    with pytest.raises(ias_b200.IasError):
        m.loss(x, y)
    out = ias_b200.FullGatherLayer.apply(x)  # world size 1: identity tuple
    assert isinstance(out, tuple) and len(out) == 1 and torch.equal(out[0], x)
