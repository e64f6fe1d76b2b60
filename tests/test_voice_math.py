"""CPU: the per-sample arithmetic the CUDA voice kernels are built from (csrc/voice_math.cuh), compiled for the
host and checked against torch's CPU ops / the Voice oracle.  This is where the rounding contract is pinned:
which ops are reproduced bit for bit and which are only correctly rounded (see DESIGN.md "Voice numerics")."""
import ctypes

import pytest
import torch

from oracle import voice as V

P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
F = ctypes.c_float


def _bits(t):
    return t.contiguous().view(torch.int32)


def _neq(a, b):
    return int((_bits(a) != _bits(b)).sum())


def test_exp2_and_pow_are_bit_exact_restatements_of_torch_cpu(voice_shim):
    g = torch.Generator().manual_seed(5)
    n = 4_000_000 - 4_000_000 % (32 * torch.get_num_threads())  # whole SIMD vectors per thread chunk (tails use libm)
    x = ((torch.rand(n, generator=g) * 223 - 69) / 12).float()
    y = torch.empty_like(x)
    voice_shim.shim_exp2_fast(P(x), P(y), ctypes.c_long(n))
    assert _neq(y, torch.exp2(x)) == 0
    x = (torch.rand(n, generator=g) * -170 + 10).float()
    voice_shim.shim_exp2_full(P(x), P(y), ctypes.c_long(n))
    assert _neq(y, torch.exp2(x)) == 0
    r = torch.rand(n, generator=g)
    a = torch.rand(n, generator=g) * 5.9 + 0.1
    r[:8] = torch.tensor([0.0, 1.0, 1e-38, 1e-30, 1e-6, 0.99999994, 0.5, 0.75])
    voice_shim.shim_pow(P(r), P(a), P(y), ctypes.c_long(n))
    assert _neq(y, torch.pow(r, a)) == 0


def test_constant_division_is_ieee(voice_shim):
    g = torch.Generator().manual_seed(6)
    x = (torch.rand(4_000_000, generator=g) * 254 - 127).float()
    y = torch.empty_like(x)
    voice_shim.shim_div_const(P(x), F(12.0), P(y), ctypes.c_long(x.numel()))
    assert _neq(y, x / 12.0) == 0
    x = (torch.rand(4_000_000, generator=g) * 80000).float()
    voice_shim.shim_div_const(P(x), F(44100.0), P(y), ctypes.c_long(x.numel()))
    assert _neq(y, x / 44100.0) == 0
    voice_shim.shim_div_const(P(x), F(48000.0), P(y), ctypes.c_long(x.numel()))
    assert _neq(y, x / 48000.0) == 0


def test_division_by_per_voice_constant_is_ieee(voice_shim):
    """ADSR ramps divide by a duration that is constant per voice: rcp + two Markstein steps == IEEE division."""
    g = torch.Generator().manual_seed(8)
    n = 8_000_000
    a = (torch.rand(n, generator=g) * 14000).float() + 1e-6          # max(n - start, 0) + eps, up to 30 s of control points
    c = torch.exp(torch.rand(n, generator=g) * 30 - 14).float()       # durations 1e-6 .. 1e7 control samples
    c[:4] = torch.tensor([882.0, 2205.0, 1.9999999, 0.99999994])      # all-ones mantissas included
    a[n // 2:] = torch.randint(0, 13230, (n - n // 2,), generator=g).float() + 1e-6
    y = torch.empty_like(a)
    voice_shim.shim_div_pre(P(a), P(c), P(y), ctypes.c_long(n))
    assert _neq(y, a / c) == 0
    x = (torch.rand(n, generator=g) * 6.2831855).float()
    voice_shim.shim_div_const(P(x), F(6.2831854820251465), P(y), ctypes.c_long(n))
    assert _neq(y, x / 6.2831854820251465) == 0
    x = (torch.rand(n, generator=g) * 900).float()
    voice_shim.shim_div_pre(P(x), P(torch.full_like(x, 441.0)), P(y), ctypes.c_long(n))
    assert _neq(y, x / 441.0) == 0


def test_large_argument_sincos(voice_shim):
    """Phases reach 3e5 rad (4 s) / 2.4e6 rad (30 s): the reduction must keep relative accuracy near sin's zeros."""
    g = torch.Generator().manual_seed(7)
    x = (torch.rand(2_000_000, generator=g) * 2.4e6).float()
    s, c = torch.empty_like(x), torch.empty_like(x)
    voice_shim.shim_sincos_arg(P(x), P(s), P(c), ctypes.c_long(x.numel()))
    xs, xc = torch.sin(x.double()), torch.cos(x.double())
    assert float((s.double() - xs).abs().max()) < 3e-7 and float((c.double() - xc).abs().max()) < 3e-7
    # near the zero crossings of sin the error must stay tiny in absolute terms: SquareSawVCO feeds sin through
    # tanh with a gain of up to ~2500 (pi * partials / 2 for the lowest notes)
    near_zero = xs.abs() < 1e-2
    assert float((s.double() - xs).abs()[near_zero].max()) < 1e-8


@pytest.fixture(scope="module")
def rendered(voice_shim):
    B, T, C = 32, 176400, 1764  # multiple of the SIMD width: torch evaluates vector tails with scalar libm instead of SLEEF
    u = V.seeded_params(0, B)
    idx = {k: i for i, k in enumerate(V.sorted_keys())}
    reg = torch.stack([u[:, idx[k]] for k in V.registration_keys()], dim=0).contiguous()
    noise = V.noise_table(32, T)
    o = V.voice_render(u, noise, intermediates=True)
    ctrl, vconst, dbg = torch.empty(B, 5, C), torch.empty(B, 16), torch.empty(B, 8, C)
    voice_shim.shim_control(P(reg), B, C, F(441.0), F(1e-6), P(ctrl), P(vconst), P(dbg))
    return dict(B=B, T=T, C=C, u=u, reg=reg, noise=noise, o=o, ctrl=ctrl, vconst=vconst, dbg=dbg)


def test_parameter_scaling_bit_exact(voice_shim, rendered):
    B = rendered["B"]
    got = torch.empty(B, 78)
    voice_shim.shim_from_0to1(P(rendered["reg"]), B, P(got))
    p = V._P(rendered["u"])
    ref = torch.stack([p(m, n) for m, n in V.registration_keys()], dim=1)
    assert _neq(got, ref) <= 2  # log2 is correctly rounded, not a bit-exact restatement: 1-ulp differences are rare


def test_control_stage_against_oracle(rendered):
    o, dbg, ctrl = rendered["o"], rendered["dbg"], rendered["ctrl"]
    n = o["adsr"][:, 0].numel()
    for i in range(6):  # the six ADSRs: pow is a bit-exact restatement, so these agree except on SIMD-tail elements
        assert _neq(dbg[:, i], o["adsr"][:, i]) <= 1e-4 * n
    # LFOs go through cos (torch: MKL VML, closed source) -> correctly rounded stand-in, 1-ulp differences allowed
    assert float((dbg[:, 6:8] - o["lfo"]).abs().max()) <= 2.4e-7
    assert float((ctrl - o["ctrl"]).abs().max()) <= 2.4e-7
    assert _neq(ctrl, o["ctrl"]) <= 0.02 * ctrl.numel()


def test_audio_stage_is_bit_exact_given_control_signals(voice_shim, rendered):
    """With identical control-rate signals both VCO phase arguments match the torch CPU path bit for bit over all
    2 x B x 176400 samples (fp64 scan is exact for 4 s clips), so the audio agrees to ~2e-7."""
    r = rendered
    B, T, C = r["B"], r["T"], r["C"]
    audio, peak, phase = torch.empty(B, T), torch.empty(B), torch.empty(B, 2, T)
    voice_shim.shim_audio(P(r["o"]["ctrl"].contiguous()), P(r["vconst"]), P(r["noise"]), 32, B, T, C, F(44100.0), 1,
                          P(audio), P(peak), P(phase))
    assert _neq(phase[:, 0], r["o"]["arg1"]) == 0
    assert _neq(phase[:, 1], r["o"]["arg2"]) == 0
    assert float((audio - r["o"]["audio"]).abs().max()) <= 1e-6
    assert float((peak - r["o"]["peak"]).abs().max()) <= 1e-6


def test_end_to_end_host_render_against_oracle(voice_shim, rendered):
    r = rendered
    B, T, C = r["B"], r["T"], r["C"]
    audio, peak = torch.empty(B, T), torch.empty(B)
    voice_shim.shim_audio(P(r["ctrl"]), P(r["vconst"]), P(r["noise"]), 32, B, T, C, F(44100.0), 1, P(audio), P(peak),
                          None)
    err = (audio - r["o"]["audio"]).abs().max(dim=1)[0]
    # The fp32 oracle itself is ill-conditioned along the pitch path: a 1-ulp change of a control value
    # moves the audio by up to 1e-2.  Contract: most voices meet 1e-4; none is worse than the oracle's own distance
    # from its fp64 evaluation.
    o64 = V.voice_render(r["u"], r["noise"], dtype=torch.float64)["audio"].float()
    self_err = (r["o"]["audio"] - o64).abs().max(dim=1)[0]
    assert int((err <= 1e-4).sum()) >= int(0.7 * B)
    assert float(err.max()) <= max(float(self_err.max()), 1e-4)
