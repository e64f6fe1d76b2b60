"""ORACLE (test infrastructure, not product code): CPU torch restatement of torchsynth's ``Voice``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  The product path (``ias_b200``) never does.

PARITY UNPINNED.  The synth arithmetic lives in the third-party package ``torchsynth`` (PyPI, *unpinned*
in the reference's ``requirements.txt:1``; latest published line 1.0.x).  It is not vendored in
``/root/reference`` and not installed in this image, and the reference has no test, fixture or golden
vector that touches synth output.  This file restates the published torchsynth v1.0.x algorithm
(``synth.py``, ``module.py``, ``parameter.py``, ``config.py``, ``util.py``) following SURVEY.md
Appendix A, anchored on the reference's call sites:

  * ``vicreg_audio_params.py:86-94,114``  ``Voice(synthconfig)(batch_idx) -> (audio, params, is_train)``
  * ``audio_to_params.py:196-203,215,238-257``  parameter API (``get_parameters`` order, ``voice(None)``)
  * ``conf/config.yaml:27``  ``nparams: 78`` (structural check: this restatement has exactly 78)
  * ``pretrain.py:72`` / ``heareval.py:15``  audio is ``[B, int(4.0 * 44100)]``

Where torchsynth's exact expression is uncertain the restatement *defines* the behaviour (marked DEFINES).
Every tensor op is written in the op order torchsynth uses so that the fp32 rounding sequence of the
CPU path is reproduced; ``dtype=torch.float64`` gives the high-precision evaluation of the same graph.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

PI = math.pi  # torch.pi is the same Python float

# --------------------------------------------------------------------------------------------------
# Parameter inventory (App. A.1, A.3).  (name, min, max, curve, symmetric) in *declaration* order.
# --------------------------------------------------------------------------------------------------
_KEYBOARD = [("midi_f0", 0.0, 127.0, 1.0, False), ("duration", 0.01, 4.0, 0.5, False)]
_ADSR = [
    ("attack", 0.0, 2.0, 0.5, False),
    ("decay", 0.0, 2.0, 0.5, False),
    ("sustain", 0.0, 1.0, 1.0, False),
    ("release", 0.0, 5.0, 0.5, False),
    ("alpha", 0.1, 6.0, 1.0, False),
]
_LFO_TYPES = ["sin", "tri", "saw", "rsaw", "sqr"]
_LFO = [
    ("frequency", 0.0, 20.0, 0.25, False),
    ("mod_depth", -10.0, 20.0, 0.5, True),
    ("initial_phase", -PI, PI, 1.0, False),
] + [(n, 0.0, 1.0, 1.0, False) for n in _LFO_TYPES]
_VCO = [
    ("tuning", -24.0, 24.0, 1.0, False),
    ("mod_depth", -96.0, 96.0, 0.2, True),
    ("initial_phase", -PI, PI, 1.0, False),
]
_SQUARESAW = _VCO + [("shape", 0.0, 1.0, 1.0, False)]
MOD_INPUTS = ["adsr_1", "adsr_2", "lfo_1", "lfo_2"]
MOD_OUTPUTS = ["vco_1_pitch", "vco_1_amp", "vco_2_pitch", "vco_2_amp", "noise_amp"]
_MODMATRIX = [(f"{i}->{o}", 0.0, 1.0, 0.5, False) for i in MOD_INPUTS for o in MOD_OUTPUTS]
_MIXER = [("vco_1", 0.0, 1.0, 1.0, False), ("vco_2", 0.0, 1.0, 1.0, False), ("noise", 0.0, 1.0, 0.1, False)]

# Voice.__init__ registration order (modules without parameters omitted).
MODULES: "OrderedDict[str, list]" = OrderedDict(
    [
        ("keyboard", _KEYBOARD),
        ("adsr_1", _ADSR),
        ("adsr_2", _ADSR),
        ("lfo_1", _LFO),
        ("lfo_2", _LFO),
        ("lfo_1_amp_adsr", _ADSR),
        ("lfo_2_amp_adsr", _ADSR),
        ("lfo_1_rate_adsr", _ADSR),
        ("lfo_2_rate_adsr", _ADSR),
        ("mod_matrix", _MODMATRIX),
        ("vco_1", _VCO),
        ("vco_2", _SQUARESAW),
        ("mixer", _MIXER),
    ]
)
NPARAMS = sum(len(v) for v in MODULES.values())
assert NPARAMS == 78  # conf/config.yaml:27


def registration_keys() -> List[Tuple[str, str]]:
    """Order of ``nn.Module.parameters()`` = order of the ``params`` tensor ``Voice.forward`` returns.  DEFINES."""
    return [(m, p[0]) for m, plist in MODULES.items() for p in plist]


def sorted_keys() -> List[Tuple[str, str]]:
    """Order of ``sorted(named_parameters())`` = the order ``randomize(seed)`` assigns random rows (App. A.2)."""
    named = sorted((f"{m}.torchparameters.{p[0]}", (m, p[0])) for m, plist in MODULES.items() for p in plist)
    return [k for _, k in named]


def get_parameters_keys() -> List[Tuple[str, str]]:
    """Order of ``Voice.get_parameters()``: sorted module name, then declaration order (audio_to_params.py:240-246)."""
    return [(m, p[0]) for m in sorted(MODULES) for p in MODULES[m]]


def spec(module: str, name: str):
    for p in MODULES[module]:
        if p[0] == name:
            return p
    raise KeyError((module, name))


# --------------------------------------------------------------------------------------------------
# Config (a1)
# --------------------------------------------------------------------------------------------------
class SynthConfigO:
    def __init__(self, batch_size=128, sample_rate=44100, buffer_size_seconds=4.0, control_rate=441,
                 reproducible=True, no_grad=True, eps=1e-6):
        if reproducible:
            assert batch_size % 32 == 0
        self.batch_size = batch_size
        self.sample_rate = sample_rate
        self.buffer_size_seconds = buffer_size_seconds
        self.buffer_size = int(buffer_size_seconds * sample_rate)
        self.control_rate = control_rate
        self.control_buffer_size = int(buffer_size_seconds * control_rate)
        self.reproducible = reproducible
        self.no_grad = no_grad
        self.eps = eps


# --------------------------------------------------------------------------------------------------
# Seeding (a2, a3, a16)
# --------------------------------------------------------------------------------------------------
def seeded_params(batch_idx: int, batch_size: int) -> torch.Tensor:
    """[B,78] in *sorted* order: row i = torch.rand(78) from a CPU generator seeded with batch_idx*B+i (App. A.2)."""
    g = torch.Generator(device="cpu")
    rows = []
    for i in range(batch_size):
        g.manual_seed(batch_idx * batch_size + i)
        rows.append(torch.rand((NPARAMS,), generator=g))
    return torch.stack(rows, dim=0)


def mt19937_uniform24(seed: int, n: int) -> np.ndarray:
    """Plain restatement of what ``torch.rand(n)`` yields on a CPU generator after ``manual_seed(seed)``:
    MT19937 seeded with the low 32 bits (Knuth LCG 1812433253), tempered 32-bit outputs, low 24 bits * 2^-24."""
    mt = [0] * 624
    mt[0] = seed & 0xFFFFFFFF
    for j in range(1, 624):
        mt[j] = (1812433253 * (mt[j - 1] ^ (mt[j - 1] >> 30)) + j) & 0xFFFFFFFF
    out = np.empty(n, dtype=np.float32)
    assert n <= 227  # first twist only needs mt[k], mt[k+1], mt[k+397]
    for k in range(n):
        y = (mt[k] & 0x80000000) | (mt[k + 1] & 0x7FFFFFFF)
        v = mt[k + 397] ^ (y >> 1) ^ (0x9908B0DF if (y & 1) else 0)
        v ^= v >> 11
        v ^= (v << 7) & 0x9D2C5680
        v ^= (v << 15) & 0xEFC60000
        v ^= v >> 18
        out[k] = np.float32((v & 0xFFFFFF) * (1.0 / 16777216.0))
    return out


def is_train(batch_idx: int, batch_size: int) -> torch.Tensor:
    idx = torch.arange(batch_idx * batch_size, (batch_idx + 1) * batch_size)
    return (idx // 32) % 10 != 9


def noise_table(rows: int, T: int, seed: int = 13) -> torch.Tensor:
    """Noise module buffer: U(-1,1) from a CPU generator seeded 13, [rows,T] row-major (App. A.9)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    n = torch.empty((rows, T))
    n.uniform_(-1.0, 1.0, generator=g)
    return n


# --------------------------------------------------------------------------------------------------
# Parameter scaling (a4)
# --------------------------------------------------------------------------------------------------
def from_0to1(u: torch.Tensor, lo: float, hi: float, curve: float, symmetric: bool) -> torch.Tensor:
    if not symmetric:
        if curve != 1.0:
            u = torch.exp2(torch.log2(u) / curve)
        return lo + (hi - lo) * u
    dist = 2.0 * u - 1.0
    # DEFINES: symmetric curve through exp2/log2 like the non-symmetric branch
    shaped = torch.sign(dist) * torch.exp2(torch.log2(torch.abs(dist)) / curve)
    return lo + (hi - lo) / 2.0 * (shaped + 1.0)


class _P:
    """Parameter lookup ``p(module, name)`` over a [B,78] sorted-order 0..1 tensor."""

    def __init__(self, params01_sorted: torch.Tensor):
        self.u = params01_sorted
        self.index = {k: i for i, k in enumerate(sorted_keys())}

    def __call__(self, module: str, name: str) -> torch.Tensor:
        _, lo, hi, curve, sym = spec(module, name)
        return from_0to1(self.u[:, self.index[(module, name)]], lo, hi, curve, sym)


# --------------------------------------------------------------------------------------------------
# Modules (a5 - a15)
# --------------------------------------------------------------------------------------------------
def _adsr(p: _P, mod: str, note_on: torch.Tensor, C: int, cr: int, eps: float) -> torch.Tensor:
    attack, decay, sustain = p(mod, "attack"), p(mod, "decay"), p(mod, "sustain")
    release, alpha = p(mod, "release"), p(mod, "alpha")
    rng = torch.arange(C)

    def ramp(duration, start=None, inverse=False):
        duration = (duration * cr).unsqueeze(1)
        start_ = (start * cr).unsqueeze(1) if start is not None else 0.0
        r = rng.expand(duration.shape[0], C) - start_
        r = torch.maximum(r, torch.tensor(0))
        r = (r + eps) / duration + eps
        r = torch.minimum(r, torch.tensor(1.0, dtype=r.dtype))
        if inverse:
            r = torch.where(duration > 0.0, 1.0 - r, r)
        return torch.pow(r, alpha.unsqueeze(1))

    new_attack = torch.minimum(attack, note_on)
    new_decay = torch.maximum(note_on - attack, torch.tensor(0, dtype=note_on.dtype))
    new_decay = torch.minimum(new_decay, decay)
    a_sig = ramp(new_attack)
    s = sustain.unsqueeze(1)
    d_sig = (1.0 - s) * ramp(new_decay, start=new_attack, inverse=True) + s
    r_sig = ramp(release, start=note_on, inverse=True)
    return a_sig * d_sig * r_sig


def _lfo(p: _P, mod: str, mod_signal: torch.Tensor, cr: int) -> torch.Tensor:
    frequency = p(mod, "frequency").unsqueeze(1)
    modulation = p(mod, "mod_depth").unsqueeze(1) * mod_signal
    f = torch.maximum(frequency + modulation, torch.tensor(0.0, dtype=mod_signal.dtype))
    arg = torch.cumsum(2 * PI * f / cr, dim=1)
    arg = arg + p(mod, "initial_phase").unsqueeze(1)
    cos = torch.cos(arg + PI)
    square = torch.sign(cos)
    cos = (cos + 1.0) / 2.0
    square = (square + 1.0) / 2.0
    saw = torch.remainder(arg, 2 * PI) / (2 * PI)
    rev_saw = 1.0 - saw
    tri = 2 * saw
    tri = torch.where(tri > 1.0, 2.0 - tri, tri)
    shapes = torch.stack([cos, tri, saw, rev_saw, square], dim=1)  # [B,5,C]
    mode = torch.stack([p(mod, t) for t in _LFO_TYPES], dim=1)  # [B,5]
    mode = torch.pow(mode, torch.tensor(2.718281828, dtype=mode.dtype))
    mode = mode / torch.sum(mode, dim=1, keepdim=True)
    return torch.matmul(mode.unsqueeze(1), shapes).squeeze(1)


def _mod_matrix(p: _P, signals: List[torch.Tensor]) -> torch.Tensor:
    w = torch.stack([p("mod_matrix", n[0]) for n in _MODMATRIX], dim=1)  # [B,20] input-major
    w = w.view(-1, 4, 5)
    w = torch.swapaxes(w, 1, 2)  # [B,5,4]
    w = w / torch.sum(w, dim=2, keepdim=True)
    return torch.matmul(w, torch.stack(signals, dim=1))  # [B,5,C]


def _upsample(x: torch.Tensor, T: int) -> torch.Tensor:
    return torch.nn.functional.interpolate(x.unsqueeze(1), size=T, mode="linear", align_corners=True).squeeze(1)


def _midi_to_hz(m: torch.Tensor) -> torch.Tensor:
    return 440.0 * torch.exp2((m - 69.0) / 12.0)


def _vco_argument(p: _P, mod: str, midi_f0: torch.Tensor, mod_signal: torch.Tensor, sr: int) -> torch.Tensor:
    m = (midi_f0 + p(mod, "tuning")).unsqueeze(1)
    modulation = p(mod, "mod_depth").unsqueeze(1) * mod_signal
    control = torch.clamp(m + modulation, 0.0, 127.0)
    hz = _midi_to_hz(control)
    arg = torch.cumsum(2 * PI * hz / sr, dim=1)
    arg += p(mod, "initial_phase").unsqueeze(1)
    return arg


def voice_render(
    params01_sorted: torch.Tensor,
    noise: torch.Tensor,
    T: int = 176400,
    C: int = 1764,
    sample_rate: int = 44100,
    control_rate: int = 441,
    eps: float = 1e-6,
    dtype: torch.dtype = torch.float32,
    intermediates: bool = False,
) -> Dict[str, torch.Tensor]:
    """Voice.output() (App. A.4-A.9) on CPU.  ``noise`` is [R,T]; row b uses noise[b % R]."""
    with torch.no_grad():
        u = params01_sorted.to(dtype)
        B = u.shape[0]
        p = _P(u)
        midi_f0, note_on = p("keyboard", "midi_f0"), p("keyboard", "duration")
        lfo_1_rate = _adsr(p, "lfo_1_rate_adsr", note_on, C, control_rate, eps)
        lfo_2_rate = _adsr(p, "lfo_2_rate_adsr", note_on, C, control_rate, eps)
        lfo_1_amp = _adsr(p, "lfo_1_amp_adsr", note_on, C, control_rate, eps)
        lfo_2_amp = _adsr(p, "lfo_2_amp_adsr", note_on, C, control_rate, eps)
        lfo_1 = _lfo(p, "lfo_1", lfo_1_rate, control_rate) * lfo_1_amp
        lfo_2 = _lfo(p, "lfo_2", lfo_2_rate, control_rate) * lfo_2_amp
        adsr_1 = _adsr(p, "adsr_1", note_on, C, control_rate, eps)
        adsr_2 = _adsr(p, "adsr_2", note_on, C, control_rate, eps)
        ctrl = _mod_matrix(p, [adsr_1, adsr_2, lfo_1, lfo_2])  # [B,5,C]
        up = [_upsample(ctrl[:, i, :], T) for i in range(5)]

        arg1 = _vco_argument(p, "vco_1", midi_f0, up[0], sample_rate)
        vco_1 = torch.cos(arg1) * up[1]

        arg2 = _vco_argument(p, "vco_2", midi_f0, up[2], sample_rate)
        max_pitch = midi_f0 + torch.maximum(p("vco_2", "mod_depth"), torch.tensor(0, dtype=dtype))
        max_f0 = _midi_to_hz(max_pitch)
        partials = (12000 / (max_f0 * torch.log10(max_f0))).unsqueeze(1)
        square = torch.tanh(PI * partials * torch.sin(arg2) / 2)
        shape = p("vco_2", "shape").unsqueeze(1)
        vco_2 = ((1 - shape / 2) * square * (1 + shape * torch.cos(arg2))) * up[3]

        R = noise.shape[0]
        nz = noise.to(dtype)
        if R != B:
            nz = nz.repeat((B + R - 1) // R, 1)[:B]
        noise_out = nz * up[4]

        levels = torch.stack([p("mixer", "vco_1"), p("mixer", "vco_2"), p("mixer", "noise")], dim=1)  # [B,3]
        mixed = torch.matmul(levels.unsqueeze(1), torch.stack([vco_1, vco_2, noise_out], dim=1)).squeeze(1)
        peak = torch.max(torch.abs(mixed), dim=1, keepdim=True)[0]
        audio = torch.where(peak > 1.0, mixed / peak, mixed)
        out = {"audio": audio, "peak": peak.squeeze(1)}
        if intermediates:
            out.update(
                ctrl=ctrl, arg1=arg1, arg2=arg2, mixed=mixed,
                adsr=torch.stack([adsr_1, adsr_2, lfo_1_amp, lfo_2_amp, lfo_1_rate, lfo_2_rate], dim=1),
                lfo=torch.stack([lfo_1, lfo_2], dim=1),
            )
        return out


def sorted_to_registration(params01_sorted: torch.Tensor) -> torch.Tensor:
    idx = {k: i for i, k in enumerate(sorted_keys())}
    cols = [idx[k] for k in registration_keys()]
    return params01_sorted[:, cols]


def voice_forward(batch_idx: int, cfg: SynthConfigO, noise: Optional[torch.Tensor] = None, dtype=torch.float32):
    """``Voice(synthconfig)(batch_idx)`` -> (audio[B,T], params[B,78] registration order, is_train[B])."""
    B = cfg.batch_size
    if noise is None:
        noise = noise_table(32 if cfg.reproducible else B, cfg.buffer_size)
    u = seeded_params(batch_idx, B)
    out = voice_render(u, noise, cfg.buffer_size, cfg.control_buffer_size, cfg.sample_rate, cfg.control_rate,
                       cfg.eps, dtype)
    return out["audio"], sorted_to_registration(u), is_train(batch_idx, B)
