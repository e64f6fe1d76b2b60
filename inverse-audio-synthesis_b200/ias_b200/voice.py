"""Drop-in for ``torchsynth.config.SynthConfig`` and ``torchsynth.synth.Voice`` as the reference uses them.

Call sites mirrored (file:line into /root/reference):
  * ``vicreg_audio_params.py:86-94``  ``SynthConfig(batch_size=, reproducible=, sample_rate=, buffer_size_seconds=)``,
    ``Voice(synthconfig=...)``, ``.to(device)``
  * ``vicreg_audio_params.py:114`` / ``pretrain.py:75`` / ``audio_to_params.py:215``  ``voice(batch_idx) ->
    (audio[B,T], params[B,78], is_train[B])``
  * ``audio_to_params.py:238-257``  ``get_parameters()``, ``getattr(voice, module).set_parameter_0to1(name, value)``,
    ``freeze_parameters(keys)``, ``voice(None)``, ``unfreeze_all_parameters()``

This module owns no synthesis arithmetic: parameters live in one ``[78, B]`` device tensor (registration order, the
layout of ``include/ias_b200.h``) that the per-name ``ModuleParameter`` objects view, and ``randomize`` / ``output``
are calls into ``libias_b200.so``.  The behavioural spec of the synth is SURVEY.md Appendix A.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Tuple

import torch
from torch import nn

from . import _lib

BASE_REPRODUCIBLE_BATCH_SIZE = 32
PI = math.pi


class SynthConfig:
    """torchsynth ``SynthConfig``: plain numbers, same attribute names."""

    def __init__(
        self,
        batch_size: int = 128,
        sample_rate: int = 44100,
        buffer_size_seconds: float = 4.0,
        control_rate: int = 441,
        reproducible: bool = True,
        no_grad: bool = True,
        debug: bool = False,
        eps: float = 1e-6,
    ):
        if reproducible and batch_size % BASE_REPRODUCIBLE_BATCH_SIZE != 0:
            raise ValueError(
                f"Reproducibility currently only supported with batch sizes that are multiples of "
                f"{BASE_REPRODUCIBLE_BATCH_SIZE}"
            )
        self.batch_size = int(batch_size)
        self.sample_rate = int(sample_rate)
        self.buffer_size_seconds = float(buffer_size_seconds)
        self.buffer_size = int(self.buffer_size_seconds * self.sample_rate)
        self.control_rate = int(control_rate)
        self.control_buffer_size = int(self.buffer_size_seconds * self.control_rate)
        self.reproducible = bool(reproducible)
        self.no_grad = bool(no_grad)
        self.debug = bool(debug)
        self.eps = float(eps)

    def to(self, device):  # torchsynth moves its tensor attributes; nothing to move here
        return self


class ModuleParameterRange:
    """torchsynth ``ModuleParameterRange``: 0..1 <-> human range with a curve.  Host-side utility; the
    device kernels carry their own copy of the forward mapping."""

    def __init__(self, minimum: float, maximum: float, curve: float = 1.0, symmetric: bool = False, name: str = "",
                 description: str = ""):
        self.minimum, self.maximum, self.curve, self.symmetric = float(minimum), float(maximum), float(curve), symmetric
        self.name, self.description = name, description

    def from_0to1(self, normalized: torch.Tensor) -> torch.Tensor:
        if not self.symmetric:
            if self.curve != 1.0:
                normalized = torch.exp2(torch.log2(normalized) / self.curve)
            return self.minimum + (self.maximum - self.minimum) * normalized
        dist = 2.0 * normalized - 1.0
        shaped = torch.sign(dist) * torch.exp2(torch.log2(torch.abs(dist)) / self.curve)
        return self.minimum + (self.maximum - self.minimum) / 2.0 * (shaped + 1.0)

    def to_0to1(self, value: torch.Tensor) -> torch.Tensor:
        normalized = (value - self.minimum) / (self.maximum - self.minimum)
        if not self.symmetric:
            if self.curve != 1.0:
                normalized = torch.pow(normalized, self.curve)
            return normalized
        dist = 2.0 * normalized - 1.0
        return (1.0 + torch.sign(dist) * torch.pow(torch.abs(dist), self.curve)) / 2.0

    def __repr__(self):
        return (f"ModuleParameterRange(name={self.name}, min={self.minimum}, max={self.maximum}, curve={self.curve}, "
                f"symmetric={self.symmetric})")


class ModuleParameter(nn.Parameter):
    """A [batch] tensor of 0..1 values with its range, viewing one row of the Voice's parameter block."""

    def __new__(cls, data: torch.Tensor, parameter_name: str, parameter_range: ModuleParameterRange):
        obj = super().__new__(cls, data, requires_grad=False)
        obj.parameter_name = parameter_name
        obj.parameter_range = parameter_range
        obj.frozen = False
        return obj

    def __deepcopy__(self, memo):
        out = ModuleParameter(self.data.clone(), self.parameter_name, self.parameter_range)
        out.frozen = self.frozen
        memo[id(self)] = out
        return out

    def __reduce_ex__(self, proto):
        return (_rebuild_module_parameter, (self.data, self.parameter_name, self.parameter_range, self.frozen))

    def from_0to1(self) -> torch.Tensor:
        return self.parameter_range.from_0to1(self.data)

    @staticmethod
    def is_parameter_frozen(parameter: "ModuleParameter") -> bool:
        return bool(getattr(parameter, "frozen", False))


def _rebuild_module_parameter(data, name, prange, frozen):
    p = ModuleParameter(data, name, prange)
    p.frozen = frozen
    return p


def _adsr_ranges():
    return [
        ModuleParameterRange(0.0, 2.0, 0.5, name="attack", description="attack time (sec)"),
        ModuleParameterRange(0.0, 2.0, 0.5, name="decay", description="decay time (sec)"),
        ModuleParameterRange(0.0, 1.0, name="sustain", description="sustain amplitude 0-1"),
        ModuleParameterRange(0.0, 5.0, 0.5, name="release", description="release time (sec)"),
        ModuleParameterRange(0.1, 6.0, name="alpha", description="envelope factor: 1 linear, >1 exponential"),
    ]


def _lfo_ranges():
    return [
        ModuleParameterRange(0.0, 20.0, 0.25, name="frequency", description="Frequency in Hz of oscillation"),
        ModuleParameterRange(-10.0, 20.0, 0.5, True, name="mod_depth", description="LFO rate modulation in Hz"),
        ModuleParameterRange(-PI, PI, name="initial_phase", description="Initial phase of LFO"),
    ] + [ModuleParameterRange(0.0, 1.0, name=n, description=f"Selection parameter for {n} LFO")
         for n in ("sin", "tri", "saw", "rsaw", "sqr")]


def _vco_ranges(shape: bool):
    r = [
        ModuleParameterRange(-24.0, 24.0, name="tuning", description="tuning adjustment in midi"),
        ModuleParameterRange(-96.0, 96.0, 0.2, True, name="mod_depth", description="depth of pitch modulation"),
        ModuleParameterRange(-PI, PI, name="initial_phase", description="initial phase"),
    ]
    if shape:
        r.append(ModuleParameterRange(0.0, 1.0, name="shape", description="square (0) to saw (1)"))
    return r


MOD_INPUTS = ("adsr_1", "adsr_2", "lfo_1", "lfo_2")
MOD_OUTPUTS = ("vco_1_pitch", "vco_1_amp", "vco_2_pitch", "vco_2_amp", "noise_amp")

# (module name, torchsynth class name, parameter ranges) in Voice.__init__ registration order
_VOICE_LAYOUT = [
    ("keyboard", "MonophonicKeyboard", lambda: [
        ModuleParameterRange(0.0, 127.0, name="midi_f0", description="pitch value in 'midi' (69 = 440Hz)"),
        ModuleParameterRange(0.01, 4.0, 0.5, name="duration", description="note-on button, in seconds")]),
    ("adsr_1", "ADSR", _adsr_ranges),
    ("adsr_2", "ADSR", _adsr_ranges),
    ("lfo_1", "LFO", _lfo_ranges),
    ("lfo_2", "LFO", _lfo_ranges),
    ("lfo_1_amp_adsr", "ADSR", _adsr_ranges),
    ("lfo_2_amp_adsr", "ADSR", _adsr_ranges),
    ("lfo_1_rate_adsr", "ADSR", _adsr_ranges),
    ("lfo_2_rate_adsr", "ADSR", _adsr_ranges),
    ("control_upsample", "ControlRateUpsample", lambda: []),
    ("mod_matrix", "ModulationMixer", lambda: [
        ModuleParameterRange(0.0, 1.0, 0.5, name=f"{i}->{o}", description=f"Modulation {i} to {o}")
        for i in MOD_INPUTS for o in MOD_OUTPUTS]),
    ("vco_1", "SineVCO", lambda: _vco_ranges(False)),
    ("vco_2", "SquareSawVCO", lambda: _vco_ranges(True)),
    ("noise", "Noise", lambda: []),
    ("vca", "VCA", lambda: []),
    ("mixer", "AudioMixer", lambda: [
        ModuleParameterRange(0.0, 1.0, 1.0, name="vco_1", description="vco_1 mix level"),
        ModuleParameterRange(0.0, 1.0, 1.0, name="vco_2", description="vco_2 mix level"),
        ModuleParameterRange(0.0, 1.0, 0.1, name="noise", description="noise mix level")]),
]


class SynthModule(nn.Module):
    """The parameter-holding shell of one torchsynth module (no arithmetic: the kernels render the whole voice)."""

    def __init__(self, kind: str, synthconfig: SynthConfig, ranges: List[ModuleParameterRange], rows: torch.Tensor):
        super().__init__()
        self.kind = kind
        self.synthconfig = synthconfig
        self.batch_size = synthconfig.batch_size
        self.parameter_ranges = ranges
        self.torchparameters = nn.ParameterDict()
        for i, r in enumerate(ranges):
            self.torchparameters[r.name] = ModuleParameter(rows[i], r.name, r)
        if kind == "ADSR":  # torchsynth keeps this persistent buffer; kept for state-dict key parity
            self.register_buffer("range", torch.arange(synthconfig.control_buffer_size))

    def get_parameter(self, parameter_id: str) -> ModuleParameter:
        return self.torchparameters[parameter_id]

    def get_parameter_0to1(self, parameter_id: str) -> torch.Tensor:
        return self.torchparameters[parameter_id].data

    def set_parameter_0to1(self, parameter_id: str, value: torch.Tensor) -> None:
        """audio_to_params.py:243-246.  Values are copied into the shared block, cast/moved as needed."""
        p = self.torchparameters[parameter_id]
        value = torch.as_tensor(value)
        if value.shape != p.data.shape:
            raise ValueError(f"{parameter_id}: expected shape {tuple(p.data.shape)}, got {tuple(value.shape)}")
        p.data.copy_(value.detach().to(device=p.device, dtype=p.dtype))

    def set_parameter(self, parameter_id: str, value: torch.Tensor) -> None:
        p = self.torchparameters[parameter_id]
        self.set_parameter_0to1(parameter_id, p.parameter_range.to_0to1(torch.as_tensor(value, dtype=p.dtype)))

    def p(self, parameter_id: str) -> torch.Tensor:
        return self.torchparameters[parameter_id].from_0to1()


class Noise(SynthModule):
    """torchsynth ``Noise``: the pre-computed white-noise table."""

    def __init__(self, synthconfig: SynthConfig, seed: int):
        super().__init__("Noise", synthconfig, [], torch.empty(0))
        rows = BASE_REPRODUCIBLE_BATCH_SIZE if synthconfig.reproducible else synthconfig.batch_size
        generator = torch.Generator(device="cpu").manual_seed(seed)
        noise = torch.empty((rows, synthconfig.buffer_size), device="cpu")
        noise.uniform_(-1.0, 1.0, generator=generator)
        self.register_buffer("noise", noise)


class Voice(nn.Module):
    """torchsynth ``Voice`` rendered by ``libias_b200.so`` (``ias_voice_seed_params`` / ``ias_voice_render``)."""

    def __init__(self, synthconfig: Optional[SynthConfig] = None, nebula: str = "default", normalize=True):
        """``normalize``: True = torchsynth's normalize_if_clipping on the rendered audio (default), False = raw mix,
        ``"defer"`` = raw mix plus ``self.row_scale`` [B] (1/peak of a clipping row, else 1) for a linear consumer to
        apply -- ``PQMF.analysis(audio, row_scale=voice.row_scale)`` gives the bands of the normalised audio without
        the second pass over the clipping rows (not a torchsynth option)."""
        super().__init__()
        if normalize not in (True, False, "defer"):
            raise ValueError("normalize must be True, False or 'defer'")
        if nebula != "default":
            raise ValueError("only the default nebula (torchsynth class defaults) is implemented")
        self.synthconfig = synthconfig if synthconfig is not None else SynthConfig()
        cfg = self.synthconfig
        self.normalize = normalize
        lib = _lib.lib()
        store = torch.rand((_lib.NPARAMS, cfg.batch_size))
        self.register_buffer("_store", store, persistent=False)  # [78, B] registration order
        self.register_buffer("_is_train", torch.zeros(cfg.batch_size, dtype=torch.uint8), persistent=False)
        row = 0
        self._rows: "OrderedDict[Tuple[str, str], int]" = OrderedDict()
        for name, kind, ranges_fn in _VOICE_LAYOUT:
            ranges = ranges_fn()
            if kind == "Noise":
                module: SynthModule = Noise(cfg, seed=13)
            else:
                module = SynthModule(kind, cfg, ranges, self._store[row:row + len(ranges)])
            for r in ranges:
                expect = lib.ias_voice_param_name(row).decode()
                if expect != f"{name}/{r.name}":
                    raise _lib.IasError(f"parameter table mismatch at row {row}: {expect} vs {name}/{r.name}")
                self._rows[(name, r.name)] = row
                row += 1
            self.add_module(name, module)
        assert row == _lib.NPARAMS
        self._workspace: Optional[torch.Tensor] = None
        self._peak: Optional[torch.Tensor] = None

    # ---- torchsynth AbstractSynth surface ------------------------------------------------------------------
    @property
    def batch_size(self) -> int:
        return self.synthconfig.batch_size

    @property
    def sample_rate(self) -> int:
        return self.synthconfig.sample_rate

    @property
    def buffer_size(self) -> int:
        return self.synthconfig.buffer_size

    @property
    def device(self) -> torch.device:
        return self._store.device

    def _named_synth_modules(self):
        for name, module in self.named_children():
            if isinstance(module, SynthModule):
                yield name, module

    def get_parameters(self, include_frozen: bool = False) -> "OrderedDict[Tuple[str, str], ModuleParameter]":
        """Sorted module name, then declaration order (audio_to_params.py:240-246 zips this with predicted params)."""
        out = []
        for module_name, module in sorted(self._named_synth_modules()):
            for parameter in module.torchparameters.values():
                if include_frozen or not ModuleParameter.is_parameter_frozen(parameter):
                    out.append(((module_name, parameter.parameter_name), parameter))
        return OrderedDict(out)

    def set_parameters(self, params: Dict[Tuple[str, str], torch.Tensor], freeze: bool = False) -> None:
        for (module_name, param_name), value in params.items():
            getattr(self, module_name).set_parameter(param_name, value)
            if freeze:
                getattr(self, module_name).get_parameter(param_name).frozen = True

    def freeze_parameters(self, params: Iterable[Tuple[str, str]]) -> None:
        for module_name, param_name in params:
            getattr(self, module_name).get_parameter(param_name).frozen = True

    def unfreeze_all_parameters(self) -> None:
        for _, module in self._named_synth_modules():
            for parameter in module.torchparameters.values():
                parameter.frozen = False

    def _batch_idx_to_is_train(self, batch_idx: int) -> torch.Tensor:
        idxs = torch.arange(self.batch_size * batch_idx, self.batch_size * (batch_idx + 1), device=self.device)
        return (idxs // BASE_REPRODUCIBLE_BATCH_SIZE) % 10 != 9

    # ---- storage plumbing ----------------------------------------------------------------------------------
    def _param_list(self) -> List[ModuleParameter]:
        plist = self.__dict__.get("_plist")
        if plist is None:
            plist = [getattr(self, m).torchparameters[n] for (m, n) in self._rows]  # registration order == row order
            self.__dict__["_plist"] = plist
        return plist

    def _tie(self) -> None:
        """Make every ModuleParameter a view of its row of ``_store`` again (after .to(), load_state_dict, or a user
        assigning ``parameter.data``); values the user put in the parameter win.  The common case (nothing moved) is
        78 pointer comparisons."""
        store = self._store
        base, stride = store.data_ptr(), store.stride(0) * store.element_size()
        for row, p in enumerate(self._param_list()):
            if p.data_ptr() != base + row * stride:
                view = store[row]
                view.copy_(p.data.to(device=store.device, dtype=store.dtype))
                p.data = view

    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        self._tie()
        self._workspace = None
        self._peak = None
        return self

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._tie()

    def _ws(self) -> torch.Tensor:
        cfg = self.synthconfig
        need = _lib.lib().ias_voice_workspace_bytes(cfg.batch_size, cfg.buffer_size, cfg.control_buffer_size)
        if self._workspace is None or self._workspace.device != self.device or self._workspace.numel() < need:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace

    # ---- randomize / output / forward ----------------------------------------------------------------------
    def randomize(self, seed: Optional[int] = None) -> None:
        """AbstractSynth.randomize: sound i gets torch-CPU-MT19937(seed*B+i) uniforms, on the device."""
        self._tie()
        if seed is None:
            if self.synthconfig.reproducible:
                raise ValueError("Reproducible mode is on, you must pass a seed")
            frozen = self._frozen_rows()
            fresh = torch.rand_like(self._store)
            keep = torch.tensor(frozen, dtype=torch.bool, device=self.device).unsqueeze(1)
            self._store.copy_(torch.where(keep, self._store, fresh))
            return
        _lib.require_cuda(self._store, "Voice parameters")
        frozen = (_lib.c_uint8 * _lib.NPARAMS)(*self._frozen_rows())
        if isinstance(seed, torch.Tensor) and seed.is_cuda:
            # the batch number stays on the device: no host read-back, and the call can be captured in a CUDA graph
            if seed.dtype != torch.int64 or seed.numel() != 1:
                raise ValueError("a device-resident batch index must be a single int64")
            with _lib.on_device(self._store):
                rc = _lib.lib().ias_voice_seed_params_dev(
                    _lib.ptr(seed), self.batch_size, frozen, _lib.ptr(self._store), _lib.ptr(self._is_train),
                    _lib.current_stream(self.device))
            _lib.check(rc, "ias_voice_seed_params_dev")
            return
        with _lib.on_device(self._store):
            rc = _lib.lib().ias_voice_seed_params(
                int(seed) * self.batch_size, self.batch_size, frozen, _lib.ptr(self._store), _lib.ptr(self._is_train),
                _lib.current_stream(self.device))
        _lib.check(rc, "ias_voice_seed_params")

    def _normalize_mode(self) -> int:
        return 2 if self.normalize == "defer" else (1 if self.normalize else 0)

    @property
    def row_scale(self) -> Optional[torch.Tensor]:
        """[B] factors of the last render when ``normalize == "defer"`` (else None)."""
        return self._peak if self.normalize == "defer" else None

    def _frozen_rows(self) -> List[int]:
        return [1 if p.frozen else 0 for p in self._param_list()]

    STAGE_CONTROL, STAGE_AUDIO, STAGE_ENVELOPES, STAGE_MODULATION = 1, 2, 4, 8

    def prepare(self, batch_idx) -> None:
        """Seed the parameters of batch ``batch_idx`` and run the control stage (envelopes, LFOs, modulation, work
        queue) into the workspace, without rendering.  ``forward(batch_idx, prepared=True)`` then renders exactly that
        batch with the audio stage alone.  A training loop knows its next batch number, so it can issue this on a side
        stream while the current batch's audio is still being consumed (PQMF, backbone, loss): the control stage is
        latency-bound and hides behind that work.  Not a torchsynth method."""
        if batch_idx is None:
            raise ValueError("Voice.prepare needs a batch index")
        on_device = isinstance(batch_idx, torch.Tensor) and batch_idx.is_cuda
        with torch.no_grad():
            self.randomize(seed=batch_idx if on_device else int(batch_idx))
            self._render(self.STAGE_CONTROL, None)
        self._prepared = batch_idx if on_device else int(batch_idx)
        self._snap_valid = False

    def prepare_envelopes(self, batch_idx) -> None:
        """First half of ``prepare`` for callers that pipeline two batches deep: seed the parameters of ``batch_idx``
        and run its six ADSR envelopes (compute bound) into the workspace.  ``prepare_modulation()`` completes the
        control stage later.  Between the two calls the parameter store holds this batch; a batch whose modulation
        stage already ran keeps its parameters in a snapshot, so it may still be waiting for its audio stage."""
        if batch_idx is None:
            raise ValueError("Voice.prepare_envelopes needs a batch index")
        on_device = isinstance(batch_idx, torch.Tensor) and batch_idx.is_cuda
        with torch.no_grad():
            self.randomize(seed=batch_idx if on_device else int(batch_idx))
            self._render(self.STAGE_ENVELOPES, None)
        self._enveloped = batch_idx if on_device else int(batch_idx)

    def prepare_modulation(self) -> None:
        """Second half of ``prepare``: LFOs, modulation matrix, per-interval records and the work queue of the batch
        ``prepare_envelopes`` seeded; snapshots its parameters / train flags (the store is reseeded before this batch
        is rendered).  ``forward(batch_idx, prepared=True)`` then renders it."""
        want = getattr(self, "_enveloped", None)
        if want is None:
            raise RuntimeError("Voice.prepare_modulation: no batch has its envelopes prepared")
        with torch.no_grad():
            self._render(self.STAGE_MODULATION, None)
            if getattr(self, "_params_snap", None) is None or self._params_snap.device != self.device:
                self._params_snap = torch.empty((self.batch_size, _lib.NPARAMS), dtype=torch.float32, device=self.device)
                self._is_train_snap = torch.empty_like(self._is_train)
            self._params_snap.copy_(self._store.t())
            self._is_train_snap.copy_(self._is_train)
        self._prepared, self._enveloped, self._snap_valid = want, None, True

    def _render(self, stages: int, audio: Optional[torch.Tensor], phase_debug=None, ctrl_in=None):
        cfg = self.synthconfig
        _lib.require_cuda(self._store, "Voice parameters")
        noise = self.noise.noise
        _lib.require_cuda(noise, "Voice noise buffer")
        if self._peak is None or self._peak.device != self.device:
            self._peak = torch.empty(cfg.batch_size, dtype=torch.float32, device=self.device)
        ws = self._ws()
        with _lib.on_device(self._store):
            rc = _lib.lib().ias_voice_render_stages(
                _lib.ptr(self._store), _lib.ptr(noise), noise.shape[0], _lib.ptr(audio), _lib.ptr(self._peak),
                cfg.batch_size, cfg.buffer_size, cfg.control_buffer_size, float(cfg.sample_rate),
                float(cfg.control_rate), float(cfg.eps), self._normalize_mode(), _lib.ptr(ctrl_in),
                _lib.ptr(phase_debug), _lib.ptr(ws), ws.numel(), stages, _lib.current_stream(self.device))
        _lib.check(rc, "ias_voice_render_stages")

    def output(self, return_peak: bool = False, phase_debug: Optional[torch.Tensor] = None,
               ctrl_in: Optional[torch.Tensor] = None, _tied: bool = False) -> torch.Tensor:
        """Voice.output(): render the current parameters to audio [B, T].  ``phase_debug`` / ``ctrl_in`` are the
        inspection hooks of ``ias_voice_render`` (parity tests only)."""
        if not _tied:
            self._tie()
        cfg = self.synthconfig
        _lib.require_cuda(self._store, "Voice parameters")
        noise = self.noise.noise
        _lib.require_cuda(noise, "Voice noise buffer")
        audio = torch.empty((cfg.batch_size, cfg.buffer_size), dtype=torch.float32, device=self.device)
        if self._peak is None or self._peak.device != self.device:
            self._peak = torch.empty(cfg.batch_size, dtype=torch.float32, device=self.device)
        ws = self._ws()
        with _lib.on_device(self._store):
            rc = _lib.lib().ias_voice_render(
                _lib.ptr(self._store), _lib.ptr(noise), noise.shape[0], _lib.ptr(audio), _lib.ptr(self._peak),
                cfg.batch_size, cfg.buffer_size, cfg.control_buffer_size, float(cfg.sample_rate),
                float(cfg.control_rate), float(cfg.eps), self._normalize_mode(), _lib.ptr(ctrl_in),
                _lib.ptr(phase_debug), _lib.ptr(ws), ws.numel(), _lib.current_stream(self.device))
        _lib.check(rc, "ias_voice_render")
        return (audio, self._peak) if return_peak else audio

    def control_signals(self) -> torch.Tensor:
        """[B, 5, C] modulation-matrix outputs (vco_1_pitch, vco_1_amp, vco_2_pitch, vco_2_amp, noise_amp)."""
        self._tie()
        cfg = self.synthconfig
        _lib.require_cuda(self._store, "Voice parameters")
        ctrl = torch.empty((cfg.batch_size, _lib.NCONTROL, cfg.control_buffer_size), dtype=torch.float32,
                           device=self.device)
        ws = self._ws()
        with _lib.on_device(self._store):
            rc = _lib.lib().ias_voice_control(
                _lib.ptr(self._store), cfg.batch_size, cfg.control_buffer_size, float(cfg.control_rate), float(cfg.eps),
                _lib.ptr(ctrl), _lib.ptr(ws), ws.numel(), _lib.current_stream(self.device))
        _lib.check(rc, "ias_voice_control")
        return ctrl

    def params01(self) -> torch.Tensor:
        """[B, 78] current 0..1 parameters in registration order (the `params` tensor of forward)."""
        self._tie()
        return self._store.t().contiguous()

    def forward(self, batch_idx=None, prepared: bool = False):
        """-> (audio[B,T], params[B,78], is_train[B] or None)  (vicreg_audio_params.py:114).  ``batch_idx`` is an int
        (or anything ``int()`` accepts, as in the reference) or a one-element int64 CUDA tensor, which is read on the
        device.  ``prepared=True``: ``prepare(batch_idx)`` already seeded this batch and ran its control stage; only the
        audio stage runs (same result, bit for bit)."""
        if self.synthconfig.reproducible and batch_idx is None:
            raise ValueError("Reproducible mode is on, you must pass a batch index")
        ctx = torch.no_grad() if self.synthconfig.no_grad else torch.enable_grad()
        if prepared:
            want = getattr(self, "_prepared", None)
            on_device = isinstance(batch_idx, torch.Tensor) and batch_idx.is_cuda
            if want is None or (not on_device and not isinstance(want, torch.Tensor) and int(batch_idx) != want):
                raise RuntimeError(f"Voice.forward(prepared=True): batch {batch_idx} was not prepared (have {want})")
            self._prepared = None
            with ctx:
                if getattr(self, "_snap_valid", False):  # prepared in two halves: the store may hold the batch after
                    is_train = self._is_train_snap.bool()
                    params = self._params_snap.clone()
                else:
                    is_train = self._is_train.bool()
                    params = self._store.t().contiguous()
                cfg = self.synthconfig
                audio = torch.empty((cfg.batch_size, cfg.buffer_size), dtype=torch.float32, device=self.device)
                self._render(self.STAGE_AUDIO, audio)
            return audio, params, is_train
        with ctx:
            if batch_idx is not None:
                on_device = isinstance(batch_idx, torch.Tensor) and batch_idx.is_cuda
                self.randomize(seed=batch_idx if on_device else int(batch_idx))  # ties the parameter views
                is_train = self._is_train.bool()
            else:
                self._tie()
                is_train = None
            params = self._store.t().contiguous()
            audio = self.output(_tied=True)
        return audio, params, is_train
