"""ORACLE tooling: freeze the Voice oracle's OWN output, so that ``oracle/voice.py`` cannot drift silently.

torchsynth is absent from the reference tree and from this image: nothing reference-held pins the synth
stage ("parity unpinned").  The next best pin is a fixture of what the restatement produces today:

  * seeded parameters and the noise table -- integer / generator arithmetic, host independent -> sha256, exact;
  * the six ADSR envelopes (control rate) -- torch's SLEEF ``pow`` (FMA build), identical on every AVX2/AVX-512 host
    we have seen; compared to 2e-7 so a non-FMA host still passes while an algorithmic change cannot;
  * the five modulation-matrix outputs -- through MKL's ``cos`` (host dependent in the last ulp), compared to 5e-7;
  * the audio of 32 one-second voices, subsampled -- compared per voice (the pitch path amplifies ulps: SURVEY H1), the
    median over voices must stay <= 1e-5.

    python oracle/make_voice_fixture.py      # rewrites tests/golden/voice_oracle.npz
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import voice as V  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "voice_oracle.npz")
B, T, C = 32, 44100, 441
SUB_C, SUB_T = 7, 101


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def compute():
    u = V.seeded_params(3, B)
    noise = V.noise_table(32, T)
    o = V.voice_render(u, noise, T, C, intermediates=True)
    return dict(
        params_sha=np.array([sha(V.seeded_params(0, 128)), sha(V.seeded_params(123456, 32)), sha(u)]),
        is_train=V.is_train(9, 64).numpy(),
        noise_sha=np.array([sha(noise), sha(V.noise_table(32, 176400)[:, ::1009])]),
        adsr=o["adsr"][:, :, ::SUB_C].numpy(),
        ctrl=o["ctrl"][:, :, ::SUB_C].numpy(),
        audio=o["audio"][:, ::SUB_T].numpy(),
        peak=o["peak"].numpy(),
    )


if __name__ == "__main__":
    np.savez_compressed(OUT, **compute())
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
