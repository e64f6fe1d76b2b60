import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "inverse-audio-synthesis_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
SHIM_SRC = os.path.join(ROOT, "tests", "shim", "voice_host.cpp")
SHIM_SO = os.path.join(ROOT, "tests", "shim", "voice_host.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """libias_b200.so, built on demand (nvcc cross-compiles without a GPU)."""
    import ias_b200

    if not os.path.exists(ias_b200._lib.LIB_PATH):
        ias_b200.build()
    return ias_b200._lib.LIB_PATH


@pytest.fixture(scope="session")
def voice_shim():
    """Host build of csrc/voice_math.cuh (test-only; see tests/shim/voice_host.cpp)."""
    import ctypes

    hdr = os.path.join(PKG, "csrc", "voice_math.cuh")
    stale = (not os.path.exists(SHIM_SO)) or os.path.getmtime(SHIM_SO) < max(os.path.getmtime(SHIM_SRC),
                                                                             os.path.getmtime(hdr))
    if stale:
        subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", SHIM_SRC, "-o",
                        SHIM_SO], check=True)
    return ctypes.CDLL(SHIM_SO)


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
