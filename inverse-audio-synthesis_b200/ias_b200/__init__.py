"""ias_b200 -- B200-native front end of inverse-audio-synthesis (Voice render -> PQMF -> VICReg loss).

Host-side mirrors of the three reference surfaces over the C ABI in include/ias_b200.h:

    from ias_b200 import SynthConfig, Voice, PQMF, VICReg, FullGatherLayer, off_diagonal

The package contains no arithmetic of its own and no fallback path: every call goes to libias_b200.so
(hand-written sm_100a kernels) and raises if the library or a CUDA device is missing.
"""
from ._lib import IasError, build, lib  # noqa: F401
from .dist import Communicator, EmbeddingExchange, StatsExchange, use_communicator, use_fused_gather  # noqa: F401
from .pqmf import PQMF  # noqa: F401
from .vicreg import FullGatherLayer, Projector, VICReg, exclude_bias_and_norm, off_diagonal, vicreg_loss  # noqa: F401
from .voice import ModuleParameter, ModuleParameterRange, SynthConfig, Voice  # noqa: F401

__version__ = "0.1.0"
