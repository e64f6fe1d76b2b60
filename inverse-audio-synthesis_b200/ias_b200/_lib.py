"""ctypes binding of libias_b200.so (include/ias_b200.h).

The shared library is built in-tree by ``csrc/Makefile`` (``__graft_entry__.build()``) and loaded from this
directory.  There is no fallback of any kind: if the library is missing, or a call returns an error, the caller
gets an exception.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IAS_B200_LIB") or os.path.join(_HERE, "libias_b200.so")  # override: tuning builds only
COMM_LIB_PATH = os.path.join(_HERE, "libias_comm.so")
CSRC = os.path.join(os.path.dirname(_HERE), "csrc")

IAS_OK = 0
IAS_ERR_INVALID, IAS_ERR_CUDA, IAS_ERR_UNSUPPORTED, IAS_ERR_WORKSPACE, IAS_ERR_NCCL = 1, 2, 3, 4, 5
NPARAMS = 78
NCONTROL = 5

_lib = None
_comm = None


class IasError(RuntimeError):
    """Raised for every failed library call; ``code`` is the IAS_ERR_* value the C ABI returned (None when the error
    was detected on the Python side)."""

    def __init__(self, message: str, code=None):
        super().__init__(message)
        self.code = code


def build(verbose: bool = False) -> None:
    """Compile every CUDA source for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
        print(proc.stderr)
    if proc.returncode != 0:
        raise IasError(f"building libias_b200.so failed (exit {proc.returncode})")


_SIGNATURES = {
    "ias_version": (c_int, []),
    "ias_last_error": (c_char_p, []),
    "ias_device_check": (c_int, [c_int]),
    "ias_prof_enable": (c_int, [c_int]),
    "ias_prof_reset": (c_int, []),
    "ias_prof_kernel_count": (c_int, []),
    "ias_prof_kernel_name": (c_char_p, [c_int]),
    "ias_prof_launches": (ctypes.c_longlong, [c_int]),
    "ias_prof_read": (c_int, [c_int, POINTER(ctypes.c_double), POINTER(ctypes.c_longlong)]),
    "ias_voice_param_name": (c_char_p, [c_int]),
    "ias_voice_sorted_index": (c_int, [c_int]),
    "ias_voice_seed_params": (c_int, [c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ias_voice_seed_params_dev": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ias_voice_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ias_voice_control": (c_int, [c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ias_voice_render": (
        c_int,
        [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_float, c_int,
         c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "ias_voice_render_stages": (
        c_int,
        [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_float, c_int,
         c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p],
    ),
    "ias_pqmf_out_len": (c_int, [c_int, c_int, c_int]),
    "ias_pqmf_analysis": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                  c_int, c_int, c_void_p]),
    "ias_pqmf_analysis_image": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ias_pqmf_synthesis": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                   c_void_p]),
    "ias_vicreg_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ias_vicreg_loss": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p,
         c_size_t, c_void_p],
    ),
    "ias_vicreg_loss_backward": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p,
         c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "ias_vicreg_gather_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ias_vicreg_loss_gather": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p,
         c_size_t, c_void_p],
    ),
    "ias_vicreg_loss_gather_backward": (
        c_int,
        [c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
         c_size_t, c_void_p],
    ),
    "ias_vicreg_stats_buffer_bytes": (c_size_t, [c_int, c_int]),
    "ias_vicreg_loss_stats": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p,
         c_void_p, c_size_t, c_void_p],
    ),
    "ias_vicreg_loss_stats_stages": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p,
         c_void_p, c_size_t, c_int, c_void_p],
    ),
    "ias_vicreg_loss_stats_backward": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p,
         c_void_p, c_size_t, c_void_p],
    ),
    "ias_vicreg_gram_reference": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ias_vicreg_gram_tc": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ias_pqmf_pool_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "ias_pqmf_analysis_pooled": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_int, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_void_p]),
    "ias_abs_avg_pool": (c_int, [c_void_p, c_void_p, c_int, ctypes.c_longlong, c_int, c_void_p]),
}

_COMM_SIGNATURES = {
    "ias_comm_unique_id": (c_int, [c_void_p]),
    "ias_comm_init": (c_int, [c_void_p, c_int, c_int, POINTER(c_void_p)]),
    "ias_comm_destroy": (c_int, [c_void_p]),
    "ias_comm_allgather": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ias_comm_reduce_scatter": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ias_comm_last_error": (c_char_p, []),
}

# every symbol include/ias_b200.h declares, by library (tests check both files export exactly these)
CORE_SYMBOLS = tuple(_SIGNATURES)
COMM_SYMBOLS = tuple(k for k in _COMM_SIGNATURES if k != "ias_comm_last_error")


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IasError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the ias_b200 hot path)"
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def comm_lib() -> ctypes.CDLL:
    global _comm
    if _comm is None:
        if not os.path.exists(COMM_LIB_PATH):
            raise IasError(f"{COMM_LIB_PATH} is missing: run __graft_entry__.build()")
        import torch  # noqa: F401  (makes sure torch's bundled libnccl.so.2 is already mapped)

        handle = ctypes.CDLL(COMM_LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in _COMM_SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _comm = handle
    return _comm


def check(rc: int, what: str = "") -> None:
    if rc != IAS_OK:
        msg = lib().ias_last_error().decode("utf-8", "replace")
        raise IasError(f"{what or 'libias_b200'} failed (code {rc}): {msg}", code=rc)


def check_comm(rc: int, what: str = "") -> None:
    if rc != IAS_OK:
        msg = comm_lib().ias_comm_last_error().decode("utf-8", "replace")
        raise IasError(f"{what or 'libias_comm'} failed (code {rc}): {msg}", code=rc)


def ptr(t) -> c_void_p:
    """Device (or host) address of a torch tensor; None -> NULL."""
    return c_void_p(0 if t is None else t.data_ptr())


def current_stream(device) -> c_void_p:
    import torch

    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def on_device(t):
    """Context manager making ``t``'s GPU the current CUDA device for the duration of a library call: kernels launch on
    the calling thread's current device, which need not be the tensor's when a process drives several GPUs."""
    import torch

    return torch.cuda.device(t.device)


def require_cuda(t, name: str) -> None:
    if not t.is_cuda:
        raise IasError(
            f"{name} is on {t.device}: the ias_b200 hot path runs on a B200 only (no CPU fallback; the CPU "
            "restatement lives in oracle/ and is test infrastructure)"
        )


__all__ = [
    "IasError", "build", "lib", "comm_lib", "check", "check_comm", "ptr", "current_stream", "require_cuda", "on_device",
    "LIB_PATH", "COMM_LIB_PATH", "NPARAMS", "NCONTROL", "CORE_SYMBOLS", "COMM_SYMBOLS", "c_uint8",
]
