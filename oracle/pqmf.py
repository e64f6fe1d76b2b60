"""ORACLE (test infrastructure, not product code): numpy restatement of the reference's ``pqmf.PQMF``.

Follows ``/root/reference/pqmf.py``:
  * filter design ........ pqmf.py:18-33  (scipy ``firwin(taps+1, cutoff, kaiser beta)``, cosine modulation with
                           the reference's ``(taps-1)/2`` centre -- its own TODO at pqmf.py:26 is *not* fixed)
  * ``analysis`` ......... pqmf.py:49-50  (strided cross-correlation, zero padding taps//2)
  * ``synthesis`` ........ pqmf.py:52-55  (zero-stuff by N with gain N, then N->1 FIR, zero padding taps//2)

Pinned by ``tests/golden/pqmf_*.npz``, produced by ``oracle/make_golden.py`` from the reference module itself.
Accumulation is float64; the reference runs float32 (oneDNN), so comparisons use the north-star tolerance
``max|a-b| / max|b| <= 1e-5``.
"""
from __future__ import annotations

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view
from scipy import signal as sig


def design(N: int = 4, taps: int = 62, cutoff: float = 0.15, beta: float = 9.0):
    """-> (H[N,taps+1], G[N,taps+1]) float32, i.e. reference buffers ``H[:,0,:]`` and ``G[0,:,:]``."""
    proto = sig.firwin(taps + 1, cutoff, window=("kaiser", beta))
    j = np.arange(taps + 1)
    H = np.zeros((N, taps + 1))
    G = np.zeros((N, taps + 1))
    for k in range(N):
        theta = (2 * k + 1) * (np.pi / (2 * N)) * (j - ((taps - 1) / 2))
        phase = (-1) ** k * np.pi / 4
        H[k] = 2 * proto * np.cos(theta + phase)
        G[k] = 2 * proto * np.cos(theta - phase)
    return H.astype(np.float32), G.astype(np.float32)


def out_len(T: int, N: int, taps: int = 62) -> int:
    return (T + 2 * (taps // 2) - (taps + 1)) // N + 1


def analysis(x: np.ndarray, H: np.ndarray, N: int) -> np.ndarray:
    """x[B,T] float32, H[N,K] -> [B,N,L];  out[b,k,n] = sum_j H[k,j] * x[b, n*N + j - K//2]."""
    B, T = x.shape
    K = H.shape[1]
    pad = K // 2  # == taps // 2 for odd K = taps + 1
    xp = np.zeros((B, T + 2 * pad), dtype=np.float64)
    xp[:, pad:pad + T] = x
    win = sliding_window_view(xp, K, axis=1)[:, ::N, :]  # [B,L,K]
    return np.einsum("blj,kj->bkl", win, H.astype(np.float64)).astype(np.float32)


def synthesis(z: np.ndarray, G: np.ndarray, N: int) -> np.ndarray:
    """z[B,N,L] float32, G[N,K] -> [B, (L-1)*N + N];  conv_transpose1d(stride N, kernel N) then FIR."""
    B, n, L = z.shape
    assert n == N
    K = G.shape[1]
    pad = K // 2
    Tout = L * N
    up = np.zeros((B, N, Tout + 2 * pad), dtype=np.float64)
    up[:, :, pad:pad + Tout:N] = (z * np.float32(N)).astype(np.float32)  # fp32 product like the reference
    win = sliding_window_view(up, K, axis=2)  # [B,N,Tout,K]
    return np.einsum("bktj,kj->bt", win, G.astype(np.float64)).astype(np.float32)


def analysis_conv1d(x, H, N: int):
    """Same contraction through torch's CPU conv1d (the library call the reference makes, pqmf.py:49-50); used for the
    CPU-baseline timing so the baseline is not handicapped by the float64 numpy restatement above."""
    import torch
    import torch.nn.functional as F

    xt = torch.as_tensor(x).unsqueeze(1)
    Ht = torch.as_tensor(H).unsqueeze(1)
    return F.conv1d(xt, Ht, padding=(H.shape[1] - 1) // 2, stride=N)


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """The tolerance metric of SURVEY 8(c): max|a-b| / max|b|."""
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))) / max(np.max(np.abs(b)), 1e-30))
