"""ORACLE (test infrastructure, not product code): numpy restatement of the reference's ``VICReg.loss``.

Follows ``/root/reference/vicreg.py``:
  * ``VICReg.loss`` ......... vicreg.py:35-58   (invariance / variance / covariance terms)
  * ``off_diagonal`` ........ vicreg.py:73-76   (row-major list of the n(n-1) off-diagonal entries)
  * ``FullGatherLayer`` ..... vicreg.py:79-95   (dead code in the reference: ``dist`` is never imported and the call at
                              vicreg.py:38-39 is commented out).  Its contract, restated here: forward = rank-ordered
                              concatenation of every rank's [B_local,D]; backward = sum over ranks of the incoming
                              gradients, sliced to the own rank.

Quirks reproduced on purpose: ``repr_loss`` uses the *local* rows (it precedes the gather, vicreg.py:36);
the covariance divisor is ``cfg.vicreg.batch_size - 1`` (vicreg.py:47-48), not the row count; the variance is the
unbiased one over the actual row count; the covariance term is divided by ``cfg.embeddim`` (vicreg.py:49).

Pinned by ``tests/golden/vicreg_*.npz`` produced by ``oracle/make_golden.py`` from the reference module itself.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np


def off_diagonal(m: np.ndarray) -> np.ndarray:
    n, k = m.shape
    assert n == k
    return m.flatten()[:-1].reshape(n - 1, n + 1)[:, 1:].flatten()


def loss(
    x: np.ndarray,
    y: np.ndarray,
    cfg_batch_size: int,
    embeddim: int,
    sim_coeff: float = 25.0,
    std_coeff: float = 25.0,
    cov_coeff: float = 1.0,
    dtype=np.float64,
    local_rows: slice | None = None,
) -> Tuple[float, float, float, float]:
    """(loss, repr_loss, std_loss, cov_loss).  ``x``/``y`` are the (gathered) [B,D]; ``local_rows`` selects the rows
    of this rank for the invariance term (default: all rows, i.e. the single-process reference)."""
    x = x.astype(dtype)
    y = y.astype(dtype)
    xl, yl = (x, y) if local_rows is None else (x[local_rows], y[local_rows])
    repr_loss = np.mean((xl - yl) ** 2, dtype=dtype)
    x = x - x.mean(axis=0, dtype=dtype)
    y = y - y.mean(axis=0, dtype=dtype)
    std_x = np.sqrt(x.var(axis=0, ddof=1, dtype=dtype) + dtype(0.0001))
    std_y = np.sqrt(y.var(axis=0, ddof=1, dtype=dtype) + dtype(0.0001))
    std_loss = np.mean(np.maximum(1 - std_x, 0)) / 2 + np.mean(np.maximum(1 - std_y, 0)) / 2
    cov_x = (x.T @ x) / (cfg_batch_size - 1)
    cov_y = (y.T @ y) / (cfg_batch_size - 1)
    cov_loss = (off_diagonal(cov_x) ** 2).sum() / embeddim + (off_diagonal(cov_y) ** 2).sum() / embeddim
    total = sim_coeff * repr_loss + std_coeff * std_loss + cov_coeff * cov_loss
    return float(total), float(repr_loss), float(std_loss), float(cov_loss)


def loss_torch(x, y, cfg_batch_size: int, embeddim: int, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0):
    """The same terms with torch CPU fp32 ops in the reference's order (vicreg.py:35-58); CPU-baseline timing path."""
    import torch
    import torch.nn.functional as F

    repr_loss = F.mse_loss(x, y)
    x = x - x.mean(dim=0)
    y = y - y.mean(dim=0)
    std_x = torch.sqrt(x.var(dim=0) + 0.0001)
    std_y = torch.sqrt(y.var(dim=0) + 0.0001)
    std_loss = torch.mean(F.relu(1 - std_x)) / 2 + torch.mean(F.relu(1 - std_y)) / 2
    cov_x = (x.T @ x) / (cfg_batch_size - 1)
    cov_y = (y.T @ y) / (cfg_batch_size - 1)
    n = cov_x.shape[0]
    offx = cov_x.flatten()[:-1].view(n - 1, n + 1)[:, 1:]
    offy = cov_y.flatten()[:-1].view(n - 1, n + 1)[:, 1:]
    cov_loss = offx.pow(2).sum() / embeddim + offy.pow(2).sum() / embeddim
    total = sim_coeff * repr_loss + std_coeff * std_loss + cov_coeff * cov_loss
    return float(total), float(repr_loss), float(std_loss), float(cov_loss)


def loss_grad(x, y, cfg_batch_size, embeddim, sim_coeff=25.0, std_coeff=25.0, cov_coeff=1.0):
    """Analytic d loss / d x, d loss / d y in float64 (single process, all rows local)."""
    x = x.astype(np.float64)
    y = y.astype(np.float64)
    B, D = x.shape
    g_repr = 2.0 * (x - y) / (B * D)

    def side(a):
        ac = a - a.mean(axis=0)
        var = (ac ** 2).sum(axis=0) / (B - 1)
        std = np.sqrt(var + 1e-4)
        # d/da of mean(relu(1-std))/2 ; centring is a projection, and sum_r ac = 0 so the mean term vanishes
        gs = np.where(std < 1.0, -1.0, 0.0) / (2.0 * D) * (1.0 / (2.0 * std)) * (2.0 / (B - 1))
        g_std = ac * gs[None, :]
        cov = (ac.T @ ac) / (cfg_batch_size - 1)
        off = cov - np.diag(np.diag(cov))
        g_cov = (4.0 / (embeddim * (cfg_batch_size - 1))) * (ac @ off)
        g_cov = g_cov - g_cov.mean(axis=0)
        return std_coeff * g_std, cov_coeff * g_cov

    sx, cx = side(x)
    sy, cy = side(y)
    return sim_coeff * g_repr + sx + cx, -sim_coeff * g_repr + sy + cy


def full_gather_forward(shards: Sequence[np.ndarray]) -> np.ndarray:
    return np.concatenate(list(shards), axis=0)


def full_gather_backward(grads_per_rank: Sequence[np.ndarray], rank: int, b_local: int) -> np.ndarray:
    """Every rank holds a gradient w.r.t. the gathered [W*B_local, D]; all-reduce(sum), keep the own slice."""
    total = np.sum(np.stack(list(grads_per_rank)), axis=0)
    return total[rank * b_local:(rank + 1) * b_local]


def pooled_statistics(shards: Sequence[np.ndarray]):
    """Global column mean, centred second moments and centred Gram of the rank-ordered concatenation, computed from
    per-rank summaries only -- the algebra of the statistics exchange (``ias_vicreg_loss_stats``), restated in float64:

        mu = sum_q B_q mu_q / B,      G = sum_q [ G_q + B_q (mu_q - mu)(mu_q - mu)^T ],      m2 = diag(G)

    with mu_q, G_q the mean and the locally-centred Gram of rank q's rows.  Equal, up to rounding, to centring the
    gathered batch as vicreg.py:40-48 does after the (intended) FullGatherLayer call at vicreg.py:38-39."""
    shards = [np.asarray(s, dtype=np.float64) for s in shards]
    B = sum(s.shape[0] for s in shards)
    mus = [s.mean(axis=0) for s in shards]
    mu = sum(s.shape[0] * m for s, m in zip(shards, mus)) / B
    G = np.zeros((shards[0].shape[1],) * 2)
    for s, m in zip(shards, mus):
        c = s - m
        d = m - mu
        G += c.T @ c + s.shape[0] * np.outer(d, d)
    return mu, np.diag(G).copy(), G
