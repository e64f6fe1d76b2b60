"""CPU: the algebra and the buffer layout of the statistics exchange (ias_vicreg_loss_stats), no GPU needed.
The N-rank == 1-rank property of the real kernels is checked on the GPU by tests/test_gpu_multi.py and, at every N, by
bench.py's `parity` block."""
import numpy as np

from oracle import make_golden as MG
from oracle import vicreg as OV


def test_pooled_formulas_equal_statistics_of_the_concatenation():
    for world, b_local, D in [(2, 16, 8), (8, 32, 24), (3, 7, 5)]:
        x_all, _ = MG.vicreg_inputs(world * b_local, D, "correlated", seed=world)
        x = x_all.double().numpy() + np.repeat(np.arange(world), b_local)[:, None] * 0.7  # rank-dependent means
        shards = [x[q * b_local:(q + 1) * b_local] for q in range(world)]
        mu, m2, G = OV.pooled_statistics(shards)
        xc = x - x.mean(axis=0)
        assert np.allclose(mu, x.mean(axis=0), rtol=0, atol=1e-12)
        assert np.allclose(G, xc.T @ xc, rtol=1e-11, atol=1e-10)
        assert np.allclose(m2, (xc ** 2).sum(axis=0), rtol=1e-11)
    # unequal shards (the kernels require equal ones; the algebra does not)
    shards = [x[:5], x[5:11], x[11:]]
    mu, m2, G = OV.pooled_statistics(shards)
    assert np.allclose(G, xc.T @ xc, rtol=1e-11, atol=1e-10)


def test_exchange_buffer_layout(built_lib):
    """ias_vicreg_stats_buffer_bytes: 1 KiB header + world x 2 packets of 4*Dp + 2*ntiles*128*128 floats (256-float
    granularity), Dp = D rounded up to the 128-wide Gram tile."""
    import ias_b200

    lib = ias_b200.lib()
    for world, D in [(1, 256), (2, 256), (8, 256), (8, 200), (4, 512)]:
        Dp = (D + 127) // 128 * 128
        dt = Dp // 128
        ntiles = dt * (dt + 1) // 2
        packet = (4 * Dp + 2 * ntiles * 128 * 128 + 255) // 256 * 256
        assert lib.ias_vicreg_stats_buffer_bytes(world, D) == 4 * (256 + world * 2 * packet)
    assert lib.ias_vicreg_stats_buffer_bytes(0, 256) == 0


def test_gather_refuses_a_per_gpu_covariance_divisor(monkeypatch):
    """With the gather on, cfg.vicreg.batch_size is the divisor of the GLOBAL batch (vicreg.py:47-48): a per-GPU value
    must not pass silently."""
    import pytest
    import torch

    import ias_b200
    from ias_b200 import vicreg as V

    monkeypatch.setattr(V, "_world", lambda: (0, 4))
    x = torch.zeros(8, 4)
    with pytest.raises(ias_b200.IasError, match="GLOBAL batch"):
        V.vicreg_loss(x, x, 8, 4, 25.0, 25.0, 1.0, gather=True)
