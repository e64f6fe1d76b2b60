"""CPU: the index algebra behind the cosine-modulated PQMF kernels and the pooled analysis epilogue, restated in numpy /
pure Python and checked against the oracle (which is pinned by the reference's goldens) and against torch.

The CUDA kernels in csrc/pqmf.cu implement exactly these formulas (UnfoldCM, SynthGeom, SynRows, the pooled epilogue
and k_pool_finalize, magic_for / magic_div); the -m gpu tests check the kernels themselves."""
import numpy as np
import pytest
import torch
from scipy import signal as sig

from oracle import pqmf as OP

K, TAPS, PAD = 63, 62, 31


def _proto_signed(N, cutoff=0.15):
    proto = sig.firwin(TAPS + 1, cutoff, window=("kaiser", 9.0))
    j = np.arange(K)
    return 2 * proto * np.where((j // (2 * N)) % 2 == 0, 1.0, -1.0)


def _phase(j, N):
    return ((PAD - j) % N + N) % N


def _offset(j, N):
    return (_phase(j, N) + j - PAD) // N


@pytest.mark.parametrize("N", [8, 16])
def test_synthesis_dct4_unfold(N):
    """k_pqmf_synthesis_cm: v[r] = (-1)^q (N/sqrt2) (U[m'] +- U[N-1-m']) with U the DCT-IV of the band values."""
    _, G = OP.design(N)
    g = _proto_signed(N)
    k = np.arange(N)[:, None]
    m = np.arange(N)[None, :]
    C = np.cos((2 * k + 1) * (2 * m + 1) * np.pi / (4 * N)) * N / np.sqrt(2)
    rng = np.random.default_rng(0)
    L = 41
    z = rng.uniform(-1, 1, (2, N, L)).astype(np.float32)
    U = np.einsum("km,bkl->bml", C, z.astype(np.float64))
    v = np.zeros((2, 2 * N, L))
    for r in range(2 * N):
        t2 = 2 * r - (K - 2)
        w = ((t2 + 2 * N) % (4 * N) + 4 * N) % (4 * N) - 2 * N
        mp, pos, neg = (abs(w) - 1) // 2, w > 0, ((w - t2) // (4 * N)) & 1
        val = U[:, mp] + (1 if pos else -1) * U[:, N - 1 - mp]
        v[:, r] = -val if neg else val
    y = np.zeros((2, L * N))
    for j in range(K):
        p, o = _phase(j, N), _offset(j, N)
        for n in range(L):
            if 0 <= n + o < L:
                y[:, n * N + p] += g[j] * v[:, j % (2 * N), n + o]
    ref = OP.synthesis(z, G, N)
    assert OP.rel_err(y.astype(np.float32), ref) <= 1e-6


@pytest.mark.parametrize("N,Q", [(2, 4), (3, 8), (3, 4), (4, 4)])
def test_synthesis_small_n_row_halves(N, Q):
    """k_pqmf_synthesis_small: the N taps that meet one (time step, row) pair are consecutive and their residues
    alternate between two fixed halves of the modulated row."""
    _, G = OP.design(N)
    g = _proto_signed(N)
    offs = [_offset(j, N) for j in range(K)]
    dmin, halo = min(offs), max(offs) - min(offs)
    jb = lambda o: PAD + N * (o - 1) + 1  # noqa: E731
    ra = jb(dmin) % (2 * N)
    res = lambda h, e: (ra + h * N + e) % (2 * N)  # noqa: E731
    c = np.array([[N * np.cos((2 * k + 1) * (np.pi / (2 * N)) * (r - (K - 2) / 2) - (1 if k % 2 == 0 else -1) * np.pi / 4)
                   for r in range(2 * N)] for k in range(N)])
    rng = np.random.default_rng(1)
    L = 29
    z = rng.uniform(-1, 1, (N, L)).astype(np.float32)
    ref = OP.synthesis(z[None], G, N)[0]

    def row(m):
        out = np.zeros((2, N))
        if 0 <= m < L:
            for h in range(2):
                for e in range(N):
                    out[h, e] = sum(c[k, res(h, e)] * float(z[k, m]) for k in range(N))
        return out

    y = np.zeros(L * N)
    for n0 in range(0, L, Q):
        acc = np.zeros((Q, N))
        for i in range(Q + halo):
            w = row(n0 + dmin + i)
            for q in range(Q):
                d = i - q
                if 0 <= d <= halo:
                    for e in range(N):
                        j = jb(dmin + d) + e
                        if 0 <= j < K:
                            assert _offset(j, N) == dmin + d and j % (2 * N) == res(d & 1, e)
                            acc[q, _phase(j, N)] += g[j] * w[d & 1, e]
        for q in range(Q):
            if n0 + q < L:
                y[(n0 + q) * N:(n0 + q + 1) * N] = acc[q]
    assert OP.rel_err(y.astype(np.float32), ref) <= 1e-6


@pytest.mark.parametrize("N", [8, 16])
def test_analysis_dct4_fold(N):
    """k_pqmf_analysis, N >= 8: out[k] = sum_m C[k][m]/sqrt2 * ((P[m]+Q[m]) - (P[N-1-m]-Q[N-1-m])) on the 2N polyphase
    partial sums, with C symmetric (the kernel reads it by rows)."""
    H, _ = OP.design(N)
    g = _proto_signed(N)
    k = np.arange(N)[:, None]
    m = np.arange(N)[None, :]
    C = np.cos((2 * k + 1) * (2 * m + 1) * np.pi / (4 * N)) / np.sqrt(2)
    assert np.allclose(C, C.T)
    rng = np.random.default_rng(2)
    T = 700
    x = rng.uniform(-1, 1, (1, T)).astype(np.float32)
    ref = OP.analysis(x, H, N)[0]
    xp = np.zeros(T + 2 * PAD)
    xp[PAD:PAD + T] = x[0]
    wrap = lambda v: ((v % (2 * N)) + 2 * N) % (2 * N)  # noqa: E731
    out = np.zeros_like(ref, dtype=np.float64)
    for n in range(ref.shape[1]):
        ps = np.zeros(2 * N)
        for j in range(K):
            ps[j % (2 * N)] += g[j] * xp[n * N + j]
        sm, df = np.zeros(N), np.zeros(N)
        for mm in range(N):
            a, b = mm + (K - 1) // 2, (K - 3) // 2 - mm
            pm = ps[wrap(a)] * (-1 if ((a - wrap(a)) // (2 * N)) & 1 else 1)
            qm = ps[wrap(b)] * (-1 if ((b - wrap(b)) // (2 * N)) & 1 else 1)
            sm[mm], df[mm] = pm + qm, pm - qm
        for mm in range(N):
            out[:, n] += C[mm, :] * (sm[mm] - df[N - 1 - mm])
    assert OP.rel_err(out.astype(np.float32), ref) <= 1e-6


def _magic(d):
    sh = 1
    while (1 << sh) < d:
        sh += 1
    return ((1 << (31 + sh)) + d - 1) // d, sh - 1


def test_magic_division_is_exact_for_31_bit_numerators():
    rng = np.random.default_rng(3)
    for d in [2, 3, 7, 255, 256, 257, 689, 4096, 58800, 176400, 1323000, 2 ** 27 - 1, 2 ** 30, 2 ** 31 - 1]:
        m, sh = _magic(d)
        assert m < 2 ** 32 and sh < 32
        xs = [0, 1, d - 1, d, d + 1, 2 ** 31 - 1, (2 ** 31 - 1) // d * d, (2 ** 31 - 1) // d * d - 1]
        xs += [int(v) for v in rng.integers(0, 2 ** 31, 20000)]
        for x in xs:
            if 0 <= x < 2 ** 31:
                assert ((x * m) >> 32) >> sh == x // d


@pytest.mark.parametrize("N,L,P,tile_n", [(3, 58800, 256, 512), (16, 11025, 256, 256), (3, 33335, 64, 512), (8, 5513, 16, 512)])
def test_pooled_epilogue_bin_algebra(N, L, P, tile_n):
    """A CTA tile of tile_n consecutive flattened elements touches at most two adaptive-pooling bins (ilo, ilo+1) when
    tile_n + 1 <= floor(S / P); per-tile slot sums recombined the way k_pool_finalize does equal torch's pooling."""
    S = N * L
    assert tile_n + 1 <= S // P and S * P < 2 ** 31
    rng = np.random.default_rng(4)
    a = np.abs(rng.standard_normal(S))
    tiles = (L + tile_n - 1) // tile_n
    partial = np.zeros((N, tiles, 2))
    for k in range(N):
        for t in range(tiles):
            first = k * L + t * tile_n
            ilo = first * P // S
            s1 = (ilo + 1) * S // P
            e0 = ((ilo + 1) * S + P - 1) // P
            for n in range(t * tile_n, min((t + 1) * tile_n, L)):
                f = k * L + n
                lo_bin, hi_bin = f * P // S, ((f + 1) * P - 1) // S  # the bins element f belongs to
                assert ilo <= lo_bin and hi_bin <= ilo + 1
                if f < e0:
                    partial[k, t, 0] += a[f]
                if f >= s1:
                    partial[k, t, 1] += a[f]
    feat = np.zeros(P)
    for i in range(P):
        s, e = i * S // P, ((i + 1) * S + P - 1) // P
        for k in range(s // L, (e - 1) // L + 1):
            na, nb = max(s, k * L) - k * L, min(e, (k + 1) * L) - k * L
            for t in range(na // tile_n, (nb - 1) // tile_n + 1):
                slot = i - (k * L + t * tile_n) * P // S
                if slot in (0, 1):
                    feat[i] += partial[k, t, slot]
        feat[i] /= e - s
    want = torch.nn.functional.adaptive_avg_pool1d(torch.from_numpy(a).reshape(1, 1, -1), P).reshape(-1).numpy()
    assert np.allclose(feat, want, rtol=1e-12, atol=0)
