"""CPU: the pairing algebra and the shared-memory layouts of the packed synthesis kernels of round 2, session 3
(k_pqmf_synthesis_n3p, k_pqmf_synthesis_n4p, k_pqmf_synthesis_cm2 in csrc/pqmf.cu), restated in numpy and checked against
the oracle (pinned by the reference's goldens), plus a simulation of the bank mapping each layout was designed for.

The kernels implement exactly these formulas; the -m gpu tests check the kernels themselves (bit-identical to the first
formulations).  Bank model (measured with ncu on the B200): a 128-bit shared access is served per quarter
warp (8 lanes over eight 16-byte bank groups), a 64-bit access per half warp (16 lanes over sixteen 8-byte bank pairs)."""
import numpy as np
import pytest
from scipy import signal as sig

from oracle import pqmf as OP

K, TAPS, PAD = 63, 62, 31
THREADS = 128


def _proto_signed(N, cutoff=0.15):
    proto = sig.firwin(TAPS + 1, cutoff, window=("kaiser", 9.0))
    j = np.arange(K)
    return 2 * proto * np.where((j // (2 * N)) % 2 == 0, 1.0, -1.0)


def _phase(j, N):
    return ((PAD - j) % N + N) % N


def _offset(j, N):
    return (_phase(j, N) + j - PAD) // N


def _geometry(N):
    offs = [_offset(j, N) for j in range(K)]
    return min(offs), max(offs) - min(offs)


def _mod_small(N):
    """c[k][r] of TapsSynSmall: N cos(theta_k (r - (K-2)/2) - (-1)^k pi/4)."""
    return np.array([[N * np.cos((2 * k + 1) * (np.pi / (2 * N)) * (r - (K - 2) / 2) - (1 if k % 2 == 0 else -1) * np.pi / 4)
                      for r in range(2 * N)] for k in range(N)])


def _tap(g, j):
    return g[j] if 0 <= j < K else 0.0


def _rows_small(N, z):
    """Modulated rows v[m][h][e] (half h, element e) of a [N, L] band block; zero outside [0, L)."""
    dmin, _ = _geometry(N)
    jb = PAD + N * (dmin - 1) + 1
    ra = jb % (2 * N)
    c = _mod_small(N)
    L = z.shape[1]

    def row(m):
        out = np.zeros((2, N))
        if 0 <= m < L:
            for h in range(2):
                for e in range(N):
                    out[h, e] = c[:, (ra + h * N + e) % (2 * N)] @ z[:, m].astype(np.float64)
        return out

    return row


def test_n3_packed_pairing_reproduces_the_oracle():
    """k_pqmf_synthesis_n3p: (y[q][2], y[q][1]) += (g[3d-1], g[3d]) (half[0], half[1]) and, for even q,
    (y[q][0], y[q+1][0]) += (g[3d+1], g[3d-2]) (half_d[2], half_{d-1}[2]) with the pair read from the row in the order
    its parity dictates: plane 1 holds (a2, b2) on even rows and (b2, a2) on odd rows."""
    N, Q = 3, 8
    dmin, halo = _geometry(N)
    assert (dmin, halo) == (-10, 21) and PAD + N * (dmin - 1) + 1 == -1  # taps 3d-1 .. 3d+1
    _, G = OP.design(N)
    g = _proto_signed(N)
    rng = np.random.default_rng(3)
    L = 40  # five blocks of Q steps
    z = rng.uniform(-1, 1, (N, L)).astype(np.float32)
    row = _rows_small(N, z)
    y = np.zeros((L, N))
    for n0 in range(0, L, Q):  # one thread: Q consecutive steps, rows n0 + dmin + i
        for i in range(Q + halo):
            v = row(n0 + dmin + i)
            plane0 = (v[0, 0], v[0, 1], v[1, 0], v[1, 1])  # a0 a1 | b0 b1
            plane1 = (v[0, 2], v[1, 2]) if i % 2 == 0 else (v[1, 2], v[0, 2])  # Q even: tile-row parity == i parity
            for q in range(Q):
                d = i - q
                if 0 <= d <= halo:
                    lo = plane0[2 * (d & 1):2 * (d & 1) + 2]
                    y[n0 + q, 2] += _tap(g, 3 * d - 1) * lo[0]
                    y[n0 + q, 1] += _tap(g, 3 * d) * lo[1]
            for q in range(0, Q, 2):
                d = i - q
                if 0 <= d <= halo:  # d == halo + 1 would pair two zero taps
                    y[n0 + q, 0] += _tap(g, 3 * d + 1) * plane1[0]
                    y[n0 + q + 1, 0] += _tap(g, 3 * d - 2) * plane1[1]
    ref = OP.synthesis(z[None], G, N)[0]
    assert OP.rel_err(y.reshape(-1).astype(np.float32), ref) <= 1e-6


def test_n4_same_parity_steps_read_one_half_per_row():
    """k_pqmf_synthesis_n4p: taps 4d .. 4d+3 on phases 3 .. 0 read row half d & 1, stored reversed; a thread that owns the
    steps s0 + par + 2q only ever meets row j of its window at d = j - 2q, i.e. on half j & 1."""
    N, Q = 4, 8
    dmin, halo = _geometry(N)
    assert (dmin, halo) == (-7, 15) and PAD + N * (dmin - 1) + 1 == 0  # taps 4d .. 4d+3
    _, G = OP.design(N)
    g = _proto_signed(N)
    rng = np.random.default_rng(4)
    L = 48  # three lane pairs of 2Q steps
    z = rng.uniform(-1, 1, (N, L)).astype(np.float32)
    row = _rows_small(N, z)
    y = np.zeros((L, N))
    for pr in range(L // (2 * Q)):
        for par in range(2):
            s0 = pr * 2 * Q + par
            for j in range(2 * (Q - 1) + halo + 1):
                v = row(s0 + dmin + j)[j & 1][::-1]  # the one half this thread reads of the row: (v3, v2, v1, v0)
                for q in range(Q):
                    d = j - 2 * q
                    if 0 <= d <= halo:
                        s = s0 + 2 * q
                        y[s, 0] += _tap(g, 4 * d + 3) * v[0]
                        y[s, 1] += _tap(g, 4 * d + 2) * v[1]
                        y[s, 2] += _tap(g, 4 * d + 1) * v[2]
                        y[s, 3] += _tap(g, 4 * d) * v[3]
    ref = OP.synthesis(z[None], G, N)[0]
    assert OP.rel_err(y.reshape(-1).astype(np.float32), ref) <= 1e-6


@pytest.mark.parametrize("N", [8, 16])
def test_two_step_ownership_for_large_n(N):
    """k_pqmf_synthesis_cm2: tap block rho (taps N rho .. N rho + N - 1) meets row s + rho on half rho & 1; with the half
    stored reversed, acc[k] += g[N rho + N-1-k] * half'[k]; steps s and s + 2 share row s + j at rho = j and j - 2."""
    dmin, halo = _geometry(N)
    assert _phase(0, N) == N - 1 and _offset(N, N) - _offset(0, N) == 1
    _, G = OP.design(N)
    g = _proto_signed(N)
    k = np.arange(N)[:, None]
    m = np.arange(N)[None, :]
    C = np.cos((2 * k + 1) * (2 * m + 1) * np.pi / (4 * N)) * N / np.sqrt(2)
    rng = np.random.default_rng(N)
    L = 24
    z = rng.uniform(-1, 1, (N, L)).astype(np.float32)
    U = np.einsum("km,kl->ml", C, z.astype(np.float64))
    v = np.zeros((2 * N, L))
    for r in range(2 * N):
        t2 = 2 * r - (K - 2)
        w = ((t2 + 2 * N) % (4 * N) + 4 * N) % (4 * N) - 2 * N
        mp, pos, neg = (abs(w) - 1) // 2, w > 0, ((w - t2) // (4 * N)) & 1
        val = U[mp] + (1 if pos else -1) * U[N - 1 - mp]
        v[r] = -val if neg else val

    def half_reversed(mrow, h):
        return v[N * h:N * h + N, mrow][::-1] if 0 <= mrow < L else np.zeros(N)

    y = np.zeros((L, N))
    for blk in range(L // 4):
        for par in range(2):
            sA = 4 * blk + par
            for j in range(halo + 3):
                hp = half_reversed(sA + dmin + j, j & 1)
                for q in range(2):
                    rho = j - 2 * q
                    if 0 <= rho <= halo:
                        for kk in range(N):
                            y[sA + 2 * q, kk] += _tap(g, N * rho + N - 1 - kk) * hp[kk]
    ref = OP.synthesis(z[None], G, N)[0]
    assert OP.rel_err(y.reshape(-1).astype(np.float32), ref) <= 1e-6


# ---- bank mapping -------------------------------------------------------------------------------------------------

def _conflicts(unit_lists, banks):
    """Number of lane groups (each served in one pass when conflict free) whose units collide modulo `banks`."""
    return sum(1 for us in unit_lists if len({u % banks for u in us}) < len(us))


def _injective(unit, rows, width=1):
    spans = sorted((unit(r), unit(r) + width) for r in range(rows))
    return all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))


def test_n3_planes_are_conflict_free():
    Q, halo = 8, 21
    rows = THREADS * Q + halo
    u0 = lambda r: r + r // Q  # noqa: E731  plane 0, 16-byte units
    u1 = lambda r: r + r // (2 * Q) + 8 * (r // (16 * Q))  # noqa: E731  plane 1, 8-byte units
    passes = (rows + THREADS - 1) // THREADS
    st0 = [[u0(t + THREADS * p) for t in range(8 * g, 8 * g + 8) if t + THREADS * p < rows]
           for p in range(passes) for g in range(THREADS // 8)]
    ld0 = [[u0(Q * t + i) for t in range(8 * g, 8 * g + 8)] for i in range(Q + halo) for g in range(THREADS // 8)]
    st1 = [[u1(t + THREADS * p) for t in range(16 * g, 16 * g + 16) if t + THREADS * p < rows]
           for p in range(passes) for g in range(THREADS // 16)]
    ld1 = [[u1(Q * t + i) for t in range(16 * g, 16 * g + 16)] for i in range(Q + halo) for g in range(THREADS // 16)]
    assert _conflicts(st0, 8) == 0 and _conflicts(ld0, 8) == 0
    assert _conflicts(st1, 16) == 0 and _conflicts(ld1, 16) == 0
    assert _injective(u0, rows) and _injective(u1, rows)
    # the two paddings that were measured before this one (ncu source pages, runs r4y and r4z): stores, then loads
    u1_r4y = lambda r: r + r // Q  # noqa: E731
    u1_r4z = lambda r: r + r // (2 * Q)  # noqa: E731
    assert _conflicts([[u1_r4y(t) for t in range(16 * g, 16 * g + 16)] for g in range(THREADS // 16)], 16) > 0
    assert _conflicts([[u1_r4z(Q * t + i) for t in range(16 * g, 16 * g + 16)]
                       for i in range(Q + halo) for g in range(THREADS // 16)], 16) > 0
    # the phase-2 bases of the kernel: u1(Q (t + a)) + i' for row Q t + Q a + i'
    for t in range(THREADS):
        for i in range(Q + halo):
            assert u1(Q * t + i) == u1(Q * (t + i // Q)) + i % Q


def test_n4_planes_are_conflict_free():
    Q, halo = 8, 15
    G = 2 * Q
    rows = THREADS * Q + halo
    unit = lambda r: r + 2 * (r // G)  # noqa: E731
    passes = (rows + THREADS - 1) // THREADS
    st = [[unit(t + THREADS * p) for t in range(8 * g, 8 * g + 8) if t + THREADS * p < rows]
          for p in range(passes) for g in range(THREADS // 8)]
    ld = [[unit(G * (t >> 1) + (t & 1) + j) for t in range(8 * g, 8 * g + 8)]
          for j in range(2 * (Q - 1) + halo + 1) for g in range(THREADS // 8)]
    assert _conflicts(st, 8) == 0 and _conflicts(ld, 8) == 0 and _injective(unit, rows)
    assert max(unit(r) for r in range(rows)) < rows + 2 * (rows // G) + 2  # UNITS of the kernel
    for t in range(THREADS):  # the kernel's constant offsets + the odd lane's bump
        pr, par = t >> 1, t & 1
        for j in range(2 * (Q - 1) + halo + 1):
            assert unit(G * pr + par + j) == (G + 2) * pr + par + j + 2 * (j // G) + (2 * par if j % G == G - 1 else 0)
    assert G * (THREADS // 2 - 1) + 1 + 2 * (Q - 1) + halo <= rows - 1  # the last thread's window stays inside the tile


@pytest.mark.parametrize("N", [8, 16])
def test_two_step_rows_are_conflict_free(N):
    _, halo = _geometry(N)
    row_units = 2 * N // 4
    unit = lambda r: (row_units + 1) * r + 2 * (r // 8) + 4 * (r // 16)  # noqa: E731
    tile = (THREADS - halo) // 4 * 4
    active = tile // 2
    st = [[unit(t) + c for t in range(8 * g, 8 * g + 8)] for c in range(row_units) for g in range(THREADS // 8)]
    ld = [[unit(4 * (t >> 1) + (t & 1) + j) + (j & 1) * (N // 4) + c for t in range(8 * g, 8 * g + 8) if t < active]
          for j in range(halo + 3) for c in range(N // 4) for g in range(THREADS // 8)]
    assert _conflicts(st, 8) == 0 and _conflicts(ld, 8) == 0
    assert _injective(unit, THREADS, row_units)
    assert 4 * ((active - 1) >> 1) + 1 + halo + 2 <= THREADS - 1  # every row a thread reads was modulated by the CTA
