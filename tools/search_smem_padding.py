#!/usr/bin/env python
"""Exhaustive search for the shared-memory row padding of k_pqmf_synthesis_cm2 (csrc/pqmf.cu), CPU only.

Row r of the modulated tile starts at 16-byte unit  pitch * r + c4 (r/4) + c8 (r/8) + c16 (r/16) + c32 (r/32).
Constraints (bank model measured with ncu on the B200: a 128-bit shared access is served per quarter warp, eight
lanes over eight 16-byte bank groups):
  * phase 1: the eight consecutive rows a quarter warp stores (one row per lane) start in eight distinct bank groups;
  * phase 2: lanes 2p / 2p+1 own the even / odd steps of a four-step block and read row 4p + par + j for
    j = 0 .. HALO + 2 -- the eight lanes of every quarter warp of the active threads start in distinct bank groups;
  * rows do not overlap.
Prints the smallest tile found per band count: (units, pitch, (c4, c8, c16, c32)).  The kernel uses
pitch = 2N/4 + 1, (0, 2, 4, 0) for N = 8 and N = 16.      python tools/search_smem_padding.py"""
import itertools

THREADS = 128


def search(N, halo, steps_per_thread=2):
    blk = 2 * steps_per_thread
    row_units = 2 * N // 4
    active = (THREADS - halo) // 4 * 4 // 2
    best = None
    for pitch in (row_units, row_units + 1, row_units + 2):
        for c in itertools.product(range(6), repeat=4):
            unit = lambda r, c=c, pitch=pitch: pitch * r + c[0] * (r // 4) + c[1] * (r // 8) + c[2] * (r // 16) + c[3] * (r // 32)  # noqa: E731
            if any(unit(r + 1) < unit(r) + row_units for r in range(THREADS + 2)):
                continue
            ok = all(len({unit(t) % 8 for t in range(8 * g, 8 * g + 8)}) == 8 for g in range(THREADS // 8))
            for j in range(2 * (steps_per_thread - 1) + halo + 1):
                if not ok:
                    break
                for g in range(THREADS // 8):
                    lanes = [t for t in range(8 * g, 8 * g + 8) if t < active]
                    if len({unit(blk * (t >> 1) + (t & 1) + j) % 8 for t in lanes}) < len(lanes):
                        ok = False
                        break
            if ok:
                size = unit(THREADS - 1) + row_units
                if best is None or size < best[0]:
                    best = (size, pitch, c)
    return best


if __name__ == "__main__":
    for N, halo in ((16, 3), (8, 7)):
        print(f"N={N} HALO={halo}:", search(N, halo))
