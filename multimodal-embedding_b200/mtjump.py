"""MT19937 jump-ahead polynomials for K6 (host-side, computed once and cached next to libmmx.so).

numpy's legacy generator is a linear recurrence over GF(2) with characteristic polynomial phi of degree 19937. For
g_J(x) = x^J mod phi(x) every word of the raw stream satisfies  x[n + J] = XOR_{i : g_J[i] = 1} x[n + i],  so the
624-word block that precedes raw offset J is a convolution of the first 19937 + 623 raw words with g_J — an
embarrassingly parallel job — instead of J sequential steps. The table holds g for J = k * SEG_WORDS, k = 1..N_POLY,
which lets `mmx_mt19937_randint_par` start one CTA per 64-block segment of the stream (296 segments = one wave of
2 CTAs per SM on a B200; longer requests are chained launches).

phi comes from Berlekamp-Massey on one output bit (tempering is a linear bijection, so any output bit is a linear
recurring sequence with the same minimal polynomial); everything is plain Python integers used as GF(2)[x] bit
vectors (bit i = coefficient of x^i). ~15 s once (cached as csrc/mt19937_jump_table.npy by build()); verified against
numpy's own sequential stream in tests/test_mtjump.py.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

DEGREE = 19937
N = 624
SEG_BLOCKS = 64                   # numpy blocks (of 624 raw words) per device segment
SEG_WORDS = N * SEG_BLOCKS
N_POLY = 295                      # 296 segments per launch = 1.18e7 raw words
TABLE_PATH = Path(__file__).resolve().parent / "csrc" / "mt19937_jump_table.npy"
# Long requests (the 9.5e7 draws of a dense model's row, the 1e9 of a 10^6-image bootstrap) take segments of 512 blocks
# with a second table: the jump convolution costs the same per SEGMENT (~0.3 ms per launch of 296), so 8x longer
# segments make it 8x cheaper per word (it was 85 % of the launch time at 64 blocks).
SEG_BLOCKS_LONG = 512
SEG_WORDS_LONG = N * SEG_BLOCKS_LONG
TABLE_PATH_LONG = Path(__file__).resolve().parent / "csrc" / "mt19937_jump_table_long.npy"


def berlekamp_massey(bits: np.ndarray) -> int:
    """Connection polynomial C (bit i = c_i, c_0 = 1) of the shortest LFSR generating `bits` over GF(2)."""
    n = len(bits)
    s = 0                                  # bits seen so far, newest at bit 0: bit j = s_{k-j}
    c, b = 1, 1
    L, m = 0, 1
    for k in range(n):
        s = (s << 1) | int(bits[k])
        # discrepancy d = sum_{i=0..L} c_i s_{k-i}
        d = ((c & s) & ((1 << (L + 1)) - 1)).bit_count() & 1
        if d:
            t = c
            c ^= b << m
            if 2 * L <= k:
                L, b, m = k + 1 - L, t, 1
            else:
                m += 1
        else:
            m += 1
    return c, L


def characteristic_polynomial() -> int:
    """phi(x) with sum_i phi_i x[n+i] = 0 for the MT19937 word stream (phi_DEGREE = 1)."""
    rs = np.random.RandomState(20240229)
    bits = (rs._bit_generator.random_raw(2 * DEGREE + 64) & 1).astype(np.uint8)
    c, L = berlekamp_massey(bits)
    if L != DEGREE:
        raise RuntimeError(f"unexpected linear complexity {L}")
    phi = 0                                 # reciprocal: phi_j = c_{L-j}
    for j in range(L + 1):
        if (c >> (L - j)) & 1:
            phi |= 1 << j
    return phi


def _mod(p: int, phi: int) -> int:
    dp = p.bit_length() - 1
    while dp >= DEGREE:
        p ^= phi << (dp - DEGREE)
        dp = p.bit_length() - 1
    return p


_SPREAD = None


def _square(p: int) -> int:
    """p(x)^2 over GF(2) = bit interleaving with zeros."""
    global _SPREAD
    if _SPREAD is None:
        _SPREAD = [sum(((v >> bit) & 1) << (2 * bit) for bit in range(8)) for v in range(256)]
    data = p.to_bytes((p.bit_length() + 7) // 8 or 1, "little")
    out = bytearray(2 * len(data))
    for i, v in enumerate(data):
        sv = _SPREAD[v]
        out[2 * i] = sv & 0xFF
        out[2 * i + 1] = sv >> 8
    return int.from_bytes(out, "little")


def _mul(a: int, b: int, phi: int) -> int:
    """a * b mod phi (schoolbook over the set bits of the sparser operand)."""
    if a.bit_count() > b.bit_count():
        a, b = b, a
    r = 0
    while a:
        low = a & -a
        r ^= b << (low.bit_length() - 1)
        a ^= low
    return _mod(r, phi)


def pow_x(e: int, phi: int) -> int:
    """x^e mod phi by square-and-shift."""
    r = 1
    for bit in format(e, "b"):
        r = _mod(_square(r), phi)
        if bit == "1":
            r = _mod(r << 1, phi)
    return r


def poly_to_words(p: int) -> np.ndarray:
    return np.frombuffer(p.to_bytes(N * 4, "little"), dtype=np.uint32).copy()


def build_table(n_poly: int = N_POLY, seg_words: int = SEG_WORDS) -> np.ndarray:
    phi = characteristic_polynomial()
    g1 = pow_x(seg_words, phi)
    table = np.zeros((n_poly, N), dtype=np.uint32)
    g = g1
    for k in range(n_poly):
        table[k] = poly_to_words(g)
        if k + 1 < n_poly:
            g = _mul(g, g1, phi)
    return table


def load_table(build_if_missing: bool = True, long: bool = False) -> np.ndarray:
    """The table for 64-block segments, or (long=True) for 512-block segments."""
    path, seg_words = (TABLE_PATH_LONG, SEG_WORDS_LONG) if long else (TABLE_PATH, SEG_WORDS)
    if path.exists():
        t = np.load(path)
        if t.shape == (N_POLY, N) and t.dtype == np.uint32:
            return t
    if not build_if_missing:
        raise FileNotFoundError(path)
    t = build_table(N_POLY, seg_words)
    np.save(path, t)
    return t


def jump_key(key: np.ndarray, g_words: np.ndarray) -> np.ndarray:
    """CPU restatement of the device convolution (test infrastructure for tests/test_mtjump.py): the 624-word block
    J raw words after `key`, given g = x^J mod phi as 624 uint32 words. Word 0 is only meaningful in its top bit
    (MT19937 never reads the low 31 bits of the first state word)."""
    x = np.zeros(33 * N, dtype=np.uint32)
    x[:N] = key
    up, lo, a = np.uint32(0x80000000), np.uint32(0x7FFFFFFF), np.uint32(0x9908B0DF)
    for n in range(N, 33 * N):
        y = (x[n - N] & up) | (x[n - N + 1] & lo)
        x[n] = x[n - N + 397] ^ (y >> np.uint32(1)) ^ (a if (y & np.uint32(1)) else np.uint32(0))
    bits = np.unpackbits(g_words.view(np.uint8), bitorder="little")
    out = np.zeros(N, dtype=np.uint32)
    for i in np.nonzero(bits)[0]:
        out ^= x[i:i + N]
    return out


PAD_POSITION = 33 * N     # index of a zero word behind the 33-block stream the device convolves (csrc/rng.cu XPAD)


def bit_positions(table: np.ndarray):
    """Set-bit positions of every polynomial: (uint16 [n_poly, ld] padded with PAD_POSITION, ld % 4 == 0; int32 counts)."""
    bits = np.unpackbits(np.ascontiguousarray(table).view(np.uint8), axis=1, bitorder="little")
    cnt = bits.sum(axis=1).astype(np.int32)
    ld = (int(cnt.max()) + 3) & ~3
    pos = np.full((table.shape[0], ld), PAD_POSITION, dtype=np.uint16)
    for k in range(table.shape[0]):
        idx = np.nonzero(bits[k])[0]
        pos[k, : len(idx)] = idx
    return pos, cnt
