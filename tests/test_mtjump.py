"""K6p host side: the MT19937 jump-ahead polynomials (multimodal-embedding_b200/mtjump.py) against numpy's own
sequential stream. The draws they serve are np.random.choice at /root/reference/main.py:588,269."""
import importlib

import numpy as np

mtjump = importlib.import_module("multimodal-embedding_b200.mtjump")


def _key_after_blocks(seed: int, blocks: int) -> np.ndarray:
    """numpy's key after `blocks` regenerations from the seeded state."""
    rs = np.random.RandomState(seed)
    st = rs.get_state()
    rs.set_state(("MT19937", st[1], 624, 0, 0.0))            # the next draw regenerates
    left = blocks * 624
    while left:
        c = min(left, 1 << 20)
        rs._bit_generator.random_raw(c)
        left -= c
    return np.asarray(rs.get_state()[1], dtype=np.uint32)


def test_characteristic_polynomial_degree():
    phi = mtjump.characteristic_polynomial()
    assert phi.bit_length() - 1 == mtjump.DEGREE and phi & 1


def test_jump_matches_sequential_stream():
    phi = mtjump.characteristic_polynomial()
    key = np.asarray(np.random.RandomState(5).get_state()[1], dtype=np.uint32)
    g1 = mtjump.pow_x(mtjump.SEG_WORDS, phi)
    g3 = mtjump._mul(mtjump._mul(g1, g1, phi), g1, phi)
    for k, g in ((1, g1), (3, g3)):
        got = mtjump.jump_key(key, mtjump.poly_to_words(g))
        want = _key_after_blocks(5, k * mtjump.SEG_BLOCKS)
        assert np.array_equal(got[1:], want[1:])
        assert (got[0] ^ want[0]) & np.uint32(0x80000000) == 0     # low 31 bits of word 0 are never read by MT19937


def test_table_rows_are_powers():
    t = mtjump.load_table()
    assert t.shape == (mtjump.N_POLY, mtjump.N) and t.dtype == np.uint32
    phi = mtjump.characteristic_polynomial()
    g1 = mtjump.pow_x(mtjump.SEG_WORDS, phi)
    assert np.array_equal(t[0], mtjump.poly_to_words(g1))
    k = mtjump.N_POLY
    assert np.array_equal(t[k - 1], mtjump.poly_to_words(mtjump.pow_x(k * mtjump.SEG_WORDS, phi)))


def test_long_table_rows_are_powers_and_jump():
    """The 512-block table (long requests): first and last rows are the right powers of x, and a jump with row 2 lands
    on numpy's key after 2 x 512 regenerations."""
    t = mtjump.load_table(long=True)
    assert t.shape == (mtjump.N_POLY, mtjump.N) and t.dtype == np.uint32
    phi = mtjump.characteristic_polynomial()
    assert np.array_equal(t[0], mtjump.poly_to_words(mtjump.pow_x(mtjump.SEG_WORDS_LONG, phi)))
    k = mtjump.N_POLY
    assert np.array_equal(t[k - 1], mtjump.poly_to_words(mtjump.pow_x(k * mtjump.SEG_WORDS_LONG, phi)))
    key = np.asarray(np.random.RandomState(9).get_state()[1], dtype=np.uint32)
    got = mtjump.jump_key(key, t[1])
    want = _key_after_blocks(9, 2 * mtjump.SEG_BLOCKS_LONG)
    assert np.array_equal(got[1:], want[1:]) and (got[0] ^ want[0]) & np.uint32(0x80000000) == 0
