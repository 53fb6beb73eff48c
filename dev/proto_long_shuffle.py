"""Prototype (numpy, CPU) of an exact PARALLEL-WITHIN-A-PERMUTATION replay of numpy's legacy RandomState.shuffle for
arrays longer than the 16 384 elements K6s handles (per-query vectors of 25 000-query models).

Fisher-Yates from the top: for i = n-1 .. 1 numpy draws j = random_interval(i) by masked rejection — raw words are
consumed until (word & mask(i)) <= i. Inside one "mask epoch" (all i with the same mask M = 2^b - 1, i from i_hi down
to i_lo = 2^(b-1)) the k-th accepted word serves target i_hi - k, so word t is accepted iff
        v_t <= i_hi - a(t),      v_t = word_t & M,   a(t) = number of accepted words before t.
a(t) is the only serial quantity. It is bracketed by L(t) <= a(t) <= U(t) and the brackets are tightened by
        surely accepted:  v_t <= i_hi - U(t)          surely rejected:  v_t > i_hi - L(t)
        L'(t) = #surely accepted before t             U'(t) = #not surely rejected before t
— two prefix sums per round, every round exact, until L == U on the prefix that holds the epoch's accepts. On a GPU a
round is a block-wide scan over the epoch's ~1.5 (i_hi - i_lo) words; the permutations themselves stay a serial chain
(each starts where the previous one stopped), so this is one CTA per test stream: ~16 epochs x ~10 rounds per
permutation instead of n dependent steps.

This script checks the scheme against numpy's own shuffle (permutation AND words consumed) and prints the rounds needed.
"""
import sys
import time

import numpy as np


def excl_cumsum(x):
    c = np.cumsum(x, dtype=np.int64)
    return c - x


def epoch_accepts(words, pos, i_hi, i_lo, stats):
    """Positions (absolute) and values of the first (i_hi - i_lo + 1) accepted words from `pos` on."""
    b = int(i_hi).bit_length()
    M = (1 << b) - 1
    c = i_hi - i_lo + 1
    W = int(c * (M + 1) / (i_lo + 1) * 1.15) + 256
    while True:
        v = (words[pos:pos + W] & np.uint32(M)).astype(np.int64)
        if len(v) < W and len(v) < c:
            raise RuntimeError("stream too short")
        L = np.zeros(len(v), dtype=np.int64)
        U = excl_cumsum((v <= i_hi).astype(np.int64))
        rounds = 0
        while True:
            rounds += 1
            sure_acc = v <= i_hi - U
            sure_rej = v > i_hi - L
            L2 = excl_cumsum(sure_acc.astype(np.int64))
            U2 = excl_cumsum((~sure_rej).astype(np.int64))
            # only the prefix up to the c-th sure accept matters
            done_at = np.searchsorted(L2 + sure_acc, c)          # first t with (accepts up to and incl. t) >= c
            if done_at < len(v) and np.array_equal(L2[:done_at + 1], U2[:done_at + 1]):
                acc = np.nonzero(sure_acc[:done_at + 1])[0]
                stats.append((b, c, done_at + 1, rounds))
                return pos + acc[:c], v[acc[:c]]
            if np.array_equal(L2, L) and np.array_equal(U2, U):
                break                                            # no progress inside this window: enlarge it
            L, U = L2, U2
        W *= 2


def shuffle_replay(words, pos, n, stats):
    """j values for i = n-1 .. 1 and the stream position after the shuffle."""
    js = np.empty(n - 1, dtype=np.int64)
    i_hi = n - 1
    k = 0
    while i_hi >= 1:
        i_lo = max(1 << (int(i_hi).bit_length() - 1), 1)
        p, vals = epoch_accepts(words, pos, i_hi, i_lo, stats)
        c = i_hi - i_lo + 1
        js[k:k + c] = vals
        k += c
        pos = int(p[-1]) + 1
        i_hi = i_lo - 1
    return js, pos


def apply(x, js):
    n = len(x)
    for k, i in enumerate(range(n - 1, 0, -1)):
        j = js[k]
        x[i], x[j] = x[j], x[i]


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    seed = 5
    ref = np.random.RandomState(seed)
    raw = np.random.RandomState(seed)
    words = raw._bit_generator.random_raw(int(P * n * 2.2) + 100_000).astype(np.uint32)
    # numpy reads 32-bit outputs one at a time for masks below 2^32
    x_ref = np.arange(n)
    x = np.arange(n)
    pos = 0
    stats = []
    t_np = t_proto = 0.0
    for p_ in range(P):
        t0 = time.perf_counter(); ref.shuffle(x_ref); t_np += time.perf_counter() - t0
        t0 = time.perf_counter(); js, pos = shuffle_replay(words, pos, n, stats); t_proto += time.perf_counter() - t0
        apply(x, js)
        assert np.array_equal(x, x_ref), f"permutation {p_} differs"
    # words consumed: numpy's generator must sit exactly `pos` raw words after the seed state
    chk = np.random.RandomState(seed)
    chk._bit_generator.random_raw(pos)
    assert chk.randint(0, 1 << 30) == ref.randint(0, 1 << 30), "stream position differs"
    big = [s for s in stats if s[1] >= min(1024, n // 4)]
    print(f"n={n} P={P}: permutations and stream position identical to numpy's shuffle")
    print(f"  large epochs (>= {min(1024, n // 4)} targets): rounds min/mean/max = {min(s[3] for s in big)}/"
          f"{np.mean([s[3] for s in big]):.1f}/{max(s[3] for s in big)}; words scanned per target "
          f"{np.mean([s[2] / s[1] for s in big]):.2f}")
    print(f"  all epochs: mean rounds {np.mean([s[3] for s in stats]):.1f}, epochs per permutation {len(stats) // P}")
    print(f"  (numpy shuffle {1e3 * t_np / P:.2f} ms per permutation on this host; the numpy prototype {1e3 * t_proto / P:.1f} ms)")
