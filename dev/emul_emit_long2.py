"""numpy emulation of emit_long_kernel v2 (csrc/shuffle.cu): one CTA of T threads, every thread owns S consecutive
words of the epoch's window (S odd); the accepts of a segment are non-increasing in the number of accepts behind it,
so a run from the upper base Ub bounds them from below and a run from the lower base Lb from above; one block-wide scan
of the packed counts per round gives the next bases (Jacobi across segments), convergence =
every thread whose lower base is below the epoch's step count has an exact base. Checked against numpy's shuffle
(permutations and stream position) before the kernel first ran on a GPU.

    python dev/emul_emit_long2.py
"""
import math
import sys

import numpy as np

T = 1024
RING = 57344
WCAP = RING - 8
SERIAL_BELOW = 64          # steps with i < 64 are walked by one thread


def plan_epochs(n):
    """(i_hi, c, M, W0) per mask epoch, top down; W0 = expected words + 8 sigma + 64."""
    out = []
    i_hi = n - 1
    while i_hi >= SERIAL_BELOW:
        b = int(i_hi).bit_length()
        M = (1 << b) - 1
        i_lo = max(1 << (b - 1), SERIAL_BELOW)
        i = np.arange(i_lo, i_hi + 1, dtype=np.float64)
        p = (i + 1.0) / (M + 1.0)
        mean = float(np.sum(1.0 / p))
        sig = math.sqrt(float(np.sum((1.0 - p) / p ** 2)))
        W0 = min(WCAP, int(mean + 8.0 * sig) + 64)
        out.append((i_hi, i_hi - i_lo + 1, M, W0))
        i_hi = i_lo - 1
    return out, i_hi        # i_hi = first step of the serial tail (SERIAL_BELOW - 1, or n - 1 if smaller)


def epoch(words, pos, n_words, i_hi, c, M, W0, stats):
    W = W0
    while True:
        W = min(W, n_words - pos)
        if W < c:
            return None, -1
        S = -(-W // T)
        S |= 1
        v = np.full(T * S, 1 << 40, dtype=np.int64)
        v[:W] = words[pos:pos + W].astype(np.int64) & M
        v = v.reshape(T, S)
        t0 = np.arange(T, dtype=np.int64) * S
        Lb = np.zeros(T, dtype=np.int64)
        Ub = np.minimum(t0, W)
        rnd = 0
        lens = np.clip(W - t0, 0, S)
        have_cnt = np.zeros(T, dtype=bool)               # frozen segments: the two runs accepted the same words
        cnt = np.zeros(T, dtype=np.int64)
        cls = np.zeros(T, dtype=bool)                    # classified segments: sure count + up to UNC uncertain thresholds
        sure_n = np.zeros(T, dtype=np.int64)
        ent = [[] for _ in range(T)]
        UNC = 4
        while True:
            lo, hi = Ub.copy(), Lb.copy()                   # two independent runs: from the upper and the lower base
            mU = np.zeros((T, S), dtype=bool)
            mL = np.zeros((T, S), dtype=bool)
            for s in range(S):
                au = (v[:, s] + lo) <= i_hi
                al = (v[:, s] + hi) <= i_hi
                mU[:, s], mL[:, s] = au, al
                lo += au
                hi += al
            cl, cu = lo - Ub, hi - Lb                        # accepts are non-increasing in the start value
            for k in np.nonzero(cls & ~have_cnt & (Lb < c))[0]:   # classified: runs over the uncertain words only
                xa, xb = Ub[k], Lb[k]
                for th in ent[k]:
                    xa += xa <= th
                    xb += xb <= th
                assert sure_n[k] + xa - Ub[k] == cl[k] and sure_n[k] + xb - Lb[k] == cu[k], (k, rnd)
            newly = (~have_cnt) & (cl == cu) & (Lb < c)
            cl = np.where(have_cnt, cnt, cl)                 # (the kernel does not walk frozen segments again)
            cu = np.where(have_cnt, cnt, cu)
            cnt = np.where(newly, cl, cnt)
            have_cnt |= newly
            if rnd >= 1:
                for k in np.nonzero(~cls & ~have_cnt & (Lb < c) & (Lb != Ub))[0]:
                    unc = mL[k] & ~mU[k]
                    if 0 < unc.sum() <= UNC:
                        cls[k] = True
                        sure_n[k] = mU[k].sum()
                        ent[k] = [int(i_hi - v[k, j] - mU[k, :j].sum()) for j in np.nonzero(unc)[0]]
            behind = Lb >= c                                 # behind the epoch's end: (0, len) is a valid bracket
            cl = np.where(behind, 0, cl)
            cu = np.where(behind, lens, cu)
            assert (cl <= cu).all()
            pred = (Lb >= c) | (Lb == Ub) | (lens == 0)
            if pred.all():
                break
            Lb = np.maximum(Lb, np.cumsum(cl) - cl)          # keep the brackets nested
            Ub = np.minimum(Ub, np.cumsum(cu) - cu)
            rnd += 1
            assert rnd <= T
        stats.append(rnd)
        # emit with exact bases
        out = np.full(c, -1, dtype=np.int64)
        end = -1
        l = Lb.copy()
        for s in range(S):
            acc = ((v[:, s] + l) <= i_hi) & (Lb < c)
            w = acc & (l < c)
            out[l[w]] = v[w, s]
            last = w & (l == c - 1)
            if last.any():
                assert last.sum() == 1
                end = int(t0[last][0]) + s + 1
            l += acc
        if end >= 0:
            assert (out >= 0).all()
            return out, end
        if W >= WCAP or pos + W >= n_words:
            return None, (-3 if W >= WCAP else -1)
        W = min(WCAP, 2 * W)


def emit_long(words, n, P):
    n_words = len(words) & ~3
    plan, i_tail = plan_epochs(n)
    jout = np.zeros(P * (n - 1), dtype=np.int64)
    pos, stats = 0, []
    for p in range(P):
        base = p * (n - 1)
        for (i_hi, c, M, W0) in plan:
            out, end = epoch(words, pos, n_words, i_hi, c, M, W0, stats)
            if out is None:
                return None, end, stats
            jout[base + (n - 1 - i_hi): base + (n - 1 - i_hi) + c] = out
            pos += end
        i = i_tail
        while i >= 1:                                       # serial tail (thread 0)
            if pos >= n_words:
                return None, -1, stats
            mask = (1 << int(i).bit_length()) - 1
            vv = int(words[pos]) & mask
            pos += 1
            if vv <= i:
                jout[base + (n - 1 - i)] = vv
                i -= 1
    return jout, pos, stats


def check(n, P, seed=7):
    ref = np.random.RandomState(seed)
    raw = np.random.RandomState(seed)
    words = raw._bit_generator.random_raw(int(P * n * 2.0) + 100000).astype(np.uint32)
    jout, pos, st = emit_long(words, n, P)
    assert jout is not None, pos
    x, xr = np.arange(n), np.arange(n)
    ok = True
    for p in range(P):
        ref.shuffle(xr)
        js = jout[p * (n - 1):(p + 1) * (n - 1)]
        for k, i in enumerate(range(n - 1, 0, -1)):
            j = js[k]
            x[i], x[j] = x[j], x[i]
        ok = ok and np.array_equal(x, xr)
    chk = np.random.RandomState(seed)
    chk._bit_generator.random_raw(pos)
    same_state = chk.randint(0, 1 << 30) == ref.randint(0, 1 << 30)
    print(n, P, ok, same_state, "rounds max", max(st) if st else 0, "mean %.2f" % (np.mean(st) if st else 0),
          "sum/perm", sum(st) / P)
    assert ok and same_state


if __name__ == "__main__":
    for n, P in ((50000, 2), (20000, 3), (16385, 2), (65536, 1), (1500, 20), (100, 50), (2, 5), (64, 9), (65, 9)):
        check(n, P)
    words = np.random.RandomState(1)._bit_generator.random_raw(30000).astype(np.uint32)
    print("short stream ->", emit_long(words, 20000, 2)[1])
