"""Oracle: confidence intervals and the permutation test (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows /root/reference/main.py:259-278 (``bootstrap_confidence_interval``) and :281-302 (``permutation_test``).
The numpy arithmetic those functions lean on (``np.mean`` pairwise summation, ``np.percentile`` linear
interpolation) is restated explicitly here because the CUDA kernels reproduce it bit for bit.
The reference draws from the GLOBAL numpy stream; every function takes an optional ``rng`` (a
``numpy.random.RandomState``) and uses ``np.random`` itself when it is None, so both usages are reproducible.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

N_BOOTSTRAP = 10000      # main.py:265
N_PERMUTATIONS = 10000   # main.py:281


def _rng(rng):
    return np.random if rng is None else rng


def pairwise_sum(a) -> float:
    """numpy's float64 add-reduce over a contiguous 1-D array (what np.mean / np.std use).

    n < 8: sequential. n <= 128: eight interleaved accumulators r[j] += a[i+j], combined
    ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail sequentially. n > 128: split at n2 = n//2 rounded down to a
    multiple of 8 and recurse. (numpy/_core/src/umath/loops_utils.h.src, *_pairwise_sum.)"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = a.shape[0]
    if n < 8:
        res = 0.0
        for i in range(n):
            res = res + float(a[i])
        return res
    if n <= 128:
        r = [float(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = r[j] + float(a[i + j])
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + float(a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise_sum(a[:n2]) + pairwise_sum(a[n2:])


def np_mean(a) -> float:
    """np.mean of a 1-D float64 array = pairwise_sum / n."""
    a = np.asarray(a, dtype=np.float64)
    return pairwise_sum(a) / a.shape[0]


def np_std(a) -> float:
    """np.std (ddof=0, main.py:651) = sqrt(pairwise_sum((a-mean)**2)/n)."""
    a = np.asarray(a, dtype=np.float64)
    d = a - np_mean(a)
    return math.sqrt(pairwise_sum(d * d) / a.shape[0])


def percentile_linear(x, q: float) -> float:
    """np.percentile(x, q) with the default 'linear' method: sort; vi = (n-1)*(q/100); lo = floor(vi); g = vi-lo;
    lerp(a, b, g) = a + (b-a)*g if g < 0.5 else b - (b-a)*(1-g)  (numpy/lib/_function_base_impl.py _lerp)."""
    xs = np.sort(np.asarray(x, dtype=np.float64))
    n = xs.shape[0]
    vi = (n - 1) * (q / 100.0)
    lo = int(math.floor(vi))
    g = vi - lo
    hi = min(lo + 1, n - 1)
    a, b = float(xs[lo]), float(xs[hi])
    d = b - a
    return a + d * g if g < 0.5 else b - d * (1.0 - g)


def ci_resample_indices(n: int, rng=None, n_bootstrap: int = N_BOOTSTRAP) -> np.ndarray:
    """The index stream behind main.py:268-269: ``np.random.choice(data, size=n)`` n_bootstrap times is the same
    stream as one ``randint(0, n, size=(n_bootstrap, n))`` (legacy MT19937 masked rejection)."""
    return _rng(rng).randint(0, n, size=(n_bootstrap, n))


def bootstrap_confidence_interval(data, confidence: float = 0.95, rng=None) -> Tuple[float, float, float]:
    """main.py:259-278. Returns (mean, lower, upper)."""
    data = np.asarray(data)
    n = len(data)
    idx = ci_resample_indices(n, rng)
    means = np.empty(N_BOOTSTRAP, dtype=np.float64)
    for r in range(N_BOOTSTRAP):
        means[r] = np.mean(data[idx[r]])          # == np_mean (checked in tests)
    mean = np.mean(data)
    alpha = (1 - confidence) / 2
    lower = np.percentile(means, alpha * 100)
    upper = np.percentile(means, (1 - alpha) * 100)
    return mean, lower, upper


def permutation_indices(n_total: int, rng=None, n_permutations: int = N_PERMUTATIONS) -> np.ndarray:
    """main.py:295-296 shuffles ``combined`` in place and cumulatively; shuffling an index array with the same
    stream gives combined_after_p == combined0[perm[p]]. Returns int32 [P, n_total]."""
    r = _rng(rng)
    idx = np.arange(n_total, dtype=np.int64)
    out = np.empty((n_permutations, n_total), dtype=np.int32)
    for p in range(n_permutations):
        r.shuffle(idx)
        out[p] = idx
    return out


def permutation_test(data1, data2, n_permutations: int = N_PERMUTATIONS, rng=None) -> float:
    """main.py:281-302, literal."""
    r = _rng(rng)
    observed_diff = np.abs(np.mean(data1) - np.mean(data2))
    combined = np.concatenate([data1, data2])
    n1 = len(data1)
    diffs = []
    for _ in range(n_permutations):
        r.shuffle(combined)
        diffs.append(np.abs(np.mean(combined[:n1]) - np.mean(combined[n1:])))
    return np.mean(np.array(diffs) >= observed_diff)
